"""Import shim: lets the reference's unmodified sources (``from torch_geometric.nn import SAGEConv``,
``from torch_geometric.loader import NeighborLoader`` — reference src/models/layers/sage.py:4, src/pipeline.py:6)
resolve to the B200-native drop-ins in ``noise_gnn_b200``.  Put ``<repo>/compat`` (and ``<repo>``) ahead of
site-packages on PYTHONPATH.  Only the two hot-path entry points and ``Data`` exist; everything else raises."""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

__version__ = "2.5.1+ngnn_b200"

from . import data, loader, nn  # noqa: E402,F401
