from noise_gnn_b200.conv import SAGEConv  # noqa: F401


class GCNConv:  # imported (never constructed) by reference src/models/layers/gcn.py:3; built by convolution.py:19-23
    def __init__(self, *a, **k):
        raise NotImplementedError("GCNConv is outside the SAGE hot path this package replaces (SURVEY §8f rank 3)")
