"""CPU oracle for the GraphSAGE mini-batch hot path of hhilsber/noise-GNN.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` may be imported, linked or executed by the
product path (``noise_gnn_b200/``); only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` use it, and only as the checker or the
timed CPU baseline.

PARITY UNPINNED.  The arithmetic of the reference's hot path lives in un-vendored third-party wheels
(torch-geometric==2.5.1 ``SAGEConv`` / ``NeighborLoader``, pyg-lib ``neighbor_sample``; reference
docs/requirements.txt:11, docs/commands.txt:14) that are not installable in this environment, and the
reference ships no tests, golden vectors or fixtures.  The oracle therefore
restates the published algorithms, anchored on the reference's call sites:

* ``sage_oracle``      SAGEConv / SAGE network / train step — the exact torch op sequence PyG lowers to
                       (index_select -> scatter_add_ -> count -> clamp(min=1) -> divide -> linear x2 -> add),
                       network structure verbatim from reference src/models/layers/sage.py:6-40,
                       train step from reference src/pipeline.py:144-173.
                       Also ``GCNConvRef`` / ``SimpleGCNRef``: PyG ``GCNConv(normalize=False)`` as built at reference
                       src/models/layers/convolution.py:19-23 (linear, then sum over in-neighbours, + bias).
* ``ct_oracle``        ``CTLoss.forward`` of reference src/utils/losses.py:19-49 restated line by line (stable argsort).
* ``structure``        COO -> CSR by destination (stable), CSR transpose — numpy.
* ``sampler``          ctypes binding of ``sampler_oracle.c`` (sequential fan-out sampler with the same
                       Philox4x32-10 stream as the CUDA sampler) plus a pure-Python twin for tiny cases.
* ``philox``           numpy Philox4x32-10, pinned to the Random123 known-answer vectors.

What is pinned: Philox against its published KAT vectors; SAGEConv against a hand-computed golden
block (tests/golden/); the C sampler against its pure-Python twin.
"""
