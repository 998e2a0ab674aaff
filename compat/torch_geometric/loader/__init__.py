from noise_gnn_b200.loader import NeighborLoader  # noqa: F401


class NeighborSampler:  # imported but unused by the reference pipelines (src/pipeline.py:6)
    def __init__(self, *a, **k):
        raise NotImplementedError("NeighborSampler is not used by the reference's hot path; use NeighborLoader")
