from noise_gnn_b200.loader import Batch, Data  # noqa: F401
