def __getattr__(name):  # Planetoid / CitationFull / Amazon need network access: use noise_gnn_b200.synthetic.make_dataset
    def _missing(*a, **k):
        raise NotImplementedError(f"torch_geometric.datasets.{name} is not provided; use noise_gnn_b200.synthetic.make_dataset")
    return _missing
