def __getattr__(name):
    def _missing(*a, **k):
        raise NotImplementedError(f"torch_geometric.transforms.{name} is not provided by the hot-path shim")
    return _missing
