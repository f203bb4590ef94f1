class Dummy:
    """Permissive stand-in: any attribute/call/item returns another Dummy."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return Dummy()

    def __getattr__(self, n):
        if n.startswith("__") and n.endswith("__"):
            raise AttributeError(n)
        return Dummy()

    def __getitem__(self, k):
        return Dummy()

    def __iter__(self):
        return iter(())
