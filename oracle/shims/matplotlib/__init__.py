"""Permissive no-op stand-in for matplotlib (reference use: dgfem/visualization.py:16-23).
Plotting is out of scope; this only has to import."""


class _Anything:
    def __getattr__(self, name):
        return _Anything()

    def __call__(self, *a, **k):
        return _Anything()

    def __iter__(self):
        return iter(())

    def __getitem__(self, k):
        return _Anything()

    def __setitem__(self, k, v):
        pass


def __getattr__(name):
    return _Anything()
