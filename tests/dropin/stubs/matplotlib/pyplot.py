class _Dummy:
    def __getattr__(self, name):
        return lambda *a, **k: _Dummy()

    def __iter__(self):
        return iter(())


def __getattr__(name):
    return lambda *a, **k: _Dummy()
