"""See matplotlib/__init__.py in this directory."""


class _Style:
    def use(self, *a, **k):
        return None


style = _Style()


def __getattr__(name):
    raise RuntimeError("matplotlib stub: plotting is not available in the build container (" + name + ")")
