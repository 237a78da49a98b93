"""TEST INFRASTRUCTURE ONLY -- live import of the read-only reference.

Loads /root/reference/solver.py and analyser.py *in this container* so that
the numpy restatement (oracle/restated.py) can be validated against the real
thing and golden vectors can be generated (oracle/make_golden.py).

The reference imports matplotlib at module top (solver.py:3-4,
analyser.py:2-3) and matplotlib is not installed here, so three stub modules
are planted in sys.modules first.  The reference is NOT present on the GPU
box: nothing under `-m gpu`, smoke() or bench.py may call into this file.
"""
import importlib
import os
import sys
import types

REFERENCE_DIR = os.environ.get("FDTD_REFERENCE_DIR", "/root/reference")


def available():
    return os.path.isfile(os.path.join(REFERENCE_DIR, "solver.py"))


class _Anything:
    """Callable/attribute sink standing in for figures, axes, images..."""

    def __call__(self, *a, **k):
        return _Anything()

    def __getattr__(self, name):
        return _Anything()

    def __iter__(self):
        return iter((_Anything(), _Anything()))


def _stub(name):
    mod = types.ModuleType(name)
    mod.__getattr__ = lambda attr: _Anything()
    return mod


def install_matplotlib_stub():
    try:
        import matplotlib  # noqa: F401
        return False
    except ImportError:
        pass
    top = _stub("matplotlib")
    top.pyplot = _stub("matplotlib.pyplot")
    top.animation = _stub("matplotlib.animation")
    sys.modules["matplotlib"] = top
    sys.modules["matplotlib.pyplot"] = top.pyplot
    sys.modules["matplotlib.animation"] = top.animation
    return True


def load():
    """Return (solver_module, analyser_module) of the reference."""
    if not available():
        raise RuntimeError("reference not present at %s" % REFERENCE_DIR)
    install_matplotlib_stub()
    out = []
    for name in ("solver", "analyser"):
        spec = importlib.util.spec_from_file_location(
            "_fdtd_reference_" + name, os.path.join(REFERENCE_DIR, name + ".py"))
        mod = importlib.util.module_from_spec(spec)
        # analyser.py does not import solver, and solver.py imports neither
        spec.loader.exec_module(mod)
        out.append(mod)
    return tuple(out)
