"""Import shim for the UNMODIFIED reference (forestagostinelli/deepxube) at /root/reference.

TEST INFRASTRUCTURE ONLY.  Used solely by the golden-vector generators under ``tools/`` (run in
the build container, where /root/reference exists) and by tests that are skipped when the
reference tree is absent.  Nothing in the product path imports this module.

The reference imports four GUI/ASP packages at module top that are not installed here
(matplotlib, clingo, wget, imageio -- deepxube/base/domain.py:8,15, domains/cube3.py:5-6,
domains/npuzzle.py:9-10, domains/grid.py:19-21, logic/asp.py:2-5).  None of them is touched by
the hot path, so they are replaced by empty stub modules before ``import deepxube``.
"""
import os
import sys
import types

REF_ROOT = os.environ.get("DXB200_REFERENCE", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "deepxube"))


class _Anything:
    """Stub class: any attribute is another stub, calling it returns a stub."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Anything()

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Anything()


def _stub(name: str, **attrs) -> types.ModuleType:
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)

    def _mod_getattr(attr):
        if attr.startswith("__"):
            raise AttributeError(attr)
        return _Anything

    mod.__getattr__ = _mod_getattr  # type: ignore[attr-defined]
    sys.modules[name] = mod
    return mod


def install_stubs() -> None:
    if "matplotlib" not in sys.modules:
        mpl = _stub("matplotlib")
        # cube3.py:227 subclasses plt.Axes at import time -> must be a real class
        plt = _stub("matplotlib.pyplot", Axes=type("Axes", (), {}))
        mpl.pyplot = plt
        for sub in ("figure", "axes", "colors", "patches", "widgets", "animation", "backend_bases"):
            setattr(mpl, sub, _stub("matplotlib." + sub))
        sys.modules["matplotlib.figure"].Figure = type("Figure", (), {})
        sys.modules["matplotlib.colors"].ListedColormap = _Anything
    if "clingo" not in sys.modules:
        cl = _stub("clingo")
        for n in ("Control", "Symbol", "Model", "SolveHandle"):
            setattr(cl, n, type(n, (), {}))
        cl.parse_term = lambda *a, **k: None
        cl.solving = _stub("clingo.solving", Model=cl.Model, SolveHandle=cl.SolveHandle)
    for n in ("wget", "imageio"):
        if n not in sys.modules:
            _stub(n)


def import_reference():
    """Returns the imported reference package ``deepxube`` (raises if the tree is absent)."""
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    install_stubs()
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import deepxube  # noqa: F401
    return deepxube
