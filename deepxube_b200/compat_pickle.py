"""Reading AND writing problem-instance / results pickles in the reference's format.

The reference pickles its own classes by module path (`deepxube.domains.cube3.Cube3State`,
`domains.grid_tutorial.GridState`, ..., deepxube/_cli.py:456-462, _solve.py:215).  The classes of this package
have the same attribute layout, so an Unpickler that redirects those module paths is enough to read the
reference's `--file` inputs and `results.pkl` outputs without the reference installed.  `dump(..., ref_classes=True)` writes
the other direction: a pickle whose class references are the reference's module paths, so that the reference's own tooling
(`deepxube viz --soln`, `deepxube solve --file`) opens files produced here (`--ref_pickle` of `solve` / `problem_inst`)."""
import io
import pickle
from typing import Any

_MODULE_MAP = {
    "deepxube.domains.cube3": "deepxube_b200.domains.cube3",
    "deepxube.domains.npuzzle": "deepxube_b200.domains.npuzzle",
    "deepxube.domains.grid": "deepxube_b200.domains.grid",
    "deepxube.domains.lightsout": "deepxube_b200.domains.lightsout",
    "domains.grid_tutorial": "deepxube_b200.domains.grid",      # tutorial copy of the grid domain, same classes
}


class _CompatUnpickler(pickle.Unpickler):
    def find_class(self, module: str, name: str) -> Any:
        return super().find_class(_MODULE_MAP.get(module, module), name)


def load(file: Any) -> Any:
    if isinstance(file, (str, bytes)):
        with open(file, "rb") as f:
            return _CompatUnpickler(f).load()
    return _CompatUnpickler(file).load()


def loads(data: bytes) -> Any:
    return _CompatUnpickler(io.BytesIO(data)).load()


def dumps(obj: Any, ref_classes: bool = False) -> bytes:
    """ref_classes: class references are written with the reference's module paths.  Protocol 2 is used for that: its GLOBAL
    opcode stores `module\nname\n` as plain text lines without a length prefix, so the module path can be rewritten in the
    byte stream (the classes keep their names and attribute layout; tutorial pickles of `domains.grid_tutorial` are written
    as `deepxube.domains.grid`, which the reference registers itself)."""
    if not ref_classes:
        return pickle.dumps(obj, protocol=-1)
    data = pickle.dumps(obj, protocol=2)
    for ref_mod, mod in _MODULE_MAP.items():
        if ref_mod.startswith("deepxube."):
            data = data.replace(b"c" + mod.encode() + b"\n", b"c" + ref_mod.encode() + b"\n")
    return data


def dump(obj: Any, file: Any, ref_classes: bool = False) -> None:
    data = dumps(obj, ref_classes)
    if isinstance(file, (str, bytes)):
        with open(file, "wb") as f:
            f.write(data)
    else:
        file.write(data)
