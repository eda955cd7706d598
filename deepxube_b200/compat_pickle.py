"""Reading problem-instance / results pickles written by the reference.

The reference pickles its own classes by module path (`deepxube.domains.cube3.Cube3State`,
`domains.grid_tutorial.GridState`, ..., deepxube/_cli.py:456-462, _solve.py:215).  The classes of this package
have the same attribute layout, so an Unpickler that redirects those module paths is enough to read the
reference's `--file` inputs and `results.pkl` outputs without the reference installed."""
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
