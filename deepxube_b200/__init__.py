"""deepxube_b200 -- B200-native batched heuristic search (batch weighted A* / Q*) behind DeepXube's
registries and call protocols.  Importing the package registers the domains (`cube3`, `npuzzle`, `grid`),
the network (`resnet_fc`), the function kinds (`heurv`, `heurq_fixout`) and the pathfinders (`graph_v`,
`graph_q`), like `import deepxube` does for the reference (deepxube/__init__.py:1-24).

The compute path is libdxb200.so (hand-written sm_100a CUDA, include/dxb200.h); there is no CPU fallback.
"""
__version__ = "0.1.0"

from . import domains  # noqa: F401,E402
from .domains import cube3, grid, lightsout, npuzzle  # noqa: F401,E402
from .nnets import resnet_fc  # noqa: F401,E402
from .pathfind_fns import fns_core  # noqa: F401,E402
from .pathfinding import graph_search, beam_search  # noqa: F401,E402
