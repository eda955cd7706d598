"""Move tables of the in-scope domains (host copies, used for problem-instance generation and tests; the
kernels build the same tables in deepxube_b200/csrc/engine.cu::make_domain_host).

cube3: next[dst] = cur[src] for the 20 stickers each quarter turn moves; action order U-1 U1 D-1 D1 L-1 L1
R-1 R1 B-1 B1 F-1 F1 (deepxube/domains/cube3.py:492, :583-671).  Stored as cycles: a quarter turn is five
4-cycles (two on the face, three around it); (a b c d) means next[b] = cur[a], next[c] = cur[b], ...
Pinned against the reference by tests/golden/domains.npz.
"""
from functools import lru_cache

import numpy as np

# one clockwise-numbered set of 4-cycles per face turn "X-1"; "X1" is the inverse permutation
_CUBE3_CYCLES_NEG = {
    "U": [(2, 0, 6, 8), (5, 1, 3, 7), (38, 20, 47, 29), (41, 23, 50, 32), (44, 26, 53, 35)],
    "D": [(11, 9, 15, 17), (14, 10, 12, 16), (45, 18, 36, 27), (48, 21, 39, 30), (51, 24, 42, 33)],
    "L": [(20, 18, 24, 26), (23, 19, 21, 25), (45, 0, 44, 9), (46, 1, 43, 10), (47, 2, 42, 11)],
    "R": [(29, 27, 33, 35), (32, 28, 30, 34), (38, 6, 51, 15), (37, 7, 52, 16), (36, 8, 53, 17)],
    "B": [(38, 36, 42, 44), (41, 37, 39, 43), (18, 2, 35, 15), (19, 5, 34, 12), (20, 8, 33, 9)],
    "F": [(47, 45, 51, 53), (50, 46, 48, 52), (29, 0, 24, 17), (28, 3, 25, 14), (27, 6, 26, 11)],
}
CUBE3_ACTION_NAMES = [f"{f}{n}" for f in "UDLRBF" for n in (-1, 1)]


@lru_cache(maxsize=None)
def cube3_perm_table() -> np.ndarray:
    """[12, 54] int64, next[:, d] = cur[:, perm[a, d]]."""
    perm = np.tile(np.arange(54, dtype=np.int64), (12, 1))
    for fi, face in enumerate("UDLRBF"):
        neg = np.arange(54, dtype=np.int64)
        for cyc in _CUBE3_CYCLES_NEG[face]:
            for k in range(4):
                neg[cyc[(k + 1) % 4]] = cyc[k]            # next[b] = cur[a]
        perm[2 * fi] = neg
        inv = np.empty(54, dtype=np.int64)
        inv[neg] = np.arange(54)
        perm[2 * fi + 1] = inv
    perm.setflags(write=False)
    return perm


@lru_cache(maxsize=None)
def npuzzle_swap_table(dim: int) -> np.ndarray:
    """[d*d, 4] int64: index the blank swaps with for U, D, L, R (row+1, row-1, col+1, col-1); a move off the
    board leaves the blank where it is and still costs 1.0 (deepxube/domains/npuzzle.py:264-321)."""
    z = np.arange(dim * dim)
    i, j = z // dim, z % dim
    lut = np.stack([np.where(i < dim - 1, z + dim, z), np.where(i > 0, z - dim, z),
                    np.where(j < dim - 1, z + 1, z), np.where(j > 0, z - 1, z)], axis=1).astype(np.int64)
    lut.setflags(write=False)
    return lut


@lru_cache(maxsize=None)
def lightsout_move_table(dim: int) -> np.ndarray:
    """[d*d, 5] int64: the cells action a flips = [a, a+d, a-d, a+1, a-1], a neighbour that would leave the board replaced
    by a itself (deepxube/domains/lightsout.py:69-78)."""
    a = np.arange(dim * dim)
    x, y = a // dim, a % dim
    mm = np.stack([a, np.where(x < dim - 1, a + dim, a), np.where(x > 0, a - dim, a),
                   np.where(y < dim - 1, a + 1, a), np.where(y > 0, a - 1, a)], axis=1).astype(np.int64)
    mm.setflags(write=False)
    return mm
