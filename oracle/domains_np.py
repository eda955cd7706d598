"""Oracle restatement of the three in-scope domains as flat uint8 arrays (numpy).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

State rows:
  cube3     54 x u8 sticker colours, goal = arange(54)//9          (ref: deepxube/domains/cube3.py:502-503)
  npuzzle.d d*d x u8 tiles, 0 = blank, goal = [1..d*d-1, 0]         (ref: deepxube/domains/npuzzle.py:63-73)
  grid.d    2 x u8 (robot_x, robot_y); goal row 2 x u8             (ref: deepxube/domains/grid.py:24-35)
  lightsout.d  d*d x u8 lights (0 / 1), goal = all off              (ref: deepxube/domains/lightsout.py:16-24, :79)

Children are produced state-major / action-minor, child index = i*A + a
(ref: deepxube/base/domain.py:288-312).  Every transition costs 1.0.
"""
from typing import Tuple

import numpy as np

# --------------------------------------------------------------------------------------------
# cube3: the 20 moved stickers of each of the 12 moves, "dst<src" (ref: cube3.py:583-671; the
# reference builds 24 (new,old) pairs per move with the 4 face corners listed twice -- this is the
# de-duplicated permutation it implies, dumped in SURVEY.md App. A.1 and pinned by
# tests/golden/domains.npz which is generated from the reference itself).
# Action order (ref: cube3.py:492): U-1 U1 D-1 D1 L-1 L1 R-1 R1 B-1 B1 F-1 F1
# --------------------------------------------------------------------------------------------
_CUBE3_MOVES = [
    "0<2 1<5 2<8 3<1 5<7 6<0 7<3 8<6 20<38 23<41 26<44 29<47 32<50 35<53 38<29 41<32 44<35 47<20 50<23 53<26",
    "0<6 1<3 2<0 3<7 5<1 6<8 7<5 8<2 20<47 23<50 26<53 29<38 32<41 35<44 38<20 41<23 44<26 47<29 50<32 53<35",
    "9<11 10<14 11<17 12<10 14<16 15<9 16<12 17<15 18<45 21<48 24<51 27<36 30<39 33<42 36<18 39<21 42<24 45<27 48<30 51<33",
    "9<15 10<12 11<9 12<16 14<10 15<17 16<14 17<11 18<36 21<39 24<42 27<45 30<48 33<51 36<27 39<30 42<33 45<18 48<21 51<24",
    "0<45 1<46 2<47 9<44 10<43 11<42 18<20 19<23 20<26 21<19 23<25 24<18 25<21 26<24 42<2 43<1 44<0 45<9 46<10 47<11",
    "0<44 1<43 2<42 9<45 10<46 11<47 18<24 19<21 20<18 21<25 23<19 24<26 25<23 26<20 42<11 43<10 44<9 45<0 46<1 47<2",
    "6<38 7<37 8<36 15<51 16<52 17<53 27<29 28<32 29<35 30<28 32<34 33<27 34<30 35<33 36<17 37<16 38<15 51<6 52<7 53<8",
    "6<51 7<52 8<53 15<38 16<37 17<36 27<33 28<30 29<27 30<34 32<28 33<35 34<32 35<29 36<8 37<7 38<6 51<15 52<16 53<17",
    "2<18 5<19 8<20 9<33 12<34 15<35 18<15 19<12 20<9 33<8 34<5 35<2 36<38 37<41 38<44 39<37 41<43 42<36 43<39 44<42",
    "2<35 5<34 8<33 9<20 12<19 15<18 18<2 19<5 20<8 33<9 34<12 35<15 36<42 37<39 38<36 39<43 41<37 42<44 43<41 44<38",
    "0<29 3<28 6<27 11<26 14<25 17<24 24<0 25<3 26<6 27<11 28<14 29<17 45<47 46<50 47<53 48<46 50<52 51<45 52<48 53<51",
    "0<24 3<25 6<26 11<27 14<28 17<29 24<17 25<14 26<11 27<6 28<3 29<0 45<51 46<48 47<45 48<52 50<46 51<53 52<50 53<47",
]


def cube3_perm() -> np.ndarray:
    """[12, 54] int64: next[:, d] = cur[:, perm[a, d]]."""
    perm = np.tile(np.arange(54, dtype=np.int64), (12, 1))
    for a, spec in enumerate(_CUBE3_MOVES):
        for pair in spec.split():
            d, s = pair.split("<")
            perm[a, int(d)] = int(s)
    return perm


def npuzzle_swap_lut(dim: int) -> np.ndarray:
    """[d*d, 4] int64: index the blank swaps with for U,D,L,R; blocked move -> blank's own index
    (ref: npuzzle.py:264-308: U -> row+1, D -> row-1, L -> col+1, R -> col-1)."""
    lut = np.zeros((dim * dim, 4), dtype=np.int64)
    for z in range(dim * dim):
        i, j = divmod(z, dim)
        lut[z, 0] = (i + 1) * dim + j if i < dim - 1 else z
        lut[z, 1] = (i - 1) * dim + j if i > 0 else z
        lut[z, 2] = i * dim + (j + 1) if j < dim - 1 else z
        lut[z, 3] = i * dim + (j - 1) if j > 0 else z
    return lut


def lightsout_moves(dim: int) -> np.ndarray:
    """[d*d, 5] int64: the cells action a toggles = [a, a+d, a-d, a+1, a-1], an out-of-board neighbour replaced by a
    itself (ref: lightsout.py:69-78).  The reference toggles through one fancy-index assignment
    `t[i, mm] = (t[i, mm] + 1) % 2` (:186), so a cell listed twice is still toggled ONCE."""
    mm = np.zeros((dim * dim, 5), dtype=np.int64)
    for a in range(dim * dim):
        x, y = a // dim, a % dim
        mm[a] = [a, a + dim if x < dim - 1 else a, a - dim if x > 0 else a, a + 1 if y < dim - 1 else a, a - 1 if y > 0 else a]
    return mm


class OracleDomain:
    """name in {"cube3", "npuzzle", "grid", "lightsout"}; dim is the puzzle / grid side (ignored for cube3)."""

    def __init__(self, name: str, dim: int = 0):
        self.name = name
        self.dim = dim
        if name == "cube3":
            self.state_len, self.num_actions = 54, 12
            self.perm = cube3_perm()
            self.goal_len = 54
        elif name == "npuzzle":
            assert 2 <= dim <= 15
            self.state_len, self.num_actions = dim * dim, 4
            self.lut = npuzzle_swap_lut(dim)
            self.goal_len = dim * dim
        elif name == "grid":
            assert 1 <= dim <= 255
            self.state_len, self.num_actions = 2, 4
            self.goal_len = 2
        elif name == "lightsout":
            assert 2 <= dim <= 15
            self.state_len, self.num_actions = dim * dim, dim * dim
            self.moves = lightsout_moves(dim)
            self.goal_len = dim * dim
        else:
            raise ValueError(name)

    # -- goals / inputs -----------------------------------------------------------------------
    def canonical_goal(self) -> np.ndarray:
        if self.name == "cube3":
            return (np.arange(54) // 9).astype(np.uint8)
        if self.name == "npuzzle":
            n = self.dim * self.dim
            return np.concatenate((np.arange(1, n), [0])).astype(np.uint8)
        if self.name == "lightsout":
            return np.zeros(self.dim * self.dim, np.uint8)
        raise ValueError("grid has no canonical goal")

    def input_info(self) -> Tuple[int, int]:
        """(number of integer inputs, one-hot depth) of the flat NN input
        (ref: cube3.py:528-529, npuzzle.py:160-161, grid.py:126-128)."""
        if self.name == "cube3":
            return 54, 6
        if self.name == "npuzzle":
            return self.dim * self.dim, self.dim * self.dim
        if self.name == "lightsout":
            return self.dim * self.dim, 1          # depth 1: the 0 / 1 value itself is the feature (pytorch_models.py:15-24)
        return 4, self.dim

    def nnet_input(self, states: np.ndarray, goals: np.ndarray) -> np.ndarray:
        """Integer NN input rows (ref: cube3.py:534-535, npuzzle.py:163-164, grid.py:130-133);
        the goal is part of the input only for grid."""
        if self.name == "grid":
            return np.concatenate([states, goals], axis=1).astype(np.int64)
        return states.astype(np.int64)

    # -- transitions ---------------------------------------------------------------------------
    def next_state(self, states: np.ndarray, actions: np.ndarray) -> np.ndarray:
        """states [N, S] u8, actions [N] int -> [N, S] u8 (ref: cube3.py:583-596, npuzzle.py:310-321,
        grid.py:74-86)."""
        states = np.ascontiguousarray(states, dtype=np.uint8)
        actions = np.asarray(actions, dtype=np.int64)
        n = states.shape[0]
        if self.name == "cube3":
            return np.take_along_axis(states, self.perm[actions], axis=1)
        if self.name == "npuzzle":
            z = np.argmax(states == 0, axis=1)
            sw = self.lut[z, actions]
            out = states.copy()
            rows = np.arange(n)
            out[rows, z] = states[rows, sw]
            out[rows, sw] = 0
            return out
        if self.name == "lightsout":
            out = states.copy()
            for k in range(n):
                cells = np.unique(self.moves[actions[k]])
                out[k, cells] ^= 1
            return out
        out = states.astype(np.int64)
        d = self.dim
        a = actions
        out[:, 0] = np.where(a == 0, np.maximum(out[:, 0] - 1, 0), np.where(a == 1, np.minimum(out[:, 0] + 1, d - 1), out[:, 0]))
        out[:, 1] = np.where(a == 2, np.maximum(out[:, 1] - 1, 0), np.where(a == 3, np.minimum(out[:, 1] + 1, d - 1), out[:, 1]))
        return out.astype(np.uint8)

    def expand(self, states: np.ndarray) -> np.ndarray:
        """[N, S] -> [N*A, S], child index i*A + a (ref: base/domain.py:288-312)."""
        n, a = states.shape[0], self.num_actions
        rep = np.repeat(states, a, axis=0)
        acts = np.tile(np.arange(a), n)
        return self.next_state(rep, acts)

    def is_solved(self, states: np.ndarray, goals: np.ndarray) -> np.ndarray:
        """(ref: cube3.py:514-517; npuzzle.py:154-158 with the goal == d*d wildcard; grid.py:68-69)."""
        if self.name == "npuzzle":
            return np.all((states == goals) | (goals == self.dim * self.dim), axis=1)
        return np.all(states == goals, axis=1)
