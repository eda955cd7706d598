"""Deterministic pseudo-heuristics used to drive reference, oracle and engine with IDENTICAL h-values
(the "injected-h" parity protocol).  TEST INFRASTRUCTURE ONLY.

Values are multiples of 1/64 in [0, 64) so that float32/float64 representations are exact and f-ties
are frequent (which stresses the (f, push order) tie-break of OPEN).
"""
import numpy as np


def _acc(states: np.ndarray) -> np.ndarray:
    s = np.ascontiguousarray(states, dtype=np.uint8).astype(np.uint64)
    w = ((np.arange(s.shape[1], dtype=np.uint64) + np.uint64(1)) * np.uint64(2654435761)) & np.uint64(0xFFFFFFFF)
    return (s * w[None, :]).sum(axis=1) & np.uint64(0xFFFFFFFF)


def pseudo_v(states: np.ndarray, goals: np.ndarray = None) -> np.ndarray:
    """[N,S] u8 -> float64 [N]"""
    acc = _acc(states)
    return (((acc >> np.uint64(20)) & np.uint64(0xFFF)).astype(np.float32) / np.float32(64.0)).astype(np.float64)


def pseudo_v_fine(states: np.ndarray, goals: np.ndarray = None) -> np.ndarray:
    """[N,S] u8 -> float64 [N]: 24-bit values (multiples of 2^-18 in [0, 64), exact in float32) from an FNV-1a hash of the
    row -- non-linear, so different states practically never tie (the linear _acc collides on states related by commuting
    moves).  For selections whose choice among exact ties is implementation-defined (beam search's np.argpartition)."""
    s = np.ascontiguousarray(states, dtype=np.uint8).astype(np.uint64)
    h = np.full(s.shape[0], 2166136261, dtype=np.uint64)
    for i in range(s.shape[1]):
        h = ((h ^ s[:, i]) * np.uint64(16777619)) & np.uint64(0xFFFFFFFF)
    h ^= h >> np.uint64(15)
    h = (h * np.uint64(2246822519)) & np.uint64(0xFFFFFFFF)
    h ^= h >> np.uint64(13)
    return ((h >> np.uint64(8)).astype(np.float64) / 262144.0)


def make_pseudo_q(num_actions: int):
    def pseudo_q(states: np.ndarray, goals: np.ndarray = None) -> np.ndarray:
        """[N,S] u8 -> float64 [N,A]"""
        acc = _acc(states)
        a = (np.arange(num_actions, dtype=np.uint64) * np.uint64(0x9E3779B1))[None, :]
        v = (acc[:, None] + a) & np.uint64(0xFFFFFFFF)
        return (((v >> np.uint64(20)) & np.uint64(0xFFF)).astype(np.float32) / np.float32(64.0)).astype(np.float64)
    return pseudo_q


def make_pseudo_q_fine(num_actions: int):
    """Tie-free counterpart of make_pseudo_q: q[state, a] = 24-bit FNV hash of (state bytes, a)."""
    def pseudo_q_fine(states: np.ndarray, goals: np.ndarray = None) -> np.ndarray:
        s = np.ascontiguousarray(states, dtype=np.uint8)
        out = np.empty((s.shape[0], num_actions), np.float64)
        for a in range(num_actions):
            out[:, a] = pseudo_v_fine(np.concatenate([s, np.full((s.shape[0], 1), a, np.uint8)], axis=1))
        return out
    return pseudo_q_fine
