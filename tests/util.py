"""Shared helpers for the parity tests."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import __graft_entry__ as graft  # noqa: E402


def pkg():
    return graft.load_package()


def bits(a):
    a = np.ascontiguousarray(a)
    return a.view({4: np.uint32, 8: np.uint64}[a.dtype.itemsize])


def splitmix64(x):
    x = (np.uint64(x) + np.uint64(0x9E3779B97F4A7C15))
    x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return x ^ (x >> np.uint64(31))


def synth_actions(n, t, seed=42):
    """SURVEY §8(d) synthetic draw: per-env per-step counter hash -> action idx in 0..2"""
    with np.errstate(over="ignore"):
        env = np.arange(n, dtype=np.uint64)
        x = splitmix64(np.uint64(seed) ^ splitmix64(env * np.uint64(0x100000001B3) + np.uint64(t)))
    return (x % np.uint64(3)).astype(np.uint8)


def unpack2(packed):
    """(N,50) uint8 2-bit packed -> (N,200) int8 board values"""
    p = np.asarray(packed, dtype=np.uint8)
    codes = np.stack([(p >> (2 * j)) & 3 for j in range(4)], axis=-1).reshape(p.shape[0], -1)
    return np.where(codes == 3, -1, codes).astype(np.int8)


# The reference's best game (README.md:54-58, trainer_gifs/very_long_double_training3.gif): 237 absolute moves, 33 apples,
# death on the bottom wall.  Recovered frame by frame in tests/test_oracle_golden.py::test_g2_score33_trajectory_replays_exactly.
G2_ACTIONS = (
    "UURUURRDDDRURULLLLDDRRRRULULLULLDDRRDRRRULULLULURURRDDRRDDDLULDLLUULLURULURRRDRDDRDDLDLLLURRULULUULURRRRD"
    "RRURDDLLDRDDLDLLLLLUUUURRDDRRUUULLLLURRRRRDRDLDRDRDLLLDLDLLUULUUURRDRUULLURRRRRRDLDLDRDRDLLDLLLULLURRRULLUL"
    "URRRURDRRDLDLDRRRDLLDLDLD")
