"""SNK_OBS_BITS (include/snake_b200.h) on the CPU: an encoder written from the header's description, fed with the oracle's
boards over a rollout with wall deaths and resets, must round-trip through the product's decoder (unpack_bits) — the host-side
half of the format; the device-side half (the kernel's records decode to the oracle's step) is tests/test_env_parity_gpu.py."""
import numpy as np
import torch

from oracle import oracle_lib as O
from tests.util import pkg, synth_actions


def encode_bits(state_i8, reward, done, mask, action):
    """state_i8 (N, 2, 10, 10) as [n][frame][c][r]; returns (N, 24) uint8 records"""
    n = state_i8.shape[0]
    rec = np.zeros((n, 24), np.uint8)
    for f in range(2):
        b = state_i8[:, f]                                            # [n][c][r]
        inner = (b[:, 1:9, 1:9] == 1).astype(np.uint8)                # [n][c-1][r-1]
        rec[:, 8 * f:8 * f + 8] = (inner << np.arange(8, dtype=np.uint8)).sum(axis=2).astype(np.uint8)   # byte c-1, bit r-1
        for i in range(n):
            cs, rs = np.nonzero(b[i] == 2)
            if len(cs):
                rec[i, 16 + f] = rs[0] | (cs[0] << 4)
    for i in range(n):
        b = state_i8[i, 1]
        ring = np.zeros((10, 10), bool)
        ring[0, :] = ring[9, :] = ring[:, 0] = ring[:, 9] = True
        cs, rs = np.nonzero((b == 1) & ring)                          # a wall death: the head sits on the wall ring
        if not len(cs):
            cs, rs = np.nonzero(b == 1)
        rec[i, 18] = rs[0] | (cs[0] << 4)
    rec[:, 19] = mask[:, 0] | (mask[:, 1] << 1) | (mask[:, 2] << 2) | (done.astype(np.uint8) << 3) | (action.astype(np.uint8) << 4)
    rec[:, 20:24] = reward.astype(np.float32).view(np.uint8).reshape(n, 4)
    return rec


def test_unpack_bits_round_trips_the_oracle_states():
    S = pkg()
    n = 300
    ora = O.OracleBatch(n, auto_reset=True)
    wall_deaths = 0
    for t in range(250):
        act = synth_actions(n, t, seed=9)
        ref = ora.step(act, obs=("i8",))
        st = ref["obs_i8"].reshape(n, 2, 10, 10)
        rec = encode_bits(st, ref["reward"], ref["done"], ref["mask"], act)
        d = S.unpack_bits(torch.from_numpy(rec))
        assert np.array_equal(d["state"].numpy(), st), t
        assert np.array_equal(d["reward"].numpy().view(np.uint32), ref["reward"].astype(np.float32).view(np.uint32)), t
        assert np.array_equal(d["done"].numpy(), ref["done"]) and np.array_equal(d["mask"].numpy(), ref["mask"]), t
        assert np.array_equal(d["action"].numpy(), act), t
        ringv = np.concatenate([st[:, 1, 0, :], st[:, 1, 9, :], st[:, 1, :, 0], st[:, 1, :, 9]], axis=1)
        wall_deaths += int((ringv == 1).any(axis=1).sum())
    assert wall_deaths > 0                                            # the wall-overwrite case was exercised
