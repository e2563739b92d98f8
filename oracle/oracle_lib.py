"""ctypes binding of oracle/libsnake_oracle.so (TEST INFRASTRUCTURE — see snake_oracle.c).

Importable only from tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libsnake_oracle.so")

DEFAULT_FOOD_RC = [  # G1; pinned in tests/test_oracle_golden.py against oracle/xoshiro_food.py and the BSON
    (7, 5), (5, 7), (7, 3), (6, 7), (5, 4), (7, 7), (4, 4), (6, 2), (4, 3), (5, 5),
    (2, 6), (3, 6), (5, 4), (5, 8), (4, 6), (4, 3), (7, 4), (2, 2), (7, 5), (7, 3),
    (6, 5), (8, 3), (4, 9), (4, 7), (8, 6), (4, 4), (6, 6), (4, 2), (2, 8), (9, 3),
    (7, 4), (4, 8), (7, 7), (4, 2), (3, 9), (4, 8), (7, 8), (2, 7), (2, 6), (9, 5),
    (9, 9), (7, 5), (8, 6), (4, 2), (7, 6), (4, 6), (8, 5), (2, 5), (2, 8), (9, 4),
]


def build(force=False):
    src = os.path.join(_HERE, "snake_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "libsnake_oracle.so"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        L = _lib
        L.or_batch_create.restype = C.c_void_p
        L.or_batch_create.argtypes = [C.c_int64, C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.or_batch_destroy.argtypes = [C.c_void_p]
        L.or_batch_reset.argtypes = [C.c_void_p]
        L.or_batch_step.argtypes = [C.c_void_p, C.c_void_p, C.c_int] + [C.c_void_p] * 9
        L.or_batch_state.argtypes = [C.c_void_p] * 4
        L.or_batch_losing_mask.argtypes = [C.c_void_p] * 3
        L.or_batch_available_actions.argtypes = [C.c_void_p] * 2
        L.or_batch_select.argtypes = [C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]
        L.or_batch_scalars.argtypes = [C.c_void_p] * 9
        L.or_batch_run_random.restype = C.c_uint64
        L.or_batch_run_random.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_uint64]
        L.or_masked_target.argtypes = [C.c_void_p] * 4 + [C.c_double, C.c_float, C.c_void_p, C.c_int64]
        L.or_center_columns.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]
        L.or_epsilon_greedy_idx.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_int]
        L.or_game_new.restype = C.c_void_p
        L.or_game_new.argtypes = [C.c_void_p, C.c_int, C.c_int]
        for name in ("or_game_delete", "or_game_assemble_state"):
            getattr(L, name).argtypes = [C.c_void_p]
        L.or_game_step.argtypes = [C.c_void_p, C.c_int]
        L.or_game_get_board.argtypes = [C.c_void_p, C.c_void_p]
        L.or_game_get_state.argtypes = [C.c_void_p, C.c_void_p]
        L.or_game_next_state.argtypes = [C.c_void_p, C.c_void_p]
        L.or_game_virtual_step.argtypes = [C.c_void_p] * 4
        L.or_available_actions.argtypes = [C.c_void_p, C.c_void_p]
        L.or_game_score.restype = C.c_int64
        L.or_game_score.argtypes = [C.c_void_p]
        L.or_game_n_hist.restype = C.c_int64
        L.or_game_n_hist.argtypes = [C.c_void_p]
        L.or_game_reward.restype = C.c_float
        L.or_game_reward.argtypes = [C.c_void_p]
        L.or_game_error.restype = C.c_uint32
        L.or_game_error.argtypes = [C.c_void_p]
        for name in ("or_game_lost", "or_game_n_food"):
            getattr(L, name).argtypes = [C.c_void_p]
        L.or_scratch_size.restype = C.c_size_t
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def food_bytes(food_rc):
    return np.asarray(food_rc, dtype=np.uint8).reshape(-1, 2).copy()


class OracleGame:
    """One reference game (SnakeGame) — used by the golden-trajectory tests."""

    def __init__(self, food_rc=None, keep_history=True):
        f = food_bytes(DEFAULT_FOOD_RC if food_rc is None else food_rc)
        self._L = lib()
        self._g = self._L.or_game_new(_p(f), len(f), int(keep_history))
        self._sc = np.zeros(self._L.or_scratch_size(), dtype=np.uint8)

    def __del__(self):
        if getattr(self, "_g", None):
            self._L.or_game_delete(self._g)
            self._g = None

    def available_actions(self):
        out = np.zeros(3, np.uint8)
        self._L.or_available_actions(self._g, _p(out))
        return out

    def step(self, d):
        self._L.or_game_step(self._g, int(d))

    def virtual_step(self):
        av = np.zeros(3, np.uint8)
        lost = np.zeros(3, np.uint8)
        self._L.or_game_virtual_step(self._g, _p(self._sc), _p(av), _p(lost))
        return av, lost

    @property
    def board(self):
        """[row][col] int64 (0-based indexing of Julia's 1-based board)"""
        out = np.zeros(100, np.int64)
        self._L.or_game_get_board(self._g, _p(out))
        return out.reshape(10, 10).T.copy()

    def next_state(self):
        out = np.zeros(200, np.int64)
        self._L.or_game_next_state(self._g, _p(out))
        return out

    def assemble_state(self):
        self._L.or_game_assemble_state(self._g)
        out = np.zeros(200, np.int64)
        self._L.or_game_get_state(self._g, _p(out))
        return out

    score = property(lambda s: s._L.or_game_score(s._g))
    lost = property(lambda s: bool(s._L.or_game_lost(s._g)))
    reward = property(lambda s: s._L.or_game_reward(s._g))
    error = property(lambda s: s._L.or_game_error(s._g))
    n_food = property(lambda s: s._L.or_game_n_food(s._g))
    n_hist = property(lambda s: s._L.or_game_n_hist(s._g))


class OracleBatch:
    """N independent reference games behind the same call shapes as the C-ABI."""

    def __init__(self, n, food_rc=None, auto_reset=True, keep_history=False):
        f = food_bytes(DEFAULT_FOOD_RC if food_rc is None else food_rc)
        self.n = int(n)
        self._L = lib()
        self._b = self._L.or_batch_create(self.n, _p(f), len(f), int(auto_reset), int(keep_history))

    def __del__(self):
        if getattr(self, "_b", None):
            self._L.or_batch_destroy(self._b)
            self._b = None

    def reset(self):
        self._L.or_batch_reset(self._b)

    def step(self, action, is_abs=False, obs=("f32",), want_mask=True):
        n = self.n
        action = np.ascontiguousarray(action, dtype=np.uint8)
        out = {
            "reward": np.zeros(n, np.float32), "done": np.zeros(n, np.uint8),
            "mask": np.zeros((n, 3), np.uint8), "ep_return": np.zeros(n, np.float32),
            "ep_score": np.zeros(n, np.int32), "av_next": np.zeros((n, 3), np.uint8),
        }
        o32 = np.zeros((n, 200), np.float32) if "f32" in obs else None
        o8 = np.zeros((n, 200), np.int8) if "i8" in obs else None
        o64 = np.zeros((n, 200), np.int64) if "i64" in obs else None
        self._L.or_batch_step(self._b, _p(action), int(is_abs), _p(out["reward"]), _p(out["done"]),
                              _p(o32), _p(o8), _p(o64), _p(out["mask"]), _p(out["ep_return"]),
                              _p(out["ep_score"]), _p(out["av_next"]))
        out["obs_f32"], out["obs_i8"], out["obs_i64"] = o32, o8, o64
        return out

    def state(self, kind="f32"):
        n = self.n
        dt = {"f32": np.float32, "i8": np.int8, "i64": np.int64}[kind]
        o = np.zeros((n, 200), dt)
        args = [None, None, None]
        args[["f32", "i8", "i64"].index(kind)] = _p(o)
        self._L.or_batch_state(self._b, *args)
        return o

    def losing_mask(self):
        m = np.zeros((self.n, 3), np.uint8)
        av = np.zeros((self.n, 3), np.uint8)
        self._L.or_batch_losing_mask(self._b, _p(m), _p(av))
        return m, av

    def available_actions(self):
        o = np.zeros((self.n, 3), np.uint8)
        self._L.or_batch_available_actions(self._b, _p(o))
        return o

    def select(self, q, eps, u, ridx):
        q = np.ascontiguousarray(q, np.float32)
        u = np.ascontiguousarray(u, np.float32)
        ridx = np.ascontiguousarray(ridx, np.uint8)
        o = np.zeros(self.n, np.uint8)
        self._L.or_batch_select(self._b, _p(q), float(eps), _p(u), _p(ridx), _p(o))
        return o

    def scalars(self):
        n = self.n
        d = {"score": np.zeros(n, np.int32), "lost": np.zeros(n, np.uint8), "error": np.zeros(n, np.uint32),
             "snake_len": np.zeros(n, np.int32), "head_rc": np.zeros((n, 2), np.uint8),
             "food_rc": np.zeros((n, 2), np.uint8), "n_food_left": np.zeros(n, np.int32),
             "n_hist": np.zeros(n, np.int32)}
        self._L.or_batch_scalars(self._b, *[_p(d[k]) for k in
                                            ("score", "lost", "error", "snake_len", "head_rc", "food_rc",
                                             "n_food_left", "n_hist")])
        return d

    def run_random(self, lo, hi, n_steps, seed=42):
        return self._L.or_batch_run_random(self._b, int(lo), int(hi), int(n_steps), int(seed))


def masked_target(q_next, mask, r, done, gamma=0.97, fill=-100.0):
    q_next = np.ascontiguousarray(q_next, np.float32)
    mask = np.ascontiguousarray(mask, np.uint8)
    r = np.ascontiguousarray(r, np.float32)
    done = np.ascontiguousarray(done, np.uint8)
    B = r.shape[0]
    y = np.zeros(B, np.float64)
    lib().or_masked_target(_p(q_next), _p(mask), _p(r), _p(done), float(gamma), float(fill), _p(y), B)
    return y


def center_columns(D_colmajor, P, K):
    """D_colmajor: flat float64 array of P*K (column k at [k*P:(k+1)*P]); centred in place."""
    mean = np.zeros(P, np.float64)
    var = np.zeros(P, np.float64)
    lib().or_center_columns(_p(D_colmajor), P, K, _p(mean), _p(var))
    return mean, var


def epsilon_greedy_idx(q3, eps, u, ridx):
    q3 = np.ascontiguousarray(q3, np.float32)
    return lib().or_epsilon_greedy_idx(_p(q3), float(eps), float(u), int(ridx))
