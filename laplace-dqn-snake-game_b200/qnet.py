"""The reference's Q-network (structs.jl:127-139) on the device — native kernels only.

    Conv((3,3), 2=>16, relu; pad=1) -> Conv((3,3), 16=>32, relu; pad=1) -> Conv((6,6), 32=>64, relu) ->
    Flux.flatten -> Dense(1600, 64, relu) -> Dense(64, 3)

`QNet` hands Flux.destructure(q_net) to snk_qnet_create and runs snk_qnet_forward (tcgen05 implicit-GEMM convolutions,
csrc/qnet.cu).  precision="f32" (default) is the Float32-faithful mode the reference's Float32 network needs for
identical greedy actions; precision="bf16" is the fast, lower-precision mode.  There is no library (cuDNN) path in the
product: the torch restatement used as a timing baseline and cross-check lives in tools/torch_qnet.py.
"""
import ctypes as C
import math

import numpy as np
import torch

from . import bson_io

SHAPES = [("conv", (3, 3, 2, 16), 1), ("conv", (3, 3, 16, 32), 1), ("conv", (6, 6, 32, 64), 0),
          ("dense", (64, 1600), None), ("dense", (3, 64), None)]
N_PARAMS = 181395
PRECISIONS = {"bf16": 0, "f32": 1}            # SNK_QNET_BF16, SNK_QNET_F32 (include/snake_b200.h)
PRECISION_NOTES = {
    "f32": "Float32-faithful: fp16 (hi, lo) split operands, four partial products on tcgen05, FP32 accumulation",
    "bf16": "bf16 operands on tcgen05, FP32 accumulation: ~1.5e-2 of max|Q|, NOT the reference's Float32 fidelity",
}


def glorot_layers(seed=0, in_frames=2):
    """Flux's default init: Glorot-uniform weights, zero bias (seeded, synthetic: the two-frame checkpoints named
    by BASELINE config 4 are missing from the reference mount)."""
    rng = np.random.default_rng(seed)
    layers = []
    for kind, shp, pad in SHAPES:
        if kind == "conv":
            k1, k2, cin, cout = shp
            if cin == 2:
                cin = in_frames
            fan_in, fan_out = k1 * k2 * cin, k1 * k2 * cout
            lim = math.sqrt(6.0 / (fan_in + fan_out))
            layers.append(("conv", {"W": rng.uniform(-lim, lim, (k1, k2, cin, cout)).astype(np.float32),
                                    "b": np.zeros(cout, np.float32), "pad": [pad, pad], "stride": [1, 1]}))
            if shp[0] == 6:
                layers.append(("flatten", {}))
        else:
            out, inn = shp
            lim = math.sqrt(6.0 / (inn + out))
            layers.append(("dense", {"W": rng.uniform(-lim, lim, (out, inn)).astype(np.float32),
                                     "b": np.zeros(out, np.float32)}))
    return layers


class QNet:
    """q_net / t_net of DQNModel (structs.jl:120-147) as a callable: obs (N,2,10,10) f32 [= Julia (10,10,2,N)] -> Q (N,3)."""

    def __init__(self, layers, device, precision="f32"):
        from . import _check, lib
        if precision not in PRECISIONS:
            raise ValueError("precision must be 'f32' or 'bf16'")
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise ValueError("QNet runs on a CUDA device only (no CPU fallback)")
        self.precision = precision
        self.theta = np.ascontiguousarray(bson_io.destructure(layers), dtype=np.float32)
        if self.theta.size != N_PARAMS:
            raise ValueError("the native Q-net is the two-frame network of structs.jl:127-139 (%d parameters), got %d"
                             % (N_PARAMS, self.theta.size))
        self.n_params = int(self.theta.size)
        self._q = C.c_void_p()
        _check(lib().snk_qnet_create(C.byref(self._q), self.theta.ctypes.data_as(C.c_void_p), self.theta.size,
                                     self.device.index or 0, PRECISIONS[precision]))

    def close(self):
        if getattr(self, "_q", None):
            try:
                from . import lib
                lib().snk_qnet_destroy(self._q)
            except Exception:          # interpreter shutdown: module globals may already be gone
                pass
            self._q = None

    __del__ = close

    @classmethod
    def from_trainer_bson(cls, path, device, which="q_net", precision="f32"):
        """q_net / t_net of ./trainers/<name>.bson (load_trainer, utils.jl:413-418)"""
        q, t = bson_io.load_trainer_nets(path)
        return cls(q if which == "q_net" else t, device, precision)

    def forward(self, obs, out=None):
        """snk_qnet_forward: obs (N,2,10,10) float32 contiguous -> Q (N,3) float32 [= Julia (3,N)]"""
        from . import _check, _ptr, lib
        n = obs.shape[0]
        if out is None:
            out = torch.empty(n, 3, dtype=torch.float32, device=self.device)
        st = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        _check(lib().snk_qnet_forward(self._q, _ptr(obs, torch.float32, n * 200, self.device), n,
                                      _ptr(out, torch.float32, 3 * n, self.device), st))
        return out

    __call__ = forward

    def sample_grads(self, states, actions, targets, planes=None, want_J=False, want_loss=True):
        """Per-sample gradients of huber(q_net(s_i)[a_i], y_i) (utils.jl:452-466), FP32, Flux.destructure order.
        states (B,2,10,10) f32, actions (B) u8 (0-based), targets (B) f64 — as ReplayBuffer.stack_exp / masked_target give.
        planes: (hi_ptr, lo_ptr, pitch) of bf16 Gram planes to fill (GramShard.planes(), GramPlan.planes()) or None.
        Returns dict(J (B, 181395) f32 if want_J, loss (B) f32 if want_loss)."""
        from . import _check, _ptr, lib
        B = states.shape[0]
        dev = self.device
        out = {}
        if want_J:
            out["J"] = torch.empty(B, N_PARAMS, dtype=torch.float32, device=dev)
        if want_loss:
            out["loss"] = torch.empty(B, dtype=torch.float32, device=dev)
        hi, lo, pitch = (C.c_void_p(planes[0]), C.c_void_p(planes[1]), int(planes[2])) if planes is not None else (None, None, 0)
        st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        _check(lib().snk_qnet_sample_grads(self._q, _ptr(states, torch.float32, B * 200, dev), _ptr(actions, torch.uint8, B, dev),
                                           _ptr(targets, torch.float64, B, dev), B, hi, lo, pitch, _ptr(out.get("J")), N_PARAMS,
                                           _ptr(out.get("loss")), st))
        return out

    def overflow(self):
        """f32 mode: True if an activation left the fp16 range of the split operands since the last call (synchronises)."""
        from . import _check, lib
        f = C.c_int(0)
        _check(lib().snk_qnet_overflow_host(self._q, C.byref(f)))
        return bool(f.value)
