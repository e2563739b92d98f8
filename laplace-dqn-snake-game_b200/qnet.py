"""The reference's Q-network (structs.jl:127-139) on the device.

    Conv((3,3), 2=>16, relu; pad=1) -> Conv((3,3), 16=>32, relu; pad=1) -> Conv((6,6), 32=>64, relu) ->
    Flux.flatten -> Dense(1600, 64, relu) -> Dense(64, 3)

`QNet` holds the parameters in torch layout (converted from Flux's with bson_io.conv_weight_to_torch) and runs
the forward pass.  backend="torch" is a LIBRARY path (cuDNN / cuBLAS through torch), kept as the reference
implementation the native kernel is checked against; it is labelled as such wherever it is timed.
"""
import ctypes as C
import math

import numpy as np
import torch
import torch.nn.functional as F

from . import bson_io

SHAPES = [("conv", (3, 3, 2, 16), 1), ("conv", (3, 3, 16, 32), 1), ("conv", (6, 6, 32, 64), 0),
          ("dense", (64, 1600), None), ("dense", (3, 64), None)]


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


BACKEND_NOTES = {
    "torch": "LIBRARY path: torch conv2d/linear (cuDNN/cuBLAS), fp32 — the baseline, not a kernel of this repo",
    "torch_bf16": "LIBRARY path: torch conv2d/linear in bf16 channels_last (cuDNN/cuBLAS)",
    "native": "this repo: tcgen05 implicit-GEMM convolutions + TMA/tcgen05 dense head (snk_qnet_forward), bf16 operands, fp32 accumulate",
}


def available_backends():
    return ["native", "torch_bf16", "torch"]


class QNet:
    def __init__(self, layers, device, dtype=torch.float32, backend="torch"):
        self.backend = backend
        if backend == "torch_bf16":
            dtype = torch.bfloat16
        self.device, self.dtype = torch.device(device), dtype
        self.params = []
        for kind, p in layers:
            if kind == "conv":
                w = torch.from_numpy(bson_io.conv_weight_to_torch(p["W"])).to(self.device, dtype)
                self.params.append(("conv", w, torch.from_numpy(p["b"]).to(self.device, dtype), int(p["pad"][0])))
            elif kind == "dense":
                self.params.append(("dense", torch.from_numpy(np.ascontiguousarray(p["W"])).to(self.device, dtype),
                                    torch.from_numpy(p["b"]).to(self.device, dtype), None))
        self.n_params = sum(w.numel() + b.numel() for _, w, b, _ in self.params)
        self._q = None
        if backend == "native":
            from . import _check, lib
            theta = np.ascontiguousarray(bson_io.destructure(layers), dtype=np.float32)
            self._q = C.c_void_p()
            _check(lib().snk_qnet_create(C.byref(self._q), theta.ctypes.data_as(C.c_void_p), theta.size,
                                         self.device.index or 0))

    def close(self):
        if getattr(self, "_q", None):
            try:
                from . import lib
                lib().snk_qnet_destroy(self._q)
            except Exception:          # interpreter shutdown: module globals may already be gone
                pass
            self._q = None

    __del__ = close

    def forward_native(self, obs):
        """snk_qnet_forward: obs (N,2,10,10) float32 contiguous -> Q (N,3) float32"""
        from . import _check, _ptr, lib
        n = obs.shape[0]
        out = torch.empty(n, 3, dtype=torch.float32, device=self.device)
        st = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        _check(lib().snk_qnet_forward(self._q, _ptr(obs, torch.float32, n * 200, self.device), n, _ptr(out), st))
        return out

    @classmethod
    def from_trainer_bson(cls, path, device, which="q_net", dtype=torch.float32):
        q, t = bson_io.load_trainer_nets(path)
        return cls(q if which == "q_net" else t, device, dtype)

    def forward_torch(self, obs):
        """obs: (N, C, 10, 10) = Julia (10,10,C,N).  Returns Q (N, 3) float32 [= Julia (3, N)]."""
        x = obs.to(self.dtype)
        if self.backend == "torch_bf16":
            x = x.contiguous(memory_format=torch.channels_last)
        convs = [p for p in self.params if p[0] == "conv"]
        denses = [p for p in self.params if p[0] == "dense"]
        for _, w, b, pad in convs:
            x = F.relu(F.conv2d(x, w, b, padding=pad))
        x = x.flatten(1)
        x = F.relu(F.linear(x, denses[0][1], denses[0][2]))
        x = F.linear(x, denses[1][1], denses[1][2])
        return x.float().contiguous()

    def __call__(self, obs):
        return self.forward_native(obs) if self.backend == "native" else self.forward_torch(obs)
