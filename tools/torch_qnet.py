"""LIBRARY restatement of the reference Q-network (torch conv2d / linear = cuDNN / cuBLAS) — NOT part of the product.

Used (a) by the tests as a second opinion next to the Float64 numpy oracle (oracle/qnet_oracle.py), on CPU in Float64 and
on the GPU in Float32, and (b) by bench.py / tools as the library timing baseline the native kernels are compared with.
"""
import numpy as np
import torch
import torch.nn.functional as F


def conv_weight_to_torch(W):
    """Flux (k1,k2,cin,cout) true-convolution kernel -> torch cross-correlation weight [cout,cin,kh,kw] for
    inputs stored (N,C,d2,d1):  wt[o,c,kh,kw] = W[K1-1-kw, K2-1-kh, c, o]."""
    return np.ascontiguousarray(np.transpose(W[::-1, ::-1, :, :], (3, 2, 1, 0)))


class TorchQNet:
    def __init__(self, layers, device, dtype=torch.float32, channels_last=False, allow_tf32=False):
        self.device, self.dtype, self.channels_last, self.allow_tf32 = torch.device(device), dtype, channels_last, allow_tf32
        self.params = []
        for kind, p in layers:
            if kind == "conv":
                w = torch.from_numpy(conv_weight_to_torch(p["W"])).to(self.device, dtype)
                self.params.append(("conv", w, torch.from_numpy(p["b"]).to(self.device, dtype), int(p["pad"][0])))
            elif kind == "dense":
                self.params.append(("dense", torch.from_numpy(np.ascontiguousarray(p["W"])).to(self.device, dtype),
                                    torch.from_numpy(p["b"]).to(self.device, dtype), None))

    def __call__(self, obs):
        """obs: (N, C, 10, 10) = Julia (10,10,C,N).  Returns Q (N, 3) float32 [= Julia (3, N)]."""
        tf32 = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
        torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = self.allow_tf32
        try:
            x = obs.to(self.dtype)
            if self.channels_last:
                x = x.contiguous(memory_format=torch.channels_last)
            convs = [p for p in self.params if p[0] == "conv"]
            denses = [p for p in self.params if p[0] == "dense"]
            for _, w, b, pad in convs:
                x = F.relu(F.conv2d(x, w, b, padding=pad))
            x = x.flatten(1)
            x = F.relu(F.linear(x, denses[0][1], denses[0][2]))
            x = F.linear(x, denses[1][1], denses[1][2])
            return x.to(torch.float32 if self.dtype != torch.float64 else torch.float64).contiguous()
        finally:
            torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32
