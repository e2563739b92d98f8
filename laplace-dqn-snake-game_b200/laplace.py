"""The numeric path of compute_D.jl + plot_traj.jl on the device (plots excluded).

compute_D.jl collects K snapshots of the flattened Q-net weights into D (P x K Float64), centres it
(`fit!` / `D .-= avg`, :66-81); plot_traj.jl takes svd(D) and uses lambda = S.^2/(K-1) (:10-16), the number of
directions that hold 99 % of the variance (:48-66) and the trajectory Y = U[:, 1:2]' * D (:69-70).

Here: snapshots -> snk_d_store_snapshot, centring -> snk_center_columns (bit-exact Float64 Welford), and instead of
an SVD of the 181,395 x 1000 matrix the K x K Gram G = D'D on the tensor cores (snk_gram).  With G = V diag(w) V':
S = sqrt(w), lambda = w/(K-1), and U' D = S V', so Y[i, :] = sqrt(w_i) * v_i — the trajectory needs no second pass
over D.  The K x K symmetric eigenproblem is a torch.linalg call (LIBRARY; a 1000 x 1000 matrix, off the hot path).
"""
import ctypes as C

import torch

from . import GramPlan, _check, _ptr, center_columns, lib


class DeviationMatrix:
    """D as Julia stores it (P x K column-major) = a (K, P) torch-contiguous float64 CUDA tensor."""

    def __init__(self, P, K, device):
        self.P, self.K = int(P), int(K)
        self.device = torch.device(device)
        self.Dt = torch.zeros(self.K, self.P, dtype=torch.float64, device=self.device)   # zeros(Float64, (P, K))
        self.position = 0                                                                # compute_D.jl:51 (0-based here)
        self.mean = self.var = None

    def store(self, theta):
        """deviation_matrix[:, position] = Float64.(theta); position += 1   (compute_D.jl:67-71)"""
        if self.position >= self.K:
            raise IndexError("all %d snapshots are already collected" % self.K)
        st = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        with torch.cuda.device(self.device):
            _check(lib().snk_d_store_snapshot(_ptr(self.Dt), self.P, self.K, self.position,
                                              _ptr(theta, torch.float32, self.P, self.device), st))
        self.position += 1

    @classmethod
    def from_numpy(cls, D, device):
        """An already collected D (P x K Float64 in Julia's shape, e.g. bson_io.load_deviation_matrix of a
        ./D_matrices/*.bson written by compute_D.jl:84) -> device; plot_traj.jl:7 starts from such a file."""
        import numpy as np
        D = np.asarray(D, dtype=np.float64)
        if D.ndim != 2:
            raise ValueError("D must be a (P, K) matrix")
        P, K = D.shape
        self = cls(P, K, device)
        self.Dt.copy_(torch.from_numpy(np.ascontiguousarray(D.T)))      # Julia column k = row k here
        self.position = K
        return self

    def center(self):
        """Welford over the columns, D .-= mean   (compute_D.jl:76-81); returns (mean, var)"""
        self.mean, self.var = center_columns(self.Dt)
        return self.mean, self.var


def spectrum(Dt, terms=3):
    """(lambda descending (K,), eigenvectors V (K, K) columns matching) of D'D/(K-1) for a CENTRED D.
    lambda == S.^2/(K-1) of plot_traj.jl:10-16."""
    K, P = Dt.shape
    G = GramPlan(K, P, Dt.device).pack(Dt).gram(terms=terms).double()
    w, V = torch.linalg.eigh(G)                       # ascending
    w = torch.clamp(w.flip(0), min=0.0)
    V = V.flip(1)
    return w / max(K - 1, 1), V, w


def n_cols_for_variance(lam, fraction=0.99):
    """compute_n_cols (plot_traj.jl:48-66): smallest n with sum(lambda[1:n]) >= fraction * sum(lambda)."""
    cum = torch.cumsum(lam, 0)
    lim = fraction * lam.sum()
    return int((cum < lim).sum().item()) + 1


def trajectory(w, V, n=2):
    """Y = U[:, 1:n]' * D  (plot_traj.jl:69-70) = sqrt(w_i) * v_i', shape (n, K); rows are defined up to sign."""
    return (V[:, :n] * torch.sqrt(w[:n])).T.contiguous()


def sample_model_weights(mean, var, Dt, z1=None, z2=None):
    """sample_model (la_utils.jl:83-95): w = mean + sqrt.(|var|) .* z1 / sqrt(2) + D * z2 / sqrt(2 (K-1)).
    Dt: centred (K, P) float64; z1 (P), z2 (K) standard-normal draws (generated with torch if not injected).
    Returns the flattened weights (P,) float64 in Flux.destructure order (feed QNet / snk_qnet_create after .float())."""
    K, P = Dt.shape
    dev = Dt.device
    z1 = torch.randn(P, dtype=torch.float64, device=dev) if z1 is None else z1
    z2 = torch.randn(K, dtype=torch.float64, device=dev) if z2 is None else z2
    w = torch.empty(P, dtype=torch.float64, device=dev)
    st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    with torch.cuda.device(dev):
        _check(lib().snk_laplace_sample_weights(_ptr(mean, torch.float64, P, dev), _ptr(var, torch.float64, P, dev),
                                                _ptr(Dt, torch.float64), P, K, _ptr(z1, torch.float64, P, dev),
                                                _ptr(z2, torch.float64, K, dev), _ptr(w), st))
    return w
