"""Oracle (test infrastructure) for the Laplace deviation-matrix path, in Float64 numpy.

compute_D.jl:66-81 builds D (P x K Float64: K snapshots of the flattened Q-net weights as columns, centred);
plot_traj.jl:10-16 takes svd(D) and lambda = S.^2/(K-1), the spectrum of D D'/(K-1), equal to the non-zero
spectrum of the K x K Gram D' D/(K-1).  Parity is UNPINNED by the reference: no D_matrices/*.bson is
committed and Julia cannot run here, so the oracle is this restatement alone.
"""
import numpy as np


def synthetic_snapshots(K, P, seed=0, drift=1e-3, step=2e-3):
    """K snapshots of a P-dim random walk with drift along RMSProp-like trajectories:
    gives a power-law-like spectrum as in images/correlation_histo.png.  Returned as A = D^T (K, P) Float64,
    values rounded through Float32 first because the reference stores Float64.(theta::Float32)."""
    rng = np.random.default_rng(seed)
    base = rng.normal(0, 0.05, P)
    dirn = rng.normal(0, 1, P) * drift
    steps = rng.normal(0, step, (K, P)) * (rng.random(P) < 0.3)
    A = base + np.cumsum(steps, axis=0) + np.arange(K)[:, None] * dirn
    return np.ascontiguousarray(A.astype(np.float32).astype(np.float64))


def gram(A):
    """G = A A^T in Float64 (A = D^T)."""
    A = np.asarray(A, dtype=np.float64)
    return A @ A.T


def spectrum(G, K):
    """lambda = eig(G)/(K-1), descending (== S.^2/(K-1) of svd(D), plot_traj.jl:16)."""
    w = np.linalg.eigvalsh(G)[::-1]
    return w / max(K - 1, 1)
