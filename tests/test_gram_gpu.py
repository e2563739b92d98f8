"""Gram of the deviation matrix on tcgen05 vs the Float64 numpy oracle (tolerances stated per test)."""
import numpy as np
import pytest
import torch

from oracle import gram_oracle as GO
from oracle import oracle_lib as O
from tests.util import pkg

pytestmark = pytest.mark.gpu


def _rel_fro(G, ref):
    return float(np.linalg.norm(G.astype(np.float64) - ref) / np.linalg.norm(ref))


@pytest.mark.parametrize("block_k", [32, 64])
@pytest.mark.parametrize("K,P", [(128, 256), (256, 4096), (200, 1000), (1000, 5003), (58, 181395), (384, 77)])
def test_gram_matches_fp64_oracle(K, P, block_k):
    S = pkg()
    rng = np.random.default_rng(K * 7 + P)
    A = rng.normal(0, 1, (K, P)) * np.exp(rng.normal(0, 1, (K, 1)))         # rows of very different scale
    ref = GO.gram(A)
    dA = torch.from_numpy(A).cuda()
    plan = S.GramPlan(K, P, dA.device).pack(dA)
    G3 = plan.gram(terms=3, block_k=block_k).cpu().numpy()
    G1 = plan.gram(terms=1, block_k=block_k).cpu().numpy()
    scale = np.sqrt(np.outer(np.diag(ref), np.diag(ref)))                   # |g_ij| <= sqrt(g_ii g_jj)
    # bf16 hi/lo split: inputs good to 2^-17, products to ~2^-16, fp32 accumulation over P terms
    assert np.max(np.abs(G3 - ref) / scale) < 2e-5, np.max(np.abs(G3 - ref) / scale)
    assert _rel_fro(G3, ref) < 1e-5
    assert np.array_equal(G3, G3.T) and np.array_equal(G1, G1.T)           # upper triangle mirrored: exactly symmetric
    # plain bf16 inputs: 2^-9 per operand
    assert np.max(np.abs(G1 - ref) / scale) < 8e-3
    assert _rel_fro(G1, ref) < 4e-3


@pytest.mark.parametrize("cta_group", [1, 2])
@pytest.mark.parametrize("K,P", [(2500, 515), (1000, 2048), (130, 300)])
def test_both_tile_engines_and_multi_panel_triangles(cta_group, K, P):
    """the single-CTA (128 x 256 tiles) and the CTA-pair (256 x 256) engine on upper-triangle tile lists that span more than
    one column panel (K = 2500: 10 tile columns)"""
    S = pkg()
    rng = np.random.default_rng(K + cta_group)
    A = rng.normal(0, 1, (K, P))
    ref = GO.gram(A)
    dA = torch.from_numpy(A).cuda()
    S.lib().snk_gram_config(cta_group)
    try:
        plan = S.GramPlan(K, P, dA.device).pack(dA)
        G3 = plan.gram(terms=3).cpu().numpy()
        G1 = plan.gram(terms=1).cpu().numpy()
    finally:
        S.lib().snk_gram_config(2)
    assert _rel_fro(G3, ref) < 1e-5 and _rel_fro(G1, ref) < 4e-3
    assert np.array_equal(G3, G3.T)


def test_gram_of_centred_snapshots_spectrum():
    """compute_D-shaped input at test size: K snapshots of a drifting random walk, centred by
    snk_center_columns (bit-exact vs the oracle), Gram by tcgen05; spectrum vs eig of the Float64 Gram."""
    S = pkg()
    K, P = 256, 20011
    A = GO.synthetic_snapshots(K, P, seed=3)
    want = A.copy().reshape(-1)
    O.center_columns(want, P, K)
    dA = torch.from_numpy(A).cuda()
    S.center_columns(dA)
    assert np.array_equal(dA.cpu().numpy().reshape(-1), want)
    ref = GO.gram(want.reshape(K, P))
    G = S.gram(dA, terms=3).cpu().numpy().astype(np.float64)
    lam, lam_ref = GO.spectrum(G, K), GO.spectrum(ref, K)
    top = lam_ref > 1e-4 * lam_ref[0]
    assert top.sum() >= 5
    assert np.max(np.abs(lam[top] - lam_ref[top]) / lam_ref[top]) < 1e-3
    assert abs(lam.sum() - lam_ref.sum()) / lam_ref.sum() < 1e-5           # trace = total variance


def test_gram_float32_input_and_explicit_splits():
    S = pkg()
    K, P = 300, 9000
    A = np.random.default_rng(1).normal(0, 1, (K, P)).astype(np.float32)
    ref = GO.gram(A)
    dA = torch.from_numpy(A).cuda()
    for splits in (1, 3, 7):
        G = S.gram(dA, terms=3, splits=splits).cpu().numpy()
        assert _rel_fro(G, ref) < 1e-5, splits


@pytest.mark.parametrize("world,K,P,terms", [(2, 512, 3000, 3), (3, 700, 5003, 3), (4, 1000, 2048, 1), (8, 1000, 9999, 3), (2, 2500, 1031, 3),
                                             (5, 33, 700, 3), (6, 1800, 515, 3)])
def test_row_sharded_gram_with_virtual_ranks(world, K, P, terms):
    """The multi-GPU algorithm (half planes ring -> block Grams -> transposed-peer mirror) with R virtual ranks
    in one process on one GPU: same kernels and pointer arithmetic, local 'peer' memory."""
    S = pkg()
    from snake_b200 import gram_sharded as GS
    rng = np.random.default_rng(world * 100 + K)
    A = rng.normal(0, 1, (K, P)) * np.exp(rng.normal(0, 0.5, (K, 1)))
    ref = GO.gram(A)
    peers = GS.LocalPeers(K, P, world, torch.device("cuda", 0))
    G = peers.gram(torch.from_numpy(A).cuda(), terms=terms).cpu().numpy()
    peers.free()
    assert G.shape == (K, K)
    tol = 1e-5 if terms == 3 else 4e-3
    assert _rel_fro(G, ref) < tol
    single = S.gram(torch.from_numpy(A).cuda(), terms=terms).cpu().numpy()
    assert _rel_fro(G, single.astype(np.float64)) < 2e-6       # same arithmetic up to split-K summation order
    assert np.array_equal(G, G.T)                              # every block below the ring's half is a copy of its mirror image


def test_transpose_block_ragged_sizes_and_leading_dimensions():
    """snk_gram_transpose_block (the mirror step of the sharded Gram): G (rows_a x rows_b, ld) = YT^T for sizes that are not
    multiples of the 32 x 32 staging tile, inside larger buffers (the untouched surroundings stay untouched)"""
    import ctypes as C
    S = pkg()
    L = S.lib()
    g = torch.Generator(device="cuda").manual_seed(0)
    for ra, rb in [(1, 1), (33, 65), (100, 7), (250, 313)]:
        ldyt, ldg = ra + 5, rb + 11
        YT = torch.randn(rb, ldyt, device="cuda", generator=g)
        G = torch.full((ra + 2, ldg), -7.0, device="cuda")
        S._check(L.snk_gram_transpose_block(C.c_void_p(YT.data_ptr()), ldyt, ra, rb, C.c_void_p(G.data_ptr()), ldg,
                                            C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        torch.cuda.synchronize()
        assert torch.equal(G[:ra, :rb], YT[:, :ra].T)
        assert bool((G[ra:] == -7.0).all()) and bool((G[:, rb:] == -7.0).all())


def test_sharded_run_with_device_barriers_two_virtual_ranks_on_two_streams():
    """snk_gram_shard_run — pack, planes ring, mirror separated by the device-side peer barrier — for two virtual ranks of
    one process, each on its own stream: rank 0's barrier kernel spins until rank 1's stream reaches its barrier.  Run twice
    (the barrier epochs keep counting)."""
    S = pkg()
    from snake_b200 import gram_sharded as GS
    K, P = 600, 4099
    rng = np.random.default_rng(3)
    A = rng.normal(0, 1, (K, P))
    ref = GO.gram(A)
    dev = torch.device("cuda", 0)
    peers = GS.LocalPeers(K, P, 2, dev)
    dA = torch.from_numpy(A).cuda()
    streams = [torch.cuda.Stream(dev) for _ in range(2)]
    torch.cuda.synchronize()
    for rep in range(2):
        outs = []
        for s, st in zip(peers.shards, streams):
            with torch.cuda.stream(st):
                lo = s.col0[s.rank]
                outs.append(s.run(dA[lo:lo + s.rows].contiguous(), terms=3))
        torch.cuda.synchronize()
        for s in peers.shards:
            s.check()                                    # no barrier timed out
        G = torch.cat(outs, 0).cpu().numpy()
        assert _rel_fro(G, ref) < 1e-5, rep
    peers.free()


def test_plot_traj_numeric_path_matches_numpy_svd():
    """compute_D.jl + plot_traj.jl end to end: snapshots -> D -> centring -> spectrum S.^2/(K-1), the 99 % column
    count and the two-direction trajectory U[:, 1:2]' D, against numpy's SVD of the centred Float64 D."""
    S = pkg()
    K, P = 200, 30011
    A = GO.synthetic_snapshots(K, P, seed=11)                    # (K, P) float64, values exactly representable in f32
    dm = S.laplace.DeviationMatrix(P, K, "cuda:0")
    for k in range(K):
        dm.store(torch.from_numpy(A[k].astype(np.float32)).cuda())
    assert dm.position == K and np.array_equal(dm.Dt.cpu().numpy(), A)
    with pytest.raises(IndexError):
        dm.store(torch.zeros(P, device="cuda"))
    dm.center()
    want = A.copy().reshape(-1)
    O.center_columns(want, P, K)
    D = want.reshape(K, P)                                       # rows = snapshots; Julia's D is this transposed
    assert np.array_equal(dm.Dt.cpu().numpy(), D)
    U, Sv, Vt = np.linalg.svd(D.T, full_matrices=False)          # svd(deviation_matrix), plot_traj.jl:10
    lam_ref = Sv ** 2 / (K - 1)
    lam, V, w = S.laplace.spectrum(dm.Dt)
    lam = lam.cpu().numpy()
    top = lam_ref > 1e-4 * lam_ref[0]
    assert np.max(np.abs(lam[top] - lam_ref[top]) / lam_ref[top]) < 1e-3
    cum = np.cumsum(lam_ref)
    n_ref = int(np.searchsorted(cum, 0.99 * cum[-1]) + 1)
    assert abs(S.laplace.n_cols_for_variance(torch.from_numpy(lam).cuda()) - n_ref) <= 1
    Y = S.laplace.trajectory(w, V, 2).cpu().numpy()
    Y_ref = U[:, :2].T @ D.T                                     # (2, K)
    for i in range(2):
        sgn = np.sign(np.dot(Y[i], Y_ref[i]))
        assert np.max(np.abs(sgn * Y[i] - Y_ref[i])) < 2e-3 * np.abs(Y_ref[i]).max()


def test_deviation_matrix_from_a_loaded_file_shape():
    """plot_traj.jl:7 starts from a stored D (P x K Float64, already centred): from_numpy keeps Julia's column k as row k and the
    spectrum of the loaded matrix matches numpy's SVD."""
    S = pkg()
    K, P = 64, 5003
    A = GO.synthetic_snapshots(K, P, seed=3)
    A = A - A.mean(axis=0, keepdims=True)                         # centred, as compute_D.jl:84 saves it
    dm = S.laplace.DeviationMatrix.from_numpy(np.asfortranarray(A.T), "cuda:0")     # Julia shape (P, K)
    assert dm.position == K and np.array_equal(dm.Dt.cpu().numpy(), A)
    lam, V, w = S.laplace.spectrum(dm.Dt)
    lam_ref = np.linalg.svd(A.T, compute_uv=False) ** 2 / (K - 1)
    top = lam_ref > 1e-4 * lam_ref[0]
    assert np.max(np.abs(lam.cpu().numpy()[top] - lam_ref[top]) / lam_ref[top]) < 1e-3


def test_sample_model_weights_matches_numpy():
    """la_utils.jl:83-95 with injected z1, z2 (the reference draws them from MvNormal(0, I))."""
    S = pkg()
    rng = np.random.default_rng(4)
    K, P = 58, 181395                                            # the reference insists on K = 58 (la_utils.jl:86)
    D = rng.normal(0, 1e-2, (K, P))
    mean, var = rng.normal(0, 0.1, P), rng.normal(0, 1e-4, P)    # some negative "variances": compute_Gamma_diag takes abs
    z1, z2 = rng.normal(0, 1, P), rng.normal(0, 1, K)
    want = mean + np.sqrt(np.abs(var)) * z1 / np.sqrt(2) + (D.T @ z2) / np.sqrt(2 * (K - 1))
    t = lambda a: torch.from_numpy(a).cuda()
    got = S.laplace.sample_model_weights(t(mean), t(var), t(D), t(z1), t(z2)).cpu().numpy()
    assert np.max(np.abs(got - want)) < 1e-13 * np.max(np.abs(want)) + 1e-15
    # the sampled weights are a valid Q-net
    net = S.qnet.QNet(S.qnet.glorot_layers(0), "cuda:0", precision="f32")
    assert got.shape == (net.n_params,)
