"""-m gpu: operator-level parity of the CUDA kernels (through the C ABI) against the CPU oracle."""
import numpy as np
import pytest

from conftest import crop_D, rel_fro

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def B():
    import background_subtraction_b200 as B
    return B


@pytest.mark.parametrize("m,n", [(640, 7), (20480, 48), (4099, 33), (12000, 130)])
def test_gram_fp64(B, m, n):
    rng = np.random.default_rng(m + n)
    D = np.asfortranarray(rng.standard_normal((m, n)).astype(np.float32).astype(np.float64))
    G = B.gram(D)
    ref = D.T @ D
    assert np.abs(G - ref).max() <= 1e-12 * np.abs(ref).max()
    S = np.asfortranarray(rng.standard_normal((m, n)).astype(np.float32).astype(np.float64))
    Y = np.asfortranarray(rng.standard_normal((m, n)).astype(np.float32).astype(np.float64))
    mu = 3.7
    G2 = B.gram(D, S, Y, mu)
    # the kernel forms W in fp32 as fma(Y, 1/mu, D - S)
    t = (D.astype(np.float32) - S.astype(np.float32)).astype(np.float64)
    W = (Y * np.float64(np.float32(1.0 / mu)) + t).astype(np.float32).astype(np.float64)
    ref2 = W.T @ W
    assert np.abs(G2 - ref2).max() <= 1e-12 * np.abs(ref2).max()


@pytest.mark.parametrize("n,k", [(1, 1), (2, 2), (7, 3), (48, 10), (48, 48), (130, 25), (300, 12), (600, 40)])
def test_eig_topk(B, n, k):
    rng = np.random.default_rng(n * 7 + k)
    # graded spectrum like the ALM Gram matrices: a few large, a long tail ~1e-10 below
    q, _ = np.linalg.qr(rng.standard_normal((n, n)))
    ev = np.concatenate([10.0 ** (-rng.uniform(0, 3, size=min(n, 6))), 10.0 ** (-rng.uniform(8, 11, size=max(n - 6, 0)))])[:n]
    ev = np.sort(ev)[::-1] * 6.4e4
    G = (q * ev) @ q.T
    G = 0.5 * (G + G.T)
    lam, vec = B.eig_topk(G, k)
    w, v = np.linalg.eigh(G)
    w, v = w[::-1], v[:, ::-1]
    assert np.abs(lam - w[:k]).max() <= 1e-13 * abs(w[0]) * max(n, 8), (lam[:5], w[:5])
    # eigenvector quality: residual and orthonormality
    R = G @ vec.T - vec.T * lam
    assert np.abs(R).max() <= 1e-11 * abs(w[0]), np.abs(R).max()
    assert np.abs(vec @ vec.T - np.eye(k)).max() <= 1e-9


def test_eig_topk_clustered(B):
    n = 64
    rng = np.random.default_rng(3)
    q, _ = np.linalg.qr(rng.standard_normal((n, n)))
    ev = np.ones(n) * 1e-6
    ev[:5] = [5.0, 2.0, 2.0, 2.0 - 1e-13, 1.0]
    G = (q * ev) @ q.T
    G = 0.5 * (G + G.T)
    lam, vec = B.eig_topk(G, 6)
    w = np.linalg.eigvalsh(G)[::-1]
    assert np.abs(lam - w[:6]).max() <= 1e-13 * 5 * n
    assert np.abs(vec @ vec.T - np.eye(6)).max() <= 1e-8
    assert np.abs(G @ vec.T - vec.T * lam).max() <= 1e-10


@pytest.mark.parametrize("shape", [(32, 40, 5), (31, 41, 3), (128, 160, 4), (7, 5, 2)])
def test_prox_flat(B, shape):
    from oracle import alm_oracle as O
    h, w, t = shape
    rng = np.random.default_rng(h * w)
    U = np.asfortranarray((rng.standard_normal((h * w, t)) * 0.05).astype(np.float32).astype(np.float64))
    groups = B.get_proximal_flat_groups_nonoverlap((h, w), (3, 3))
    for lam in (0.04, 0.3, 1e-4):
        ref = O.prox_flat(U, lam, groups)
        out = B.prox_flat(U, lam, groups)
        assert np.abs(out - ref).max() <= 2e-7 * max(np.abs(U).max(), lam), (lam, np.abs(out - ref).max())
    # arbitrary partition (generic CSR path): random group ids incl. id 0 = no group
    g2 = rng.integers(0, 40, size=h * w).astype(np.int32)
    ref = O.prox_flat(U, 0.1, g2)
    out = B.prox_flat(U, 0.1, g2)
    assert np.abs(out - ref).max() <= 5e-7


def test_prox_flat_golden(B, golden_cases):
    G = golden_cases["bs_G"]
    groups = B.get_proximal_flat_groups_nonoverlap((32, 40), (3, 3))
    out = B.prox_flat(G, 0.04, groups)
    assert np.abs(out - golden_cases["pf_out"]).max() <= 2e-7 * np.abs(G).max()


def test_block_shrink(B, golden_cases):
    from oracle import alm_oracle as O
    G = golden_cases["bs_G"]
    labels = golden_cases["gs_a_labels"].astype(np.int32)
    ptr = golden_cases["gs_a_lam_ptr"]
    lam = golden_cases["gs_a_lam"]
    n, m = labels.shape
    blocks = [[labels[f] == b + 1 for b in range(ptr[f + 1] - ptr[f])] for f in range(n)]
    lambdas = [[lam[ptr[f] + b] for b in range(ptr[f + 1] - ptr[f])] for f in range(n)]
    out = B.block_shrinkage_operator(G, blocks, lambdas, 3.0, 0.02)
    assert rel_fro(out, golden_cases["bs_out"]) <= 1e-6
    assert np.abs(out - golden_cases["bs_out"]).max() <= 1e-6 * np.abs(G).max()


def test_foreground_mask(B, golden_cases, watersurface_u8):
    D, shp = crop_D(watersurface_u8, golden_cases["flat_a_crop"])
    L, S = golden_cases["flat_a_L"], golden_cases["flat_a_S"]
    mask = B.foreground_mask(D, L, S)
    ref = np.unpackbits(golden_cases["flat_a_mask"])[:D.size].reshape(D.shape, order='F').astype(bool)
    assert (mask == ref).mean() >= 0.999, (mask == ref).mean()
    for k in (2, 3):
        D2, _ = crop_D(watersurface_u8, golden_cases["gs_a_crop"])
        mk = B.foreground_mask(D2, golden_cases["gs_a_L"], golden_cases["gs_a_S"], k)
        rf = np.unpackbits(golden_cases["gs_a_mask%d" % k])[:D2.size].reshape(D2.shape, order='F').astype(bool)
        assert (mk == rf).mean() >= 0.999


@pytest.mark.parametrize("shape", [(24, 30, 3), (9, 7, 2), (40, 33, 2)])
def test_prox_graph(B, shape):
    from oracle import alm_oracle as O
    h, w, t = shape
    rng = np.random.default_rng(h + w)
    U = np.asfortranarray((rng.standard_normal((h * w, t)) * 0.05).astype(np.float32).astype(np.float64))
    graph = B.getGraphSPAMS_all_groups((h, w), (3, 3))
    gc = O.graph_from_spams_dict(graph)
    for lam in (0.004, 0.05):
        ref = O.prox_graph(U, lam, gc, tol=1e-12)
        out, sw = B.prox(U, lam, graph, return_sweeps=True, tol=1e-6)
        err = np.abs(out - ref).max()
        assert err <= 2e-5 * max(np.abs(U).max(), lam), (lam, err, sw)
        # uncovered last row / column is the identity
        out3 = out.reshape((h, w, t), order='F')
        U3 = U.reshape((h, w, t), order='F')
        assert np.array_equal(out3[-1, :, :].astype(np.float32), U3[-1, :, :].astype(np.float32))
        assert np.array_equal(out3[:, -1, :].astype(np.float32), U3[:, -1, :].astype(np.float32))


def test_prox_by_frame_center_windows(B):
    """prox_by_frame (inexact_alm_lsd.py:60-68) with the per-frame centre-window graphs of get_proximal_graph_group_centers
    (lsd_improvement.py:74-120): weights 1 / 1.5, exact border clipping, pixels outside every window untouched."""
    from oracle import alm_oracle as O
    h, w, t = 18, 23, 4
    rng = np.random.default_rng(9)
    U = np.asfortranarray((rng.standard_normal((h * w, t)) * 0.05).astype(np.float32).astype(np.float64))
    weights = [np.where(rng.random((h, w)) < 0.3, rng.choice([1.0, 1.5], (h, w)), -1.0) for _ in range(t)]
    for wm in weights:
        wm[0, 0], wm[-1, -1], wm[0, w // 2] = 1.5, 1.0, 1.0           # windows clipped at corners and an edge
    graphs = [B.get_proximal_graph_group_centers((h, w), 1, wm) for wm in weights]
    bare = [{k: v for k, v in g.items() if not k.startswith('_')} for g in graphs]      # as the reference would hand them over
    gcs = [O.graph_group_centers((h, w), 1, wm) for wm in weights]
    for lam in (0.004, 0.05):
        ref = O.prox_by_frame(U, lam, gcs, tol=1e-12)
        for gl in (graphs, bare):
            out = B.prox_by_frame(U, lam, gl)
            assert np.abs(out - ref).max() <= 2e-5 * max(np.abs(U).max(), lam)
        one = B.prox(U[:, [1]], lam, bare[1])                           # single-graph entry point with a centre graph
        assert np.abs(one - ref[:, [1]]).max() <= 2e-5 * max(np.abs(U).max(), lam)
        for f in range(t):                                              # uncovered pixels: identity
            cov = np.zeros(h * w, dtype=bool)
            cov[gcs[f][1]] = True
            assert np.array_equal(out[~cov, f].astype(np.float32), U[~cov, f].astype(np.float32))


def test_prox_graph_golden(B, golden_cases):
    G = golden_cases["bs_G"][:, :4]
    graph = B.getGraphSPAMS_all_groups((32, 40), (3, 3))
    out = B.prox(G, 0.04, graph, tol=1e-6)
    assert np.abs(out - golden_cases["pg_out"]).max() <= 2e-5 * np.abs(G).max()


def test_svd_k_largest(B):
    rng = np.random.default_rng(5)
    A = np.asfortranarray((rng.standard_normal((3000, 4)) @ rng.standard_normal((4, 20)) + 0.01 * rng.standard_normal((3000, 20))))
    A = A.astype(np.float32).astype(np.float64)
    u, s, vh = B.svd_k_largest(A, 6)
    s_ref = np.linalg.svd(A, compute_uv=False)[:6]
    assert np.abs(s - s_ref).max() <= 1e-9 * s_ref[0]
    assert rel_fro((u * s) @ vh, (lambda U, S, V: (U[:, :6] * S[:6]) @ V[:6])(*np.linalg.svd(A, full_matrices=False))) <= 1e-8


@pytest.mark.parametrize("n,ldq", [(44, 256), (128, 64 * 7), (130, 64 * 40), (300, 64 * 33), (300, 64 * 1000), (257, 64 * 50),
                                   (384, 64 * 9), (600, 64 * 20)])
def test_gram_i8_exact(B, n, ldq):
    """tcgen05 int8 slice Gram: exact integer equality with NumPy (classes i + j >= 3 of the 4x4 slice pairs)."""
    import ctypes
    from background_subtraction_b200 import _cabi as C
    rng = np.random.default_rng(n + ldq)
    sl = rng.integers(-128, 128, size=(4, n, ldq), dtype=np.int8)
    G = np.empty((n, n), dtype=np.int64)
    C.check(C.load().bsub_gram_i8_test(sl.ctypes.data_as(ctypes.c_void_p), n, ldq, G.ctypes.data_as(ctypes.c_void_p)))
    ref = np.zeros((n, n), dtype=np.int64)
    s64 = sl.astype(np.int64)
    for i in range(4):
        for j in range(4):
            if i + j >= 3:
                ref += (s64[i] @ s64[j].T) << (8 * (i + j - 3))
    assert np.array_equal(G, ref), (np.abs(G - ref).max(), np.argwhere(G != ref)[:5])


@pytest.mark.parametrize("shape,xi_mb", [((70, 95, 5), None), ((70, 95, 5), "1"), ((33, 64, 3), None), ((131, 45, 2), None)])
def test_prox_graph_tiled_multi_tile(B, shape, xi_mb, monkeypatch):
    """The tile-local dual BCD (prox_graph3_tile_kernel): images of several 32-pixel tiles in both directions, so that windows
    straddle tile borders in every one of the three shifted tilings; with BSUB_GRAPH_XI_MB=1 the frames are walked in chunks
    whose duals fit a 1 MB buffer (2 chunks here).  Checked against the oracle's sequential sweeps at 1e-12."""
    from oracle import alm_oracle as O
    h, w, t = shape
    if xi_mb is not None:
        monkeypatch.setenv("BSUB_GRAPH_XI_MB", xi_mb)
    rng = np.random.default_rng(h * w)
    U = (rng.standard_normal((h * w, t)) * 0.05)
    U[rng.random((h * w, t)) < 0.02] += 0.4                     # a few strong pixels, like foreground
    U = np.asfortranarray(U.astype(np.float32).astype(np.float64))
    graph = B.getGraphSPAMS_all_groups((h, w), (3, 3))
    gc = O.graph_from_spams_dict(graph)
    for lam in (0.003, 0.08):
        ref = O.prox_graph(U, lam, gc, tol=1e-12)
        out, sw = B.prox(U, lam, graph, return_sweeps=True, tol=1e-6)
        err = np.abs(out - ref).max()
        assert err <= 2e-5 * max(np.abs(U).max(), lam), (lam, err, sw)
        assert 1 <= sw < 20000
