"""-m gpu: whole-solve parity at the shape classes the round-1 review found uncovered ("Parity gaps"):
  * BASELINE config 2: the converged group-sparse solve on the committed highway fixture (golden_summary.json was
    produced by the reference's own inexact_alm_group_sparse_RPCA, tests/golden/make_golden.py) + masks k=2,3 +
    the compute_score F pair;
  * n = 300 frames: the shape class of the headline bench (3-block multicast int8 Gram + streamed shrink<8,48,28> +
    8-CTA eigensolver together) against the CPU oracle;
  * n = 600 (5-block Gram, 16-CTA eigensolver) against the oracle, not only against the fallback kernels;
  * rank sequences, ||D||_2 and the row-sum norm asserted against the reference-generated goldens;
  * Y after one iteration for generic groups (Y0 must not be dropped on the SPILL path).
Every call goes Python mirror -> ctypes -> C ABI (libbsub_b200.so)."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, rel_fro

pytestmark = pytest.mark.gpu

TOL_F = 1e-4          # north_star: rel-Frobenius(L), (S) <= 1e-4 with fp32 storage


@pytest.fixture(scope="module")
def B():
    import background_subtraction_b200 as B
    return B


@pytest.fixture(scope="module")
def summary():
    with open(os.path.join(GOLDEN, "golden_summary.json")) as f:
        return json.load(f)


def _synthetic_D(rows, cols, n, seed, n_rect):
    from background_subtraction_b200 import synth
    video, gt = synth.make_clip(rows, cols, n, seed=seed, n_rect=n_rect, return_gt=True)
    return np.asfortranarray(synth.preprocess_u8(video).T.astype(np.float64)), gt


def test_group_sparse_highway_converged(B, highway_fixture, summary):
    """precomputed_main.py:64-74 on the committed fixture: D -> inexact_alm_group_sparse_RPCA(blocks, lambdas) ->
    foreground_mask k = 2, 3; reference run: 17 iterations, converged."""
    from oracle import alm_oracle as O
    from oracle import score_oracle as SC
    gold = summary["highway_half_group_sparse"]
    frames = highway_fixture["frames"]
    h, w, t = frames.shape
    D, x_norm, _mean = O.normalize_and_center(frames)
    labels, ptr, lam = highway_fixture["labels"], highway_fixture["lam_ptr"], highway_fixture["lam"]
    assert int(ptr[-1]) == gold["blocks_total"] == len(lam)
    dec = B.group_sparse_decomposition(D, None, None, delta=10, labels=(labels, ptr, np.append(lam, 0.0)))
    st, log = dec.status(), dec.log()
    L, S = dec.download('L'), dec.download('S')
    print("[highway gs] iters", st.iter, "conv", st.converged, "svp", [l['svp'] for l in log])
    print("[highway gs] err tail gpu", ["%.3e" % l['err'] for l in log][-3:], "ref", gold["err"][-3:])
    assert st.iter == gold["iters"] and bool(st.converged) == gold["converged"]
    assert [l['svp'] for l in log] == gold["svp"]
    assert np.allclose([l['err'] for l in log], gold["err"], rtol=2e-3)          # the golden log keeps 4 digits
    assert abs(np.linalg.norm(L) - gold["normL"]) <= 1e-5 * gold["normL"]
    assert abs(np.linalg.norm(S) - gold["normS"]) <= 1e-5 * gold["normS"]
    masks = {}
    for k, key in ((2, "gs_mask2"), (3, "gs_mask3")):
        mk = dec.mask(k)
        ref = np.unpackbits(highway_fixture[key])[:D.size].reshape(D.shape, order='F').astype(bool)
        agree = float((mk == ref).mean())
        print("[highway gs] mask k=%d fraction gpu %.6f ref %.6f agreement %.6f" % (k, mk.mean(), gold["mask%d_fraction" % k], agree))
        assert agree >= 0.999 and abs(float(mk.mean()) - gold["mask%d_fraction" % k]) <= 1e-4
        masks[k] = (mk, ref)
    # full matrices against the oracle (itself pinned to the reference loop in tests/test_oracle.py)
    blocks = [[labels[f] == b + 1 for b in range(ptr[f + 1] - ptr[f])] for f in range(t)]
    lambdas = [[lam[ptr[f] + b] for b in range(ptr[f + 1] - ptr[f])] for f in range(t)]
    Lr, Sr, itr, convr = O.inexact_alm_group_sparse_RPCA(D, blocks, lambdas, delta=10)
    print("[highway gs] relF(L) %.3e relF(S) %.3e" % (rel_fro(L, Lr), rel_fro(S, Sr)))
    assert itr == gold["iters"] and convr
    assert rel_fro(L, Lr) <= TOL_F and rel_fro(S, Sr) <= TOL_F
    # compute_score.py F-measure of both masks against an implementation-independent pseudo ground truth
    # (temporal-median difference): |F_gpu - F_ref| <= 0.001
    med = np.median(x_norm, axis=2, keepdims=True)
    gt = np.abs(x_norm - med) > 0.12
    cube = lambda a: np.asarray(a).reshape((h, w, t), order='F')          # noqa: E731
    for k in (2, 3):
        f_gpu, f_ref = SC.mean_fscore(cube(masks[k][0]), gt), SC.mean_fscore(cube(masks[k][1]), gt)
        print("[highway gs] k=%d F gpu %.4f ref %.4f" % (k, f_gpu, f_ref))
        assert abs(f_gpu - f_ref) <= 1e-3


def test_highway_flat_lsd_stage1(B, highway_fixture, summary):
    """Stage 1 of the same flow (lsd_improvement.py --alg_ver 0 with flat groups): 21 iterations, rank sequence, mask."""
    from oracle import alm_oracle as O
    gold = summary["highway_half_lsd_flat"]
    frames = highway_fixture["frames"]
    h, w, t = frames.shape
    D, _x, mean = O.normalize_and_center(frames)
    assert abs(mean - gold["mean"]) <= 1e-12
    dec = B.lsd_decomposition(D, groups=B.get_proximal_flat_groups_nonoverlap((h, w), (3, 3)), img_shape=(h, w))
    st, log = dec.status(), dec.log()
    print("[highway flat] iters", st.iter, "svp", [l['svp'] for l in log], dec.debug_info())
    assert st.iter == gold["iters"] and bool(st.converged) == gold["converged"]
    assert [l['svp'] for l in log] == gold["svp"]
    assert np.allclose([l['err'] for l in log], gold["err"], rtol=2e-3)
    L, S = dec.download('L'), dec.download('S')
    assert abs(np.linalg.norm(L) - gold["normL"]) <= 1e-5 * gold["normL"]
    assert abs(np.linalg.norm(S) - gold["normS"]) <= 1e-5 * gold["normS"]
    mk = dec.mask(2)
    ref = np.unpackbits(highway_fixture["lsd_mask"])[:D.size].reshape(D.shape, order='F').astype(bool)
    assert (mk == ref).mean() >= 0.999 and abs(float(mk.mean()) - gold["mask_fraction"]) <= 1e-4


def test_watersurface_norms_and_rank_sequence(B, watersurface_u8, summary):
    """init block (row a2) and rank logic (row a5): ||D||_2, ||D||_F, the induced inf-norm (max row sum) and the whole
    svp / err sequence against what the reference printed (golden_summary.json)."""
    from oracle import alm_oracle as O
    D, _x, _mean = O.normalize_and_center(watersurface_u8)
    for key, delta in (("watersurface_flat_delta10", 10), ("watersurface_flat_delta1", 1)):
        gold = summary[key]
        dec = B.lsd_decomposition(D, groups=B.get_proximal_flat_groups_nonoverlap((128, 160), (3, 3)), delta=delta)
        st, log = dec.status(), dec.log()
        assert abs(st.norm_two - gold["norm_two"]) <= 1e-6 * gold["norm_two"]          # D is stored in fp32
        assert abs(st.norm_fro - gold["norm_fro"]) <= 1e-6 * gold["norm_fro"]
        assert abs(st.norm_rowsum - gold["norm_inf_rowsum"]) <= 1e-6 * gold["norm_inf_rowsum"]
        assert st.iter == gold["iters"] and bool(st.converged) == gold["converged"]
        assert [l['svp'] for l in log] == gold["svp"]
        assert np.allclose([l['err'] for l in log], gold["err"], rtol=2e-3)
        mk = dec.mask(2)
        assert abs(float(mk.mean()) - gold["mask_fraction"]) <= 1e-4


def test_whole_solve_n300_headline_kernels(B):
    """The kernel combination of the headline bench (256 < n <= 384: gram_i8_c3 + shrink_stream<8,48,28> + 8-CTA
    eigensolver + the warm-started subspace path) against the CPU oracle on a 144x240x300 clip of the bench generator."""
    from oracle import alm_oracle as O
    rows, cols, n = 144, 240, 300
    D, gt = _synthetic_D(rows, cols, n, seed=0, n_rect=6)
    groups = B.get_proximal_flat_groups_nonoverlap((rows, cols), (3, 3))
    dec = B.lsd_decomposition(D, groups=groups, img_shape=(rows, cols))
    info = dec.debug_info()
    st, log = dec.status(), dec.log()
    L, S = dec.download('L'), dec.download('S')
    olog = []
    Lr, Sr, itr, convr = O.inexact_alm_lsd(D, groups=groups, log=olog)
    print("[n300]", info, "iters", st.iter, itr, "relF(L) %.3e relF(S) %.3e" % (rel_fro(L, Lr), rel_fro(S, Sr)))
    print("[n300] svp gpu", [l['svp'] for l in log], "oracle", [l['svp'] for l in olog])
    assert info["use_stream"] == 1 and info["use_i8"] == 1 and info["stream_R"] == 48 and info["stream_FC"] == 28
    assert info["eig_cluster"] == 8
    assert st.iter == itr and bool(st.converged) == convr
    assert [l['svp'] for l in log] == [l['svp'] for l in olog]
    assert [l['sv'] for l in log] == [l['sv'] for l in olog]
    assert np.allclose([l['err'] for l in log], [l['err'] for l in olog], rtol=1e-3)
    assert rel_fro(L, Lr) <= TOL_F and rel_fro(S, Sr) <= TOL_F
    mk, mref = dec.mask(2), O.foreground_mask(D, Lr, Sr)
    assert (mk == mref).mean() >= 0.999


@pytest.mark.parametrize("rows,cols,n", [(48, 60, 600), (96, 63, 130), (60, 45, 200)])
def test_whole_solve_other_frame_counts_vs_oracle(B, rows, cols, n):
    """n = 600 (5 frame blocks, 16-CTA eigensolver, rank cap of long clips), n = 130 (2 blocks), n = 200 (config 5)."""
    from oracle import alm_oracle as O
    D, _gt = _synthetic_D(rows, cols, n, seed=33, n_rect=2)
    groups = B.get_proximal_flat_groups_nonoverlap((rows, cols), (3, 3))
    dec = B.lsd_decomposition(D, groups=groups, img_shape=(rows, cols))
    st, log = dec.status(), dec.log()
    L, S = dec.download('L'), dec.download('S')
    olog = []
    Lr, Sr, itr, convr = O.inexact_alm_lsd(D, groups=groups, log=olog)
    print("[n=%d] iters %d/%d relF(L) %.3e relF(S) %.3e" % (n, st.iter, itr, rel_fro(L, Lr), rel_fro(S, Sr)), dec.debug_info())
    assert st.iter == itr and bool(st.converged) == convr
    assert [l['svp'] for l in log] == [l['svp'] for l in olog]
    assert rel_fro(L, Lr) <= TOL_F and rel_fro(S, Sr) <= TOL_F
    assert (dec.mask(2) == O.foreground_mask(D, Lr, Sr)).mean() >= 0.999


def test_eig_fast_path_equals_full_path(B):
    """The warm-started subspace eigensolver must reproduce the full tridiagonalisation path: same iteration count, rank
    sequence and (to fp32 storage accuracy) the same L and S; and it must actually be taken on most iterations."""
    rows, cols, n = 96, 120, 300
    D, _gt = _synthetic_D(rows, cols, n, seed=5, n_rect=4)
    groups = B.get_proximal_flat_groups_nonoverlap((rows, cols), (3, 3))
    old = os.environ.get("BSUB_NO_EIG_FAST")
    try:
        os.environ.pop("BSUB_NO_EIG_FAST", None)
        d1 = B.lsd_decomposition(D, groups=groups, img_shape=(rows, cols))
        s1, log1 = d1.status(), d1.log()
        L1, S1 = d1.download('L'), d1.download('S')
        fast = d1.eig_fast_count()
        os.environ["BSUB_NO_EIG_FAST"] = "1"
        d0 = B.lsd_decomposition(D, groups=groups, img_shape=(rows, cols))
        s0, log0 = d0.status(), d0.log()
        L0, S0 = d0.download('L'), d0.download('S')
        assert d0.eig_fast_count() == 0
    finally:
        if old is None:
            os.environ.pop("BSUB_NO_EIG_FAST", None)
        else:
            os.environ["BSUB_NO_EIG_FAST"] = old
    print("[eig fast] iterations on the fast path: %d of %d" % (fast, s1.iter), "relF(L) %.2e relF(S) %.2e" % (rel_fro(L1, L0), rel_fro(S1, S0)))
    assert s1.iter == s0.iter and s1.converged == s0.converged
    assert [l['svp'] for l in log1] == [l['svp'] for l in log0]
    assert rel_fro(L1, L0) <= 2e-6 and rel_fro(S1, S0) <= 2e-5
    assert fast >= s1.iter // 2


def test_generic_groups_first_multiplier(B, watersurface_u8):
    """with a generic partition the solver switches to the two-phase (SPILL) shrink after
    bsub_create; Y1 must still be Y0 + mu Z with Y0 = D / dual_norm (inexact_alm_lsd.py:108-112, 162-163)."""
    from oracle import alm_oracle as O
    D, _x, _mean = O.normalize_and_center(watersurface_u8[:40, :50, :12])
    groups = O.flat_groups_nonoverlap((40, 50), (4, 2))
    m, n = D.shape
    lam = 1.0 / (np.sqrt(max(m, n)) * 10)
    norm_two = np.linalg.norm(D, 2)
    dual = max(norm_two, np.linalg.norm(D, np.inf) / lam)
    mu = 12.5 / norm_two
    Y0 = D / dual
    W = D + Y0 / mu
    u, s, vh = np.linalg.svd(W, full_matrices=False)
    svp, _sv = O.rank_logic(s[:10], 10, mu, min(m, n))
    Lr = (u[:, :svp] * (s[:svp] - 1 / mu)) @ vh[:svp, :]
    Sr = O.prox_flat(D - Lr + Y0 / mu, lam / mu, groups)
    Y1 = Y0 + mu * (D - Lr - Sr)
    dec = B.lsd_decomposition(D, groups=groups, max_iter=1)
    Y = dec.download('Y')
    print("[generic Y1] relF(Y) %.3e  (dropping Y0 would give %.3e)" % (rel_fro(Y, Y1), rel_fro(Y1 - Y0, Y1)))
    assert rel_fro(Y, Y1) <= 1e-5
    assert rel_fro(dec.download('S'), Sr) <= 1e-5


def test_group_sparse_through_sharded_driver(B, highway_fixture, summary):
    """The l2-block mode through dist.ShardedLSD (world size 1: same kernels, the split shrink pass and the block-sum buffer that
    the multi-GPU driver all-reduces): identical to the reference run on the highway fixture."""
    import torch
    from background_subtraction_b200 import dist as bdist
    from oracle import alm_oracle as O
    gold = summary["highway_half_group_sparse"]
    frames = highway_fixture["frames"]
    h, w, t = frames.shape
    D, _x, _mean = O.normalize_and_center(frames)
    labels, ptr, lam = highway_fixture["labels"], highway_fixture["lam_ptr"], highway_fixture["lam"]
    s = bdist.CudaStepSolver(h, w, t, h * w, blocks=(labels, ptr, np.append(lam, 0.0)))
    s.load(np.ascontiguousarray(D.T, dtype=np.float32))

    class Solo:
        world, rank = 1, 0

        def all_reduce_sum(self, t_):
            pass

        def all_reduce_max(self, t_):
            pass
    assert s.block_sums_view() is not None and s.block_sums_view().numel() == t * (int(np.diff(ptr).max()) + 1)
    drv = bdist.ShardedLSD(s, Solo())
    drv.solve()
    drv.finish(2.0, want_mask=False)
    torch.cuda.synchronize()
    st, log = s.status(), s.dec.log()
    assert st.iter == gold["iters"] and bool(st.converged) == gold["converged"]
    assert [l['svp'] for l in log] == gold["svp"]
    L = s.dec.download('L')
    assert abs(np.linalg.norm(L) - gold["normL"]) <= 1e-5 * gold["normL"]


def test_graph_mode_through_sharded_driver(B, watersurface_u8):
    """The overlapping-window LSD through dist.ShardedLSD (world size 1: the split shrink pass, bsub_step_prox_buffers /
    bsub_step_prox_frames -- the pieces the multi-GPU driver puts an all-to-all between): same result as bsub_run."""
    import torch
    from background_subtraction_b200 import dist as bdist
    from oracle import alm_oracle as O
    h, w, t = 64, 60, 12
    D, _x, _mean = O.normalize_and_center(watersurface_u8[:h, :w, :t])
    graph = B.getGraphSPAMS_all_groups((h, w), (3, 3))
    dec = B.lsd_decomposition(D, graphs=graph)
    st0, log0 = dec.status(), dec.log()
    S0 = dec.download('S')
    s = bdist.CudaStepSolver(h, w, t, h * w, graph_cols=w)
    s.load(np.ascontiguousarray(D.T, dtype=np.float32))

    class Solo:
        world, rank = 1, 0

        def all_reduce_sum(self, t_):
            pass

        def all_reduce_max(self, t_):
            pass
    drv = bdist.ShardedLSD(s, Solo())
    drv.solve()
    drv.finish(2.0, want_mask=False)
    torch.cuda.synchronize()
    st, log = s.status(), s.dec.log()
    assert st.iter == st0.iter and bool(st.converged) and bool(st0.converged)
    assert [l['svp'] for l in log] == [l['svp'] for l in log0]
    assert rel_fro(s.dec.download('S'), S0) <= 1e-6
