"""-m "not gpu": pins oracle/flow_oracle.py (the CPU restatements of the stages around the decomposition, SURVEY.md 8f rows 1,
3, 4) against OpenCV / SciPy themselves and -- when /root/reference is present (build container only) -- against the
unmodified reference functions (filter_sparse_map, computeSCube, run_motion_saliency_check, merge_masks)."""
import importlib
import sys

import numpy as np
import pytest

from flow_cases import blob_video, random_masks
from oracle import flow_oracle as F
from oracle import ref_harness as R

cv2 = pytest.importorskip("cv2")


@pytest.mark.parametrize("h,w,ratio", [(48, 64, 0.5), (37, 53, 0.5), (60, 45, 1 / 3), (50, 70, 0.37), (128, 160, 0.25), (33, 47, 0.8),
                                       (20, 30, 2.0), (17, 23, 1.5)])
def test_resize_restatement_matches_cv2(h, w, ratio):
    rng = np.random.default_rng(h * w)
    cube = rng.random((h, w, 3))
    got = F.resize_with_cv2(cube, ratio)
    size = [int(np.ceil(h * ratio)), int(np.ceil(w * ratio))]
    for t in range(3):
        ref = cv2.resize(cube[:, :, t], size[::-1], interpolation=cv2.INTER_AREA if ratio < 1 else cv2.INTER_CUBIC)
        assert got[:, :, t].shape == ref.shape
        assert np.abs(got[:, :, t] - ref).max() <= (1e-7 if ratio < 1 else 3e-6)     # OpenCV keeps its tap weights in float32


def test_connected_components_numbering_and_stats_match_cv2():
    rng = np.random.default_rng(5)
    for _ in range(150):
        h, w = int(rng.integers(3, 40)), int(rng.integers(3, 40))
        fr = (rng.random((h, w)) < rng.uniform(0.15, 0.7)).astype(np.uint8) * 255
        n, lab, st, _c = cv2.connectedComponentsWithStats(fr, 8, cv2.CV_32S)
        n2, lab2, st2 = F.connected_components(fr)
        assert n == n2 and np.array_equal(lab, lab2) and np.array_equal(st, st2)


def test_filter_sparse_map_restatement():
    m = random_masks(3, 36, 44, 5, density=(0.3, 0.5))
    for thresh in (None, 3, 20):
        got = F.filter_sparse_map(m, thresh)
        # same thing through OpenCV, as the reference does it (utils.py:404-420)
        th = (36 * 44) // 200 if thresh is None else thresh
        ref = np.zeros_like(m)
        for i in range(m.shape[2]):
            n, lab, st, _c = cv2.connectedComponentsWithStats(m[:, :, i].astype(np.uint8) * 255, 8, cv2.CV_32S)
            for j in range(1, n):
                if st[j, cv2.CC_STAT_AREA] > th:
                    ref[:, :, i][lab == j] = True
        assert np.array_equal(got, ref)
        if R.available():
            with R.quiet():
                assert np.array_equal(got, R.load()["utils"].filter_sparse_map(m, thresh))


@pytest.mark.parametrize("W,H,T", [(40, 50, 12), (30, 47, 9), (52, 61, 7), (21, 25, 3)])
def test_scube_separable_equals_dense_scipy(W, H, T):
    rng = np.random.default_rng(W + H)
    xt, yt = rng.standard_normal((W, H, T)), rng.standard_normal((H, W, T))
    dense = F.compute_scube_dense(xt, yt)
    sep = F.compute_scube_separable(xt, yt)
    assert np.abs(dense - sep).max() <= 1e-13 * np.abs(dense).max()
    k = int(min(H, W) / 10)
    assert np.allclose(F.gkern(k), np.einsum('i,j,k->ijk', F.gauss_taps(k), F.gauss_taps(k), F.gauss_taps(k)), rtol=1e-13, atol=0)


@pytest.mark.skipif(not R.available(), reason="reference not present")
def test_scube_against_reference():
    R.load()
    sys.path.insert(0, R.REF_DIR)
    try:
        ref_mod = importlib.import_module("computeSCube")
    finally:
        sys.path.remove(R.REF_DIR)
    rng = np.random.default_rng(9)
    xt, yt = rng.standard_normal((40, 30, 6)), rng.standard_normal((30, 40, 6))
    with R.quiet():
        ref = ref_mod.computeSCube(xt, yt)
    assert np.abs(ref - F.compute_scube_separable(xt, yt)).max() <= 1e-13 * np.abs(ref).max()


def _same_groups(a, b, wa, wb):
    assert len(a) == len(b)
    for f in range(len(a)):
        assert len(a[f]) == len(b[f]) and len(wa[f]) == len(wb[f])
        for g1, g2, w1, w2 in zip(a[f], b[f], wa[f], wb[f]):
            assert np.array_equal(g1, g2)
            assert abs(w1 - w2) <= 1e-12 * abs(w2)


@pytest.mark.skipif(not R.available(), reason="reference not present")
def test_motion_saliency_against_reference():
    ref = R.load()["motion_saliency_check"]
    for seed in (1, 2, 3):
        mask, cube = blob_video(seed, 60, 80, 6)
        data = np.zeros(mask.shape)
        with R.quiet():
            gb, wb = ref.run_motion_saliency_check(data, mask, cube)
        gb2, wb2 = F.run_motion_saliency_check(data, mask, cube)
        assert sum(len(g) for g in gb) > 0
        _same_groups(gb2, gb, wb2, wb)


def test_morphology_equals_definition():
    """Dilation / erosion with out-of-image pixels ignored (what the device kernel computes) == scipy 'reflect' for a disk."""
    m = random_masks(7, 30, 26, 2, density=(0.02, 0.08))
    r = F.disk_radius(0.2, 30)
    assert r == 3
    fp = F.disk(r).astype(bool)
    def morph(x, erode):
        out = np.zeros_like(x)
        h, w = x.shape
        for i in range(h):
            for j in range(w):
                vals = [x[i + di, j + dj] for di in range(-r, r + 1) for dj in range(-r, r + 1)
                        if fp[di + r, dj + r] and 0 <= i + di < h and 0 <= j + dj < w]
                out[i, j] = min(vals) if erode else max(vals)
        return out
    got = F.apply_morph_ops(m, 0.2)
    for f in range(2):
        x = morph(morph(morph(m[:, :, f], False), False), True)
        assert np.array_equal(got[:, :, f], x)
    assert F.disk(2).tolist() == [[0, 0, 1, 0, 0], [0, 1, 1, 1, 0], [1, 1, 1, 1, 1], [0, 1, 1, 1, 0], [0, 0, 1, 0, 0]]


@pytest.mark.skipif(not R.available(), reason="reference not present")
def test_merge_masks_against_reference():
    ref = R.load()["lsd_improvement"]
    a, b = random_masks(1, 10, 12, 3)[..., 0], random_masks(2, 10, 12, 3)[..., 0]
    assert np.array_equal(ref.merge_masks((a, a | b), (1, 1.5)), F.merge_masks((a, a | b), (1, 1.5)))


def highway_saliency_inputs(fx):
    """Inputs of the run_motion_saliency_check call that produced the fixture's labels (tests/golden/make_golden.py)."""
    frames = fx["frames"]
    h, w, t = frames.shape
    x = np.asfortranarray(frames.astype(np.float64))
    x -= np.min(x)
    x *= 1.0 / np.max(x)
    sal = np.abs(x - np.median(x, axis=2, keepdims=True))
    sal /= sal.sum()
    mask1 = np.unpackbits(fx["lsd_mask"])[:h * w * t].reshape((h * w, t), order='F').reshape((h, w, t), order='F').astype(bool)
    return x - np.mean(x), mask1, sal


def test_motion_saliency_golden_highway(highway_fixture):
    """The labels / lambdas committed with the fixture came from the reference's own run_motion_saliency_check."""
    xc, mask1, sal = highway_saliency_inputs(highway_fixture)
    gb, wb = F.run_motion_saliency_check(xc, mask1, sal)
    labels, ptr, lam = highway_fixture["labels"], highway_fixture["lam_ptr"], highway_fixture["lam"]
    assert [len(g) for g in gb] == np.diff(ptr).tolist()
    for f in range(len(gb)):
        for b, g in enumerate(gb[f]):
            assert np.array_equal(g, labels[f] == b + 1)
            assert abs(wb[f][b] - lam[ptr[f] + b]) <= 1e-12 * lam[ptr[f] + b]


def saliency_slices(seed=3, h=48, w=64, t=40):
    """[t, h, w] float64 video (0..255, like import_video_as_frames) from the seeded synthetic generator."""
    from background_subtraction_b200 import synth
    video, _ = synth.make_clip(h, w, t, seed=seed, n_rect=2)
    return video.reshape(t, w, h).transpose(0, 2, 1).astype(np.float64)


def test_rank_capped_rpca_restatement():
    """The batch engine's oracle: without the cap it IS inexact_alm_rpca (oracle port, and the reference when present)."""
    from oracle import alm_oracle as O
    xt = saliency_slices().transpose(2, 1, 0)
    D = np.asfortranarray(xt[5] - xt[5].mean())
    L, S, it, conv = F.inexact_alm_rpca_capped(D, 1.0)
    Lo, So, ito, convo = O.inexact_alm_rpca(D, 1.0)
    assert (it, conv) == (ito, convo) and np.array_equal(L, Lo) and np.array_equal(S, So)
    if R.available():
        with R.quiet():
            Lr, Sr, itr, convr = R.load()["lsd_improvement"].inexact_alm_rpca(D, 1.0)
        assert (it, conv) == (itr, convr) and np.allclose(L, Lr, atol=1e-12) and np.allclose(S, Sr, atol=1e-12)
    L1, S1, it1, conv1 = F.inexact_alm_rpca_capped(xt[5], 1.0, max_rank=1, tol_l1=xt.shape[1] * xt.shape[2] * 1e-4)
    assert conv1 and np.linalg.matrix_rank(L1) == 1 and np.sum(np.abs(xt[5] - L1 - S1)) <= xt.shape[1] * xt.shape[2] * 1e-4
