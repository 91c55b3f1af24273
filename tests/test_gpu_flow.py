"""-m gpu: the stages on either side of the decomposition (csrc/post.cu through flow.py) against the
CPU oracle (oracle/flow_oracle.py, pinned to OpenCV / SciPy / the reference in tests/test_oracle_flow.py) on seeded inputs."""
import numpy as np
import pytest

from conftest import crop_D, rel_fro
from flow_cases import blob_video, random_masks

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def B():
    import background_subtraction_b200 as B
    return B


@pytest.fixture(scope="module")
def F():
    from oracle import flow_oracle as F
    return F


@pytest.mark.parametrize("h,w,t,ratio", [(48, 64, 5, 0.5), (37, 53, 3, 0.5), (60, 45, 4, 1 / 3), (50, 70, 2, 0.37), (128, 160, 6, 0.25),
                                         (33, 47, 3, 0.8), (20, 30, 2, 2.0), (17, 23, 3, 1.5), (1080, 1920, 2, 1 / 4)])
def test_resize_with_cv2(B, F, h, w, t, ratio):
    rng = np.random.default_rng(h + w)
    cube = rng.random((h, w, t))
    got = B.resize_with_cv2(cube, ratio)
    ref = F.resize_with_cv2(cube, ratio)
    assert got.shape == ref.shape and got.dtype == np.float64
    assert np.abs(got - ref).max() <= 2e-6                                 # float32 on the device, values in [0, 1)
    first = B.resize_with_cv2_by_first_axis(np.ascontiguousarray(cube.transpose(2, 0, 1)), ratio)
    assert np.abs(first - ref.transpose(2, 0, 1)).max() <= 2e-6
    try:
        import cv2
    except ImportError:
        return
    size = [int(np.ceil(h * ratio)), int(np.ceil(w * ratio))]
    cvr = cv2.resize(cube[:, :, 0], size[::-1], interpolation=cv2.INTER_AREA if ratio < 1 else cv2.INTER_CUBIC)
    assert np.abs(got[:, :, 0] - cvr).max() <= 5e-6


@pytest.mark.parametrize("h,w,t", [(36, 44, 5), (7, 5, 3), (128, 160, 4), (241, 322, 2)])
def test_connected_components_and_filter(B, F, h, w, t):
    m = random_masks(h * w, h, w, t, density=(0.25, 0.55))
    cc = B.connected_components(m)
    lab = cc.labels.cpu().numpy().reshape(t, w, h).transpose(2, 1, 0)      # [h, w, t]
    for f in range(t):
        n, ref_lab, ref_st = F.connected_components(m[:, :, f])
        assert cc.num[f] == n - 1
        st, idx = cc.stats_cv2(f)
        assert np.array_equal(st, ref_st[1:])                               # OpenCV's numbering and stats columns
        # device label l (numbered by smallest pixel index) <-> OpenCV label: same partition
        order = np.empty(len(idx) + 1, dtype=np.int64)
        order[0] = 0
        order[idx - cc.offsets[f] + 1] = np.arange(1, len(idx) + 1)
        assert np.array_equal(order[lab[:, :, f]], ref_lab)
    for thresh in (None, 2, 15):
        got = B.filter_sparse_map(m, thresh)
        assert got.dtype == m.dtype and np.array_equal(got, F.filter_sparse_map(m, thresh))
    assert not B.filter_sparse_map(np.zeros((h, w, t), dtype=bool)).any()


@pytest.mark.parametrize("W,H,T", [(40, 50, 12), (30, 47, 9), (52, 61, 7), (21, 25, 3), (320, 240, 20)])
def test_computeSCube(B, F, W, H, T):
    rng = np.random.default_rng(W + H)
    xt, yt = rng.standard_normal((W, H, T)), rng.standard_normal((H, W, T))
    xt[np.abs(xt) < 0.8] = 0                                                # sparse, like the stage-2 output
    got = B.computeSCube(xt, yt)
    ref = F.compute_scube_separable(xt, yt)
    assert got.shape == (T, H, W)
    assert np.abs(got - ref).max() <= 2e-6 * np.abs(ref).max()
    if W * H * T <= 40 * 50 * 12:
        assert np.abs(got - F.compute_scube_dense(xt, yt)).max() <= 2e-6 * np.abs(ref).max()   # the reference's dense scipy call


@pytest.mark.parametrize("seed,h,w,t", [(1, 60, 80, 6), (2, 60, 80, 6), (3, 96, 128, 10), (4, 240, 320, 4)])
def test_run_motion_saliency_check(B, F, seed, h, w, t):
    mask, cube = blob_video(seed, h, w, t)
    data = np.zeros(mask.shape)
    gb, wb = B.run_motion_saliency_check(data, mask, cube)
    gr, wr = F.run_motion_saliency_check(data, mask, cube)
    assert sum(len(g) for g in gr) > 0 and len(gb) == t
    for f in range(t):
        assert len(gb[f]) == len(gr[f])
        for g1, g2, w1, w2 in zip(gb[f], gr[f], wb[f], wr[f]):
            assert np.array_equal(g1, g2) and abs(w1 - w2) <= 1e-5 * abs(w2)
    # the device form feeds the group-sparse solver directly
    labels, ptr, lam = B.motion_saliency_blocks(mask.shape, mask, cube)
    assert labels.shape == (t, h * w) and ptr[-1] == sum(len(g) for g in gr) and len(lam) == ptr[-1] + 1


@pytest.mark.parametrize("h,w,t,pct", [(30, 26, 2, 0.2), (128, 160, 3, 0.05), (240, 320, 2, 0.05), (64, 48, 2, 0.0)])
def test_apply_morph_ops(B, F, h, w, t, pct):
    m = random_masks(h, h, w, t, density=(0.005, 0.03))
    got = B.apply_morph_ops(m, percetage=pct)
    assert got.dtype == bool and np.array_equal(got, F.apply_morph_ops(m, pct))
    with pytest.raises(Exception, match="disk"):
        B.apply_morph_ops(m, footprint_name='diamond')


def test_LSD_improved_two_pass(B, F, watersurface_u8):
    """LSD_improved (lsd_improvement.py:441-487), alg_ver 2: flat LSD -> mask -> morphology -> weight map -> LSD with background."""
    from oracle import alm_oracle as O
    cube = np.asfortranarray(watersurface_u8[40:88, 50:110, 0:16].astype(np.float64))
    D, _x, _mean = O.normalize_and_center(cube)
    shape = cube.shape
    wm_ref, it1_ref, _c = F.improved_LSD_weight_mask(D, shape)
    wm, it1, _c2 = B.improved_LSD_weight_mask(D, shape, (1, 1.5), delta=1.0,
                                              proximal_object=B.get_proximal_flat_groups_nonoverlap(shape[:2], (3, 3)),
                                              mode="NONOVERLAPPING_GROUPS")
    assert abs(it1 - it1_ref) <= 1 and (wm == wm_ref).mean() >= 0.999
    out = B.LSD_improved(cube.copy(order='F'), 0, 15, 1, alg_ver=2)
    S, S_mask, L_recon, _im, _mean2, shp, it, conv, git, gconv = out
    assert shp == shape and git == it1 and S_mask.shape == shape and conv
    L_ref, S_ref, it_ref, conv_ref, mask_ref = F.LSD_improved_from_weight_mask(D, shape, wm)    # same weight map on both sides
    assert abs(it - it_ref) <= 1 and conv == conv_ref
    assert rel_fro(L_recon.reshape(D.shape, order='F'), L_ref) <= 1e-4 and rel_fro(S, S_ref) <= 1e-4
    assert (S_mask == mask_ref).mean() >= 0.999
    with pytest.raises(Exception, match="alg ver"):
        B.LSD_improved(cube.copy(order='F'), 0, 15, 1, alg_ver=3)


def test_motion_saliency_golden_highway(B, highway_fixture):
    """Device run_motion_saliency_check on the inputs of the reference run that produced the committed labels / lambdas."""
    from test_oracle_flow import highway_saliency_inputs
    xc, mask1, sal = highway_saliency_inputs(highway_fixture)
    labels, ptr, lam = B.motion_saliency_blocks(xc.shape, mask1, sal)
    assert np.array_equal(ptr, highway_fixture["lam_ptr"])
    assert np.array_equal(labels.cpu().numpy(), highway_fixture["labels"])
    assert np.allclose(lam[:-1], highway_fixture["lam"], rtol=2e-6, atol=0)        # float32 saliency cube on the device
    gb, wb = B.run_motion_saliency_check(xc, mask1, sal)
    assert [len(g) for g in gb] == np.diff(ptr).tolist() and gb[0][0].dtype == bool if len(gb[0]) else True


@pytest.mark.parametrize("h,w,t", [(48, 64, 40), (60, 45, 33), (240, 320, 200)])
def test_saliency_rpca_batch(B, F, h, w, t):
    """compute_RPCA / executeSaliencyRPCA (computeRPCADecomposition.py:12-95): every X-T and Y-T slice through the rank-1-capped
    inexact_alm_rpca, one cluster per slice, against the float64 oracle with an exact SVD."""
    from test_oracle_flow import saliency_slices
    from background_subtraction_b200 import flow
    video = saliency_slices(7, h, w, t)                                      # [t, h, w], 0..255
    xt = np.ascontiguousarray(video.transpose(2, 1, 0))                      # [w, h, t]
    nsl = min(xt.shape[0], 12)
    tol_l1 = xt.shape[1] * xt.shape[2] * 1e-4
    L, S, info = flow.inexact_alm_rpca_batch(xt, delta=1.0, tol_l1=tol_l1, return_info=True)
    assert L.shape == xt.shape and np.all(info["iters"] > 0) and np.all(info["rank"] == 1)
    sel = np.linspace(0, xt.shape[0] - 1, nsl).astype(int)
    Lr, Sr, its = F.compute_RPCA(xt[sel], tol_l1)
    print("[saliency batch] iters gpu", info["iters"][sel].tolist(), "oracle", its.tolist())
    assert np.all(np.abs(info["iters"][sel] - its) <= 1)
    for k, i in enumerate(sel):
        same_iters = info["iters"][i] == its[k]
        assert rel_fro(L[i], Lr[k]) <= 1e-4
        if same_iters:
            assert rel_fro(S[i], Sr[k]) <= 1e-4
        assert np.sum(np.abs(xt[i] - L[i] - S[i])) <= 2 * tol_l1            # float32 L + S against the float64 input
    # the drop-in pair and the chain into computeSCube
    Lc, Sc = B.compute_RPCA(xt[sel], True, tol_l1)
    assert np.array_equal(Lc, L[sel]) and np.array_equal(Sc, S[sel])
    if h * w * t <= 60 * 45 * 40:
        xl, xs, yl, ys = B.executeSaliencyRPCA(video, 1)
        rxl, rxs, ryl, rys = F.executeSaliencyRPCA(video, 1)
        assert xs.shape == (w, h, t) and ys.shape == (h, w, t)
        cube = B.computeSCube(xs, ys)
        ref = F.compute_scube_separable(rxs, rys)
        assert np.abs(cube - ref).max() <= 2e-3 * np.abs(ref).max()
    with pytest.raises(Exception, match="rank"):
        flow.inexact_alm_rpca_batch(xt[:2], max_rank=2)
