"""oracle/flow_oracle.py -- TEST INFRASTRUCTURE ONLY (never imported by the product path).

CPU restatements (NumPy / SciPy, float64) of the stages on either side of the decomposition.
Each function cites the reference lines it follows; tests/test_oracle.py pins them against the unmodified reference
functions (oracle/ref_harness.py) and against OpenCV / SciPy themselves where the reference only wraps those.
"""
import math

import numpy as np
from scipy import ndimage


# ---------------------------------------------------------------- resize (utils.py:119-136 -> cv2.resize)
def _area_tab(ssize, dsize):
    """OpenCV computeResizeAreaTab: list over destination cells of (source index, weight)."""
    scale = 1.0 / (dsize / ssize)
    tab = []
    for d in range(dsize):
        fsx1 = d * scale
        fsx2 = fsx1 + scale
        cell = min(scale, ssize - fsx1)
        sx1, sx2 = math.ceil(fsx1), math.floor(fsx2)
        sx2 = min(sx2, ssize - 1)
        sx1 = min(sx1, sx2)
        taps = []
        if sx1 - fsx1 > 1e-3:
            taps.append((sx1 - 1, (sx1 - fsx1) / cell))
        for sx in range(sx1, sx2):
            taps.append((sx, 1.0 / cell))
        if fsx2 - sx2 > 1e-3:
            taps.append((sx2, min(min(fsx2 - sx2, 1.0), cell) / cell))
        tab.append(taps)
    return tab


def _cubic_tab(ssize, dsize):
    """OpenCV INTER_CUBIC (A = -0.75), BORDER_REPLICATE."""
    scale, A = 1.0 / (dsize / ssize), -0.75
    tab = []
    for d in range(dsize):
        fx = (d + 0.5) * scale - 0.5
        sx = math.floor(fx)
        fx = float(np.float32(fx - sx))
        x0, x1, x2 = fx + 1.0, fx, 1.0 - fx
        c = [((A * x0 - 5 * A) * x0 + 8 * A) * x0 - 4 * A, ((A + 2) * x1 - (A + 3)) * x1 * x1 + 1, ((A + 2) * x2 - (A + 3)) * x2 * x2 + 1]
        c.append(1.0 - sum(c))
        tab.append([(min(max(sx - 1 + k, 0), ssize - 1), c[k]) for k in range(4)])
    return tab


def _tab_matrix(tab, ssize):
    M = np.zeros((len(tab), ssize))
    for d, taps in enumerate(tab):
        for s, w in taps:
            M[d, s] += w
    return M


def resize_image(img, dst_h, dst_w, cubic=False):
    """cv2.resize(img, (dst_w, dst_h), interpolation=INTER_AREA | INTER_CUBIC) for one 2-D float image."""
    tab = _cubic_tab if cubic else _area_tab
    My = _tab_matrix(tab(img.shape[0], dst_h), img.shape[0])
    Mx = _tab_matrix(tab(img.shape[1], dst_w), img.shape[1])
    return My @ np.asarray(img, dtype=np.float64) @ Mx.T


def resize_with_cv2(images, ratio):
    """/root/reference/utils.py:129-136 ([h, w, T] cube)."""
    size = [int(np.ceil(images.shape[i] * ratio)) for i in (0, 1)]
    out = np.empty(size + [images.shape[2]])
    for t in range(images.shape[2]):
        out[:, :, t] = resize_image(images[:, :, t], size[0], size[1], cubic=not (ratio < 1))
    return out


# ---------------------------------------------------------------- connected components (cv2.connectedComponentsWithStats, 8-way)
_EIGHT = np.ones((3, 3), dtype=bool)


def connected_components(frame):
    """Labels of one binary 2-D frame numbered like OpenCV (raster order of the first 2x2 block of each component) and the
    stats columns (left, top, width, height, area) -- cv2.connectedComponentsWithStats(frame, 8, cv2.CV_32S)."""
    lab, k = ndimage.label(np.asarray(frame) != 0, structure=_EIGHT)
    if k == 0:
        return 1, lab.astype(np.int32), np.zeros((1, 5), dtype=np.int64)
    h, w = lab.shape
    ii, jj = np.nonzero(lab)
    key = (ii // 2) * ((w + 1) // 2) + (jj // 2)
    first = np.full(k + 1, np.iinfo(np.int64).max)
    np.minimum.at(first, lab[ii, jj], key)
    order = np.argsort(first[1:], kind="stable")          # component ids (0-based) in OpenCV order
    new_id = np.empty(k + 1, dtype=np.int32)
    new_id[0] = 0
    new_id[order + 1] = np.arange(1, k + 1)
    lab = new_id[lab]
    stats = np.zeros((k + 1, 5), dtype=np.int64)
    for l in range(0, k + 1):                              # row 0 = the background, like OpenCV
        yy, xx = np.nonzero(lab == l)
        if yy.size:
            stats[l] = (xx.min(), yy.min(), xx.max() - xx.min() + 1, yy.max() - yy.min() + 1, yy.size)
    return k + 1, lab, stats


def filter_sparse_map(sparse_array, size_thresh=None):
    """/root/reference/utils.py:404-420 ([h, w, t] boolean cube)."""
    if size_thresh is None:
        size_thresh = (sparse_array.shape[0] * sparse_array.shape[1]) // 200
    out = np.zeros_like(sparse_array)
    for t in range(sparse_array.shape[2]):
        lab, k = ndimage.label(sparse_array[:, :, t] != 0, structure=_EIGHT)
        if k:
            area = np.bincount(lab.ravel(), minlength=k + 1)
            keep = area > size_thresh
            keep[0] = False
            out[:, :, t] = keep[lab]
    return out


# ---------------------------------------------------------------- computeSCube (computeSCube.py:9-50, 82-92)
def gkern(l=10, sig=1.):
    ax = np.linspace(-(l - 1) / 2., (l - 1) / 2., l)
    xx, yy, zz = np.meshgrid(ax, ax, ax)
    kernel = np.exp(-0.5 * (np.square(xx) + np.square(yy) + np.square(zz)) / np.square(sig))
    return kernel / np.sum(kernel)


def scube_product(sparse_xt, sparse_yt):
    """build_sparse_xt_cube, build_sparse_yt_cube, build_final_cube (computeSCube.py:22-50): [t, h, w] cube with unit sum."""
    cube = np.abs(sparse_xt.transpose([2, 1, 0])) * np.abs(sparse_yt.transpose([2, 0, 1]))
    return cube / np.sum(cube)


def compute_scube_dense(sparse_xt, sparse_yt):
    """computeSCube.py:82-92 verbatim in behaviour: the dense l^3-tap scipy convolution (small shapes only: O(l^3) per voxel)."""
    cube = scube_product(sparse_xt, sparse_yt)
    k = int(min(cube.shape[1], cube.shape[2]) / 10)
    return ndimage.convolve(cube, gkern(k), mode='reflect')


def gauss_taps(l, sig=1.):
    """1-D factor of gkern: gkern(l) == outer product of three copies of this vector (exactly separable, incl. normalisation)."""
    ax = np.linspace(-(l - 1) / 2., (l - 1) / 2., l)
    g = np.exp(-0.5 * np.square(ax) / np.square(sig))
    return g / g.sum()


def conv_shift(l):
    """ndimage.convolve(x, w)[i] = sum_k w[l-1-k] x[i + k - shift]: shift = l//2 for odd l, l//2 - 1 for even l (scipy moves the
    origin of an even-sized kernel by one when it flips it)."""
    return l // 2 if l % 2 else l // 2 - 1


def reflect_index(s, n):
    per = 2 * n
    s = np.mod(s, per)
    return np.where(s >= n, per - 1 - s, s)


def conv1d_reflect(a, w, axis):
    """One axis of the separable form (what bsub_conv1d_reflect_dev computes): out[i] = sum_k w[k] a[reflect(i + k - shift)]."""
    l = len(w)
    n = a.shape[axis]
    out = np.zeros_like(a, dtype=np.float64)
    idx = np.arange(n)
    for k in range(l):
        out += w[k] * np.take(a, reflect_index(idx + k - conv_shift(l), n), axis=axis)
    return out


def compute_scube_separable(sparse_xt, sparse_yt):
    cube = scube_product(sparse_xt, sparse_yt)
    k = int(min(cube.shape[1], cube.shape[2]) / 10)
    w = gauss_taps(k)[::-1]                     # convolution flips the kernel (symmetric here)
    for axis in (2, 1, 0):
        cube = conv1d_reflect(cube, w, axis)
    return cube


# ---------------------------------------------------------------- motion saliency check (motion_saliency_check.py:19-120, utils.py:340-401)
def contained_in(cc1, cc2):
    x2, y2, w2, h2 = cc2
    x1, y1, w1, h1 = cc1
    return bool(x2 < x1 and y2 < y1 and x1 + w1 < x2 + w2 and y1 + h1 < y2 + h2)


def nested_relabel_map(stats):
    """unite_nestedCCs (utils.py:351-401) reduced to the label -> new label map it applies: every edge (n1, n2) of the spanning
    tree of the bbox-nesting graph relabels the ORIGINAL label n2 as n1 (no chaining), in networkx's edge order."""
    import networkx as nx
    k = stats.shape[0]
    cc = {i: tuple(int(v) for v in stats[i, :4]) for i in range(1, k)}
    nested = [(l1, l2) for l1 in cc for l2 in cc if l1 != l2 and contained_in(cc[l1], cc[l2])]
    graph = nx.Graph()
    for l1, l2 in nested:
        graph.add_edge(l2, l1)
    remap = {i: i for i in range(1, k)}
    for n1, n2 in nx.minimum_spanning_tree(graph).edges():
        remap[n2] = n1
    return remap


def compute_groups_per_frame(mask_frame, cube_frame, frame_idx):
    """motion_saliency_check.py:19-49: (frame_idx, weight, area, mask_1d F-order) per (nest-merged) component."""
    k, lab, stats = connected_components(mask_frame)
    remap = nested_relabel_map(stats)
    new_lab = lab.copy()
    for old, new in remap.items():
        if new != old:
            new_lab[lab == old] = new
    groups = []
    for l in np.unique(new_lab):
        if l == 0:
            continue
        m2 = new_lab == l
        area = int(m2.sum())
        groups.append((frame_idx, float(np.sum(cube_frame[m2]) / area), area, m2.flatten(order='F')))
    return groups


def filter_groups(groups, size_thresh):
    """motion_saliency_check.py:52-63."""
    w = np.array([g[1] for g in groups])
    thr = np.mean(w) + np.std(w)
    kept = [g for g in groups if g[1] > thr and g[2] > size_thresh]
    return kept, min(g[1] for g in kept)


def run_motion_saliency_check(data, sparse_binary_mat, sparse_cube, delta=10):
    """motion_saliency_check.py:66-120 -> (groups_by_frame, weights_by_frame)."""
    shape = data.shape
    n = shape[2]
    size_thresh = (shape[0] * shape[1]) / 1500
    groups = []
    for f in range(n):
        groups.extend(compute_groups_per_frame(sparse_binary_mat[:, :, f], sparse_cube[:, :, f], f))
    kept, min_w = filter_groups(groups, size_thresh)
    kept.sort(key=lambda g: g[0])
    norm = 1.0 / (delta * np.sqrt(max(shape[0] * shape[1], shape[2]))) * min_w
    gb, wb = [], []
    for f in range(n):
        fg = [g for g in kept if g[0] == f]
        gb.append([g[3] for g in fg])
        wb.append([norm / g[1] for g in fg])
    return gb, wb


# ---------------------------------------------------------------- morphology (lsd_improvement.py:307-335 -> skimage disk / dilation / closing)
def disk(radius):
    """skimage.morphology.disk(radius) (default strict radius): X^2 + Y^2 <= radius^2 on a (2r+1)^2 grid."""
    L = np.arange(-radius, radius + 1)
    X, Y = np.meshgrid(L, L)
    return ((X ** 2 + Y ** 2) <= radius ** 2).astype(np.uint8)


def disk_radius(percentage, im_height):
    """get_footprint('disk', percentage * im_height) (lsd_improvement.py:307-320): disk(ceil(size) // 2)."""
    return int(math.ceil(percentage * im_height)) // 2


def apply_morph_ops(mask_hwt, percentage=0.05):
    """apply_morph_ops (lsd_improvement.py:323-335) with the 'disk' footprint expanded along time ([d, d, 1]): dilation, then
    closing (dilation followed by erosion), every frame on its own.  skimage's dilation / erosion are scipy's grey_dilation /
    grey_erosion with mode='reflect'; for a disk that equals ignoring the pixels outside the image."""
    fp = disk(disk_radius(percentage, mask_hwt.shape[0])).astype(bool)[:, :, None]
    x = np.asarray(mask_hwt).astype(np.uint8)
    x = ndimage.grey_dilation(x, footprint=fp, mode='reflect')
    x = ndimage.grey_dilation(x, footprint=fp, mode='reflect')
    x = ndimage.grey_erosion(x, footprint=fp, mode='reflect')
    return x.astype(bool)


def merge_masks(masks, weights, background_marker=-1):
    """lsd_improvement.py:338-351."""
    merged = np.ones(masks[0].shape) * background_marker
    for i in range(len(masks) - 1, -1, -1):
        merged[masks[i]] = weights[i]
    return merged


# ---------------------------------------------------------------- the two-pass LSD (lsd_improvement.py:369-487)
def improved_LSD_weight_mask(D, shape, weights=(1, 1.5), alg_ver=2):
    """First pass of build_improved_LSD_graphs (lsd_improvement.py:369-404): decomposition, mask, morphology with the
    mask-percentage back-off, merged weight map (-1 = background).  Returns (weight_mask, iterations, converged)."""
    from . import alm_oracle as O
    h, w, _t = shape
    if alg_ver == 2:
        L, S, it, conv = O.inexact_alm_lsd(D, groups=O.flat_groups_nonoverlap((h, w), (3, 3)), delta=1.0)
    else:
        L, S, it, conv = O.inexact_alm_rpca(D, delta=10.0)
    S_mask = O.foreground_mask(D, L, S, 2).reshape(shape, order='F')
    ratio, total, cur = 0.05, 5, 1
    morph = apply_morph_ops(S_mask, ratio)
    wm = merge_masks((S_mask, morph), weights)
    pct = np.sum(wm > 0) / wm.size * 100
    while pct > 20 and cur < total:
        ratio -= 0.01
        total += 1
        if ratio * h <= 0:
            break
        morph = apply_morph_ops(S_mask, ratio)
        wm = merge_masks((S_mask, morph), weights)
        pct = np.sum(wm > 0) / wm.size * 100
    return wm, it, conv


def LSD_improved_from_weight_mask(D, shape, wm):
    """Second pass (lsd_improvement.py:406-436, 477-483): per-frame centre-window graphs + background masks from the weight map,
    inexact_alm_lsd_with_background, mask."""
    from . import alm_oracle as O
    h, w, t = shape
    graphs = [O.graph_group_centers((h, w), 1, wm[:, :, i]) for i in range(t)]
    bg = [(wm[:, :, i] < 0).flatten(order='F') for i in range(t)]
    L, S, it, conv = O.inexact_alm_lsd_with_background(D, graphs, bg)
    return L, S, it, conv, O.foreground_mask(D, L, S, 2).reshape(shape, order='F')


# ---------------------------------------------------------------- stage-2 saliency RPCA (computeRPCADecomposition.py:12-95)
def inexact_alm_rpca_capped(D0, delta=1.0, max_rank=None, tol_out=1e-7, tol_l1=0.0, max_iter=500, rho=1.2):
    """inexact_alm_rpca (lsd_improvement.py:123-196, use_sv_prediction=False) with an optional cap on the rank of L and an optional
    second stopping test sum|Z| <= tol_l1 (the test RobustPCA(tol=...) applies in compute_RPCA).  max_rank=None, tol_l1=0 is the
    reference function itself (pinned in tests/test_oracle_flow.py).  Returns (L, S, iterations, converged)."""
    D = np.asarray(D0, dtype=np.float64)
    m, n = D.shape
    lam = (np.sqrt(max(m, n)) * delta) ** (-1)
    norm_two = np.linalg.norm(D, ord=2)
    norm_inf = np.linalg.norm(D, ord=np.inf) / lam
    Y = D / max(norm_two, norm_inf)
    mu = 1.25 / norm_two
    L = np.zeros_like(D)
    S = np.zeros_like(D)
    for it in range(1, max_iter + 1):
        u, s, vh = np.linalg.svd(D - S + Y / mu, full_matrices=False)
        svp = int(np.sum(s - 1 / mu > 0))
        if max_rank is not None:
            svp = min(svp, max_rank)
        L = (u[:, :svp] * (s[:svp] - 1 / mu)) @ vh[:svp, :]
        G = D - L + Y / mu
        S = np.maximum(G - lam / mu, 0) + np.minimum(G + lam / mu, 0)
        Z = D - L - S
        Y = Y + mu * Z
        mu = min(mu * rho, mu * 1e7)
        err = np.linalg.norm(Z, 'fro') / np.linalg.norm(D, 'fro')
        if err < tol_out or (tol_l1 > 0 and np.sum(np.abs(Z)) <= tol_l1):
            return L, S, it, True
    return L, S, max_iter, False


def compute_RPCA(image_array, max_error, max_iter=500):
    """compute_RPCA (computeRPCADecomposition.py:12-48), grayscale branch, with the engine this repo substitutes for the
    un-installable RobustPCA package (PARITY UNPINNED at that boundary): rank-1-capped inexact_alm_rpca per slice, stopped by
    sum|M - L - S| <= max_error."""
    L = np.zeros(image_array.shape)
    S = np.zeros(image_array.shape)
    its = []
    for i in range(image_array.shape[0]):
        L[i], S[i], it, conv = inexact_alm_rpca_capped(image_array[i], 1.0, 1, 1e-7, max_error, max_iter)
        its.append(it if conv else -it)
    return L, S, np.array(its)


def executeSaliencyRPCA(video_thw, downsample_ratio):
    """executeSaliencyRPCA (computeRPCADecomposition.py:52-95), grayscale."""
    xt = video_thw.transpose([2, 1, 0])
    yt = video_thw.transpose([1, 2, 0])
    if downsample_ratio != 1:
        xt = resize_with_cv2(xt, 1 / downsample_ratio)
        yt = resize_with_cv2(yt, 1 / downsample_ratio)
    xt = xt.astype(np.float64)
    yt = yt.astype(np.float64)
    xl, xs, _ = compute_RPCA(xt, xt.shape[1] * xt.shape[2] * 0.0001)
    yl, ys, _ = compute_RPCA(yt, yt.shape[1] * yt.shape[2] * 0.0001)
    return xl, xs, yl, ys
