"""-m "not gpu": host-side logic of the package and the C-ABI library surface (no compute calls without a GPU)."""
import ctypes
import os
import re

import numpy as np
import pytest

import background_subtraction_b200 as B
from background_subtraction_b200 import _cabi as C
from background_subtraction_b200 import dist as bdist
from oracle import alm_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "bsub_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(bsub_[A-Za-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 40
    lib = C.load()                                     # builds with nvcc if the .so is missing; raises otherwise
    for name in declared:
        assert hasattr(lib, name), name
        assert name in C.SIGNATURES, "ctypes signature missing for " + name
    assert lib.bsub_version() >= 100
    cfg = C.Config()
    lib.bsub_default_config(ctypes.byref(cfg))
    assert (cfg.delta, cfg.mu_scale, cfg.rho, cfg.tol, cfg.max_iter, cfg.sv0) == (10.0, 12.5, 1.6, 1e-7, 500, 10)
    sizes = (ctypes.c_int32 * 3)()
    lib.bsub_abi_sizes(sizes)
    assert list(sizes) == [ctypes.sizeof(C.Config), ctypes.sizeof(C.Status), ctypes.sizeof(C.IterLog)]


def test_no_silent_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    D = np.zeros((36, 4), order='F')
    with pytest.raises(Exception, match="no CUDA device"):
        B.inexact_alm_lsd(D, groups=B.get_proximal_flat_groups_nonoverlap((6, 6), (3, 3)))
    with pytest.raises(Exception, match="no CUDA device"):
        B.foreground_mask(D, D, D)
    with pytest.raises(Exception, match="no CUDA device"):
        B.inexact_alm_lsd_batch([D, D], groups=B.get_proximal_flat_groups_nonoverlap((6, 6), (3, 3)))
    w = -np.ones((6, 6)); w[2, 3] = 1.0
    with pytest.raises(Exception, match="no CUDA device"):
        B.inexact_alm_lsd_with_background(D, [B.get_proximal_graph_group_centers((6, 6), 1, w)] * 4,
                                          [(w < 0).flatten(order='F')] * 4)
    with pytest.raises(Exception, match="graphs must be list/array"):
        B.inexact_alm_lsd_with_background(D, B.get_proximal_graph_group_centers((6, 6), 1, w), [(w < 0).flatten(order='F')] * 4)
    src = "".join(open(os.path.join(ROOT, "background-subtraction_b200", f)).read()
                  for f in ("api.py", "dist.py", "_cabi.py", "synth.py", "build.py", "__init__.py"))
    assert "oracle" not in src.replace("oracle's", "")          # the product never imports the checker


def test_reference_error_convention():
    D = np.zeros((36, 4), order='F')
    with pytest.raises(Exception, match="one of graphs or groups must not be None"):
        B.inexact_alm_lsd(D)
    with pytest.raises(Exception, match="only one of graphs or groups must not be None"):
        B.inexact_alm_lsd(D, graphs=B.getGraphSPAMS_all_groups((6, 6), (3, 3)), groups=np.ones(36, dtype=np.int32))
    with pytest.raises(Exception, match="Input lengths are incorrect"):
        B.get_proximal_flat_groups_nonoverlap((6, 6, 1), (3, 3))
    with pytest.raises(Exception, match="Input lengths are incorrect"):
        B.getGraphSPAMS_all_groups((6,), (3, 3))


@pytest.mark.parametrize("shape", [(128, 160), (31, 41), (5, 6), (3, 3), (4, 7), (240, 320), (2, 9)])
def test_group_and_graph_builders(shape):
    g = B.get_proximal_flat_groups_nonoverlap(shape, (3, 3))
    assert g.dtype == np.int32 and np.array_equal(g, O.flat_groups_nonoverlap(shape, (3, 3)))
    assert B.detect_flat_tiling(g) == shape or shape[1] <= 3
    assert B.detect_flat_tiling(g, shape) == shape
    ip, ix = B.window_csc(shape)
    ip2, ix2, eta = O.graph_all_groups(shape, (3, 3))
    assert np.array_equal(ip, ip2) and np.array_equal(ix, ix2)
    graph = B.getGraphSPAMS_all_groups(shape, (3, 3))
    assert set(graph) == {"eta_g", "groups", "groups_var"} and graph["groups"].nnz == 0
    det = B.detect_window_graph(graph, shape[0] * shape[1], shape)
    assert det is not None and det[:2] == shape and np.all(det[2] == 1.0)


def test_detect_rejects_other_inputs():
    g = O.flat_groups_nonoverlap((12, 12), (4, 2))
    assert B.detect_flat_tiling(g) is None
    rng = np.random.default_rng(0)
    assert B.detect_flat_tiling(rng.integers(1, 5, size=144).astype(np.int32)) is None
    graph = B.getGraphSPAMS_all_groups((8, 9), (3, 3))
    graph["groups_var"] = graph["groups_var"][:, :-1]
    assert B.detect_window_graph(graph, 72) is None


def test_center_window_graphs():
    """get_proximal_graph_group_centers mirror: same CSC arrays as the oracle restatement (utils.py:234-246,
    lsd_improvement.py:74-120), recoverable from a bare SPAMS dict, not confused with the all-windows graph."""
    import background_subtraction_b200 as B
    from oracle import alm_oracle as O
    rng = np.random.default_rng(0)
    for shape in ((7, 9), (12, 5), (3, 3)):
        w = np.where(rng.random(shape) < 0.35, rng.choice([1.0, 1.5], shape), -1.0)
        w[0, 0], w[-1, -1] = 1.5, 1.0
        p, i, e = B.center_window_csc(shape, w)
        p2, i2, e2 = O.graph_group_centers(shape, 1, w)
        assert np.array_equal(p, p2) and np.array_equal(i, i2) and np.array_equal(e, e2)
        g = B.get_proximal_graph_group_centers(shape, 1, w)
        bare = {k: v for k, v in g.items() if not k.startswith('_')}
        emap = np.where(w > 0, w, 0).astype(np.float32).flatten(order='F')
        for cand in (g, bare):
            det = B.detect_center_windows(cand, w.size)
            assert det is not None and det[0] == shape and np.array_equal(det[1], emap)
    assert B.detect_center_windows(B.getGraphSPAMS_all_groups((7, 9), (3, 3)), 63) is None


def test_labels_from_blocks():
    m, n = 10, 3
    b0 = np.zeros(m, bool); b0[:4] = True
    b1 = np.zeros(m, bool); b1[3:6] = True           # overlaps b0 at pixel 3 -> later block wins (sequential assignment)
    labels, ptr, lam = B.labels_from_blocks([[b0, b1], [], [b1]], [[0.1, 0.2], [], [0.3]], m)
    assert labels.tolist()[0] == [1, 1, 1, 2, 2, 2, 0, 0, 0, 0] and labels[1].sum() == 0
    assert ptr.tolist() == [0, 2, 2, 3] and lam[:3].tolist() == [0.1, 0.2, 0.3]
    lo, po = O.blocks_to_labels([[b0, b1], [], [b1]], m)
    assert np.array_equal(lo, labels) and np.array_equal(po, ptr)
    with pytest.raises(Exception):
        B.labels_from_blocks([[b0]], [[0.1, 0.2]], m)


@pytest.mark.parametrize("cols,world", [(1920, 8), (1920, 3), (160, 2), (7, 4), (3840, 8), (10, 1)])
def test_shard_columns(cols, world):
    spans = [bdist.shard_columns(cols, world, r) for r in range(world)]
    assert spans[0][0] == 0 and spans[-1][1] == cols
    for (a0, a1), (b0, b1) in zip(spans[:-1], spans[1:]):
        assert a1 == b0 and a0 % 3 == 0 and a1 % 3 == 0 or a1 == cols      # boundaries on whole column triples
    widths = [b - a for a, b in spans]
    assert max(widths) - min(widths) <= 3 or cols < 3 * world


def test_frames_major_is_zero_copy_for_fortran_input():
    from background_subtraction_b200 import api
    D = np.asfortranarray(np.arange(12, dtype=np.float64).reshape(4, 3))
    A = api._frames_major(D)
    assert A.shape == (3, 4) and A.flags.c_contiguous and np.shares_memory(A, D)
    A2 = api._frames_major(np.ascontiguousarray(D))
    assert np.array_equal(A2, A)


def test_synthetic_clip_is_deterministic():
    from background_subtraction_b200 import synth
    v1, g1 = synth.make_clip(24, 30, 6, seed=3, n_rect=2, return_gt=True)
    v2, _ = synth.make_clip(24, 30, 6, seed=3, n_rect=2)
    assert v1.dtype == np.uint8 and v1.shape == (6, 720) and np.array_equal(v1, v2) and 0 < g1.mean() < 0.5
    D = synth.preprocess_u8(v1)
    cube = np.asfortranarray(v1.reshape(6, 30, 24).transpose(2, 1, 0))
    Dref, _x, _m = O.normalize_and_center(cube)
    assert np.abs(D.T.astype(np.float64) - Dref).max() <= 1e-7
