"""-m "not gpu": the pixel-column sharded driver (background-subtraction_b200/dist.py) on 2 CPU ranks over gloo.
Each rank runs the real ShardedLSD choreography (what is all-reduced, when, and the stop logic) around a NumPy step
solver; the stitched result must equal the single-process oracle on the whole matrix."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, rows, cols, n, out_dir):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from background_subtraction_b200 import dist as bdist, synth
    from np_step_solver import NumpyStepSolver
    from oracle import alm_oracle as O
    video, _ = synth.make_clip(rows, cols, n, seed=11, n_rect=2)
    cube = np.asfortranarray(video.reshape(n, cols, rows).transpose(2, 1, 0))
    D, _x, _mean = O.normalize_and_center(cube)                      # every rank derives the same global statistics
    c0, c1 = bdist.shard_columns(cols, world, rank)
    D_local = D[c0 * rows:c1 * rows, :]
    solver = NumpyStepSolver(D_local, rows, c1 - c0, n, rows * cols)
    driver = bdist.ShardedLSD(solver, bdist.TorchComm(), run_ahead=1)
    driver.solve()
    mask = driver.finish(2.0)
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), L=solver.L, S=solver.S, mask=mask, c0=c0, c1=c1, iters=solver.iter,
             conv=solver.converged, svp=[l[1] for l in solver.log], err=[l[2] for l in solver.log])
    dist.destroy_process_group()


@pytest.mark.parametrize("cols", [20, 15])
def test_sharded_driver_matches_oracle(tmp_path, cols):
    rows, n, world = 18, 12, 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, rows, cols, n, str(tmp_path)), nprocs=world, join=True)
    sys.path.insert(0, ROOT)
    from background_subtraction_b200 import synth
    from oracle import alm_oracle as O
    video, _ = synth.make_clip(rows, cols, n, seed=11, n_rect=2)
    cube = np.asfortranarray(video.reshape(n, cols, rows).transpose(2, 1, 0))
    D, _x, _mean = O.normalize_and_center(cube)
    log = []
    L, S, it, conv = O.inexact_alm_lsd(D, groups=O.flat_groups_nonoverlap((rows, cols), (3, 3)), log=log)
    mask = O.foreground_mask(D, L, S)
    parts = [np.load(os.path.join(str(tmp_path), "rank%d.npz" % r)) for r in range(world)]
    assert parts[0]["c1"] == parts[1]["c0"] and parts[0]["c1"] % 3 == 0
    Ls = np.vstack([p["L"] for p in parts]); Ss = np.vstack([p["S"] for p in parts]); Ms = np.vstack([p["mask"] for p in parts])
    for p in parts:
        assert int(p["iters"]) == it and bool(p["conv"]) == conv and p["svp"].tolist() == [l["svp"] for l in log]
        assert np.allclose(p["err"], [l["err"] for l in log], rtol=1e-6)
    assert np.abs(Ls - L).max() <= 1e-9 and np.abs(Ss - S).max() <= 1e-9
    assert np.array_equal(Ms, mask)


def _blocks_for(rows, cols, n):
    """two rectangular blocks per frame (one of them crossing the shard border at cols / 2), a frame without blocks, lambdas"""
    rng = np.random.default_rng(5)
    labels = np.zeros((n, cols, rows), dtype=np.int64)
    lams = []
    for f in range(n):
        if f == 3:
            lams.append([])
            continue
        c0 = cols // 2 - 2 + (f % 3)
        labels[f, c0:c0 + 4, 2:8] = 1
        labels[f, 1:3, 10:15] = 2
        lams.append([float(0.02 + 0.01 * rng.random()), float(0.03 + 0.01 * rng.random())])
    return labels, lams


def _worker_l2(rank, world, port, rows, cols, n, out_dir):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from background_subtraction_b200 import dist as bdist, synth
    from np_step_solver import NumpyStepSolver
    from oracle import alm_oracle as O
    video, _ = synth.make_clip(rows, cols, n, seed=11, n_rect=2)
    cube = np.asfortranarray(video.reshape(n, cols, rows).transpose(2, 1, 0))
    D, _x, _mean = O.normalize_and_center(cube)
    labels, lams = _blocks_for(rows, cols, n)
    c0, c1 = bdist.shard_columns(cols, world, rank)
    solver = NumpyStepSolver(D[c0 * rows:c1 * rows, :], rows, c1 - c0, n, rows * cols,
                             labels=labels[:, c0:c1, :].reshape(n, -1), lambdas_by_frame=lams)
    driver = bdist.ShardedLSD(solver, bdist.TorchComm(), run_ahead=1)
    driver.solve()
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), L=solver.L, S=solver.S, c0=c0, c1=c1, iters=solver.iter,
             conv=solver.converged, svp=[l[1] for l in solver.log])
    dist.destroy_process_group()


def test_sharded_group_sparse_matches_oracle(tmp_path):
    """l2-block mode on 2 ranks: the per-(frame, group) sums of squares are all-reduced between the two halves of the shrink pass
    (blocks and the frame-wide complement group span the shards, /root/reference/group_sparse_RPCA.py:29-40)."""
    rows, cols, n, world = 18, 21, 12, 2
    port = _free_port()
    mp.spawn(_worker_l2, args=(world, port, rows, cols, n, str(tmp_path)), nprocs=world, join=True)
    sys.path.insert(0, ROOT)
    from background_subtraction_b200 import synth
    from oracle import alm_oracle as O
    video, _ = synth.make_clip(rows, cols, n, seed=11, n_rect=2)
    cube = np.asfortranarray(video.reshape(n, cols, rows).transpose(2, 1, 0))
    D, _x, _mean = O.normalize_and_center(cube)
    labels, lams = _blocks_for(rows, cols, n)
    flat = labels.reshape(n, -1)
    blocks = [[flat[f] == b + 1 for b in range(len(lams[f]))] for f in range(n)]
    log = []
    L, S, it, conv = O.inexact_alm_group_sparse_RPCA(D, blocks, lams, log=log)
    parts = [np.load(os.path.join(str(tmp_path), "rank%d.npz" % r)) for r in range(world)]
    Ls = np.vstack([p["L"] for p in parts]); Ss = np.vstack([p["S"] for p in parts])
    for p in parts:
        assert int(p["iters"]) == it and bool(p["conv"]) == conv
        assert p["svp"].tolist() == [l["svp"] for l in log if l["err"] is not None]
    assert np.abs(Ls - L).max() <= 1e-9 and np.abs(Ss - S).max() <= 1e-9


def _worker_graph(rank, world, port, rows, cols, n, out_dir):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from background_subtraction_b200 import dist as bdist, synth
    from np_step_solver import NumpyStepSolver
    from oracle import alm_oracle as O
    video, _ = synth.make_clip(rows, cols, n, seed=11, n_rect=2)
    cube = np.asfortranarray(video.reshape(n, cols, rows).transpose(2, 1, 0))
    D, _x, _mean = O.normalize_and_center(cube)
    c0, c1 = bdist.shard_columns(cols, world, rank)
    solver = NumpyStepSolver(D[c0 * rows:c1 * rows, :], rows, c1 - c0, n, rows * cols, graph_cols=cols)
    driver = bdist.ShardedLSD(solver, bdist.TorchComm(), run_ahead=2)
    driver.solve()
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), L=solver.L, S=solver.S, iters=solver.iter, conv=solver.converged,
             svp=[l[1] for l in solver.log])
    dist.destroy_process_group()


def test_sharded_graph_mode_matches_oracle(tmp_path):
    """Overlapping-window LSD on 2 ranks: G_S is re-sharded by frames (all-to-all; 7 frames -> 4 + 3), the prox runs on whole
    frames, S comes back by columns (SURVEY 8e "Exceptions", solved exactly instead of with a halo)."""
    rows, cols, n, world = 12, 15, 7, 2
    port = _free_port()
    mp.spawn(_worker_graph, args=(world, port, rows, cols, n, str(tmp_path)), nprocs=world, join=True)
    sys.path.insert(0, ROOT)
    from background_subtraction_b200 import synth
    from oracle import alm_oracle as O
    video, _ = synth.make_clip(rows, cols, n, seed=11, n_rect=2)
    cube = np.asfortranarray(video.reshape(n, cols, rows).transpose(2, 1, 0))
    D, _x, _mean = O.normalize_and_center(cube)
    log = []
    L, S, it, conv = O.inexact_alm_lsd(D, graphs=O.graph_all_groups((rows, cols), (3, 3)), log=log)
    parts = [np.load(os.path.join(str(tmp_path), "rank%d.npz" % r)) for r in range(world)]
    Ls = np.vstack([p["L"] for p in parts]); Ss = np.vstack([p["S"] for p in parts])
    for p in parts:
        assert int(p["iters"]) == it and bool(p["conv"]) == conv and p["svp"].tolist() == [l["svp"] for l in log]
    assert np.abs(Ls - L).max() <= 1e-9 and np.abs(Ss - S).max() <= 1e-9
