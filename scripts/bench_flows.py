"""Extra bench.py workloads (BASELINE.json configs 2 and 5); each prints the same JSON line as the headline workload.

  highway_flow       config 2 on the committed half-resolution fixture of the reference's input/ frames 1-289
                     (tests/golden/highway_half_u8.npz): stage 1 flat LSD -> mask -> run_motion_saliency_check -> group-sparse
                     RPCA -> masks k = 2, 3 -> filter_sparse_map, scored with compute_score's F-measure
                     (/root/reference/precomputed_main.py:53-92).  The stage-2 saliency cube is the deterministic stand-in the
                     fixture was generated with (|x - temporal median|, tests/golden/make_golden.py): the reference's stage 2
                     needs the un-installable RobustPCA package.
  batch64_qvga_200   config 5: 64 independent 240x320x200 clips, one decomposition per clip, clips dealt round-robin to the
                     ranks (replicas only, no collective), several solver handles in flight per GPU.
"""
import json
import os
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _line(args, world, value, ms, workload_cfg, extra):
    out = {"metric": "frames/s decomposed", "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
           "dtype": "f32 (state) / exact int8-slice or fp64 Gram + f64 eigensolve", "config": workload_cfg}
    out.update(extra)
    return out


# ------------------------------------------------------------------------------------------------------------------
def highway_flow(args, rank, world, sampler_cls):
    import torch
    import background_subtraction_b200 as B
    from background_subtraction_b200 import _cabi as C
    if world > 1:
        raise SystemExit("highway_flow is a single-GPU workload (76 800 x 289 at full size: one decomposition)")
    fx = np.load(os.path.join(ROOT, "tests", "golden", "highway_half_u8.npz"))
    frames = fx["frames"]                                                   # [h, w, t] uint8
    h, w, t = frames.shape
    m = h * w
    x = np.asfortranarray(frames.astype(np.float64))
    B.normalizeImage(x)
    mean = float(np.mean(x))
    D = np.asfortranarray((x - mean).reshape((m, t), order='F'))
    med = np.median(x, axis=2, keepdims=True)
    sal = np.abs(x - med)
    sal /= sal.sum()                                                        # stand-in for the stage-2 saliency cube (see module doc)
    groups = B.get_proximal_flat_groups_nonoverlap((h, w), (3, 3))
    Dh = torch.from_numpy(np.ascontiguousarray(D.T, dtype=np.float32)).pin_memory().numpy()     # [t][m] float32, pinned
    sal_dev = torch.from_numpy(np.ascontiguousarray(sal, dtype=np.float32)).cuda()
    video_thw = torch.from_numpy(np.ascontiguousarray(frames.transpose(2, 0, 1), dtype=np.float32)).cuda()   # raw 0..255, [t, h, w]
    stream = torch.cuda.current_stream()
    stages = {}

    def timed(name, fn):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        r = fn()
        e1.record(stream)
        stages.setdefault(name, []).append((e0, e1))
        return r

    def flow(record):
        tm = timed if record else (lambda name, fn: fn())
        # stage 1: flat LSD + mask (lsd_improvement.py --alg_ver 0 with flat groups)
        cfg1 = B.make_config(m, t, C.PROX_FLAT_LINF, h, w)
        d1 = B.Decomposition(cfg1)
        d1.set_flat_groups(groups)

        def s1():
            d1.load(Dh)
            d1.run()
            return torch.from_numpy(d1.mask(2).T.copy()).cuda()            # [t][m] bool -> device
        mask1 = tm("stage1_lsd_flat", s1)
        mask1_hwt = mask1.view(t, w, h).permute(2, 1, 0)
        # stage 2: saliency RPCA of every X-T / Y-T slice + computeSCube (timed; the parity fields below use the stand-in cube the
        # golden labels were generated with, the cube of this stage is scored separately as F*_rpca_saliency)
        def s2():
            xl, xs, yl, ys = B.executeSaliencyRPCA(video_thw, 1)
            return B.computeSCube(xs, ys, return_device=True).permute(1, 2, 0).contiguous()      # [h, w, t] like precomputed_main.py:42
        cube2 = tm("stage2_saliency_rpca_scube", s2)
        # stage 3a: groups and lambdas from the mask and the saliency cube
        labels, ptr, lam = tm("motion_saliency", lambda: B.motion_saliency_blocks((h, w, t), mask1_hwt, sal_dev))
        # stage 3b: group-sparse RPCA, masks, size filter
        def s3():
            # = api.group_sparse_decomposition, fed from the pinned float32 copy of D
            dec = B.Decomposition(B.make_config(m, t, C.PROX_BLOCK_L2, h, w, delta=10, mu_scale=1.25, break_on_rank0=True,
                                                use_sv_prediction=True))
            dec.set_blocks(labels.cpu().numpy(), ptr, lam)
            dec.load(Dh)
            dec.run()
            st = dec.status()
            mk = [torch.from_numpy(dec.mask(k).T.copy()).cuda().view(t, w, h).permute(2, 1, 0) for k in (2, 3)]
            return dec, st, mk
        dec, st, mk = tm("group_sparse", s3)
        filt = tm("filter_sparse_map", lambda: [B.filter_sparse_map(v) for v in mk])
        torch.cuda.synchronize()
        return d1, labels, ptr, lam, dec, st, mk, filt, mask1_hwt, cube2

    for _ in range(max(args.warmup, 1)):
        flow(False)
    sampler = sampler_cls(0)
    sampler.start()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        d1, labels, ptr, lam, dec, st, mk, filt, mask1_hwt, cube2 = flow(True)
    e1.record(stream)
    torch.cuda.synchronize()
    clocks = sampler.stop()
    ms = e0.elapsed_time(e1) / args.steps
    # parity against what the reference's own functions produced for this fixture (tests/golden/make_golden.py)
    gold_labels, gold_ptr, gold_lam = fx["labels"], fx["lam_ptr"], fx["lam"]
    lab_h = labels.cpu().numpy()
    from oracle import score_oracle as SC                                  # checker only (F-measure of compute_score.py)
    gt = np.abs(x - med) > 0.12
    golden = {k: np.unpackbits(fx[key])[:m * t].reshape((m, t), order='F').reshape((h, w, t), order='F').astype(bool)
              for k, key in ((2, "gs_mask2"), (3, "gs_mask3"))}
    parity = {"blocks": int(ptr[-1]), "blocks_reference": int(gold_ptr[-1]),
              "labels_identical": bool(np.array_equal(lab_h, gold_labels) and np.array_equal(ptr, gold_ptr)),
              "lambda_max_rel_diff": float(np.max(np.abs(lam[:-1] - gold_lam) / gold_lam)) if len(lam) - 1 == len(gold_lam) else None,
              "group_sparse_iters": int(st.iter), "group_sparse_converged": bool(st.converged), "stage1_iters": int(d1.status().iter)}
    for i, k in enumerate((2, 3)):
        g = mk[i].cpu().numpy()
        parity["mask%d_agreement" % k] = float((g == golden[k]).mean())
        parity["F%d_gpu" % k] = SC.mean_fscore(g, gt)
        parity["F%d_reference" % k] = SC.mean_fscore(golden[k], gt)
        parity["F%d_filtered_gpu" % k] = SC.mean_fscore(filt[i].cpu().numpy(), gt)
    # the same flow once more (untimed) with the saliency cube of stage 2 instead of the stand-in
    try:
        lab2, ptr2, lam2 = B.motion_saliency_blocks((h, w, t), mask1_hwt, cube2)
        dec2 = B.Decomposition(B.make_config(m, t, C.PROX_BLOCK_L2, h, w, delta=10, mu_scale=1.25, break_on_rank0=True, use_sv_prediction=True))
        dec2.set_blocks(lab2.cpu().numpy(), ptr2, lam2)
        dec2.load(Dh)
        dec2.run()
        st2 = dec2.status()
        parity["rpca_saliency"] = {"blocks": int(ptr2[-1]), "iters": int(st2.iter), "converged": bool(st2.converged),
                                   "F2": SC.mean_fscore(dec2.mask(2).reshape((h, w, t), order='F'), gt),
                                   "F3": SC.mean_fscore(dec2.mask(3).reshape((h, w, t), order='F'), gt)}
    except Exception as ex:                                                 # e.g. no group survives the weight filter
        parity["rpca_saliency"] = {"error": str(ex)}
    stage_ms = {k: float(np.mean([a.elapsed_time(b) for a, b in v])) for k, v in stages.items()}
    cfg = {"workload": "highway_flow", "rows": h, "cols": w, "frames": t,
           "flow": "flat LSD -> mask -> saliency RPCA batch + computeSCube -> run_motion_saliency_check -> group-sparse RPCA -> masks k=2,3 -> filter_sparse_map",
           "fixture": "half-resolution input/ frames 1-289 (tests/golden/highway_half_u8.npz); stage-2 saliency = |x - median| stand-in",
           "timing": "CUDA events on the launching stream around the whole flow (host glue included); pinned float32 D in, masks out"}
    return _line(args, world, t / (ms * 1e-3), ms, cfg,
                 {"data": "reference fixture (input/ frames, half resolution)", "stages_ms": stage_ms, "parity_vs_reference": parity,
                  "clocks": clocks, "roofline": None, "cpu_baseline": None,
                  "e2e": {"value": t / (ms * 1e-3), "unit": "frames/s", "h2d_bytes_per_step": int(2 * Dh.nbytes),
                          "d2h_bytes_per_step": int(3 * m * t), "what": "the flow is the host-buffer API: value == e2e"},
                  "gpu_launches": None})


# ------------------------------------------------------------------------------------------------------------------
def batch_clips(args, rank, world, local_rank, sampler_cls, nclips=64, rows=240, cols=320, frames=200, seed0=100, nrect=3):
    import torch
    import torch.distributed as dist
    import background_subtraction_b200 as B
    from background_subtraction_b200 import _cabi as C, synth
    m = rows * cols
    mine = list(range(rank, nclips, world))
    in_flight = int(os.environ.get("BSUB_BATCH_IN_FLIGHT", "4"))
    groups = B.get_proximal_flat_groups_nonoverlap((rows, cols), (3, 3))
    clips = []
    for c in mine:                                                          # resident inputs: float32 [frames][m] on the device
        video, _ = synth.make_clip(rows, cols, frames, seed=seed0 + c, n_rect=nrect)
        clips.append(torch.from_numpy(synth.preprocess_u8(video)).cuda())
    # one solver handle per in-flight slot, reused for every clip it takes
    decs = []
    for _ in range(min(in_flight, len(mine))):
        d = B.Decomposition(B.make_config(m, frames, C.PROX_FLAT_LINF, rows, cols))
        d.set_flat_groups(groups)
        decs.append(d)
    results = [None] * len(mine)

    def run_all():
        nxt, lock, errs = [0], threading.Lock(), []

        def worker(slot):
            try:
                torch.cuda.set_device(local_rank)
                with torch.cuda.stream(torch.cuda.Stream()):
                    dec = decs[slot]
                    dec._stream = None
                    while True:
                        with lock:
                            i = nxt[0]
                            nxt[0] += 1
                        if i >= len(mine):
                            break
                        dec.load(clips[i])
                        dec.run()
                        st = dec.status()
                        mask = torch.empty((frames, m), dtype=torch.uint8, device="cuda")
                        C.check(dec.lib.bsub_finalize(dec.h, dec.stream()))
                        C.check(dec.lib.bsub_mask_stats_local(dec.h, 0, dec.stream()))      # max |S|, then count / sum / sum of squares
                        C.check(dec.lib.bsub_mask_stats_local(dec.h, 1, dec.stream()))
                        C.check(dec.lib.bsub_mask_dev(dec.h, 2.0, __import__("ctypes").c_void_p(mask.data_ptr()), dec.stream()))
                        torch.cuda.current_stream().synchronize()
                        results[i] = (int(st.iter), bool(st.converged), int(st.svp), float(mask.float().mean().item()))
            except Exception as ex:
                errs.append(ex)
        th = [threading.Thread(target=worker, args=(s,)) for s in range(len(decs))]
        for t_ in th:
            t_.start()
        for t_ in th:
            t_.join()
        if errs:
            raise errs[0]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 1)):
        run_all()
    barrier()
    sampler = sampler_cls(local_rank)
    if rank == 0:
        sampler.start()
    stream = torch.cuda.current_stream()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        run_all()
        torch.cuda.synchronize()                                            # the workers' streams are done; e1 then orders after them
    e1.record(stream)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    tmax = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms = float(tmax.item()) / args.steps
    if rank != 0:
        return None
    cfg = {"workload": "batch64_qvga_200", "clips": nclips, "rows": rows, "cols": cols, "frames": frames, "prox": "flat 3x3 l_inf (LSD)",
           "sharding": "whole clips dealt round-robin to %d GPU(s): replicas only, no collective" % world, "in_flight_per_gpu": len(decs),
           "timing": "CUDA events on rank 0's stream around all clips of a step (device-resident inputs; L, S and the mask materialised "
                     "on the device), max over ranks"}
    its = [r[0] for r in results]
    return _line(args, world, nclips * frames / (ms * 1e-3), ms, cfg,
                 {"data": "synthetic", "clips_this_rank": len(mine), "iters_min_max": [min(its), max(its)],
                  "all_converged": all(r[1] for r in results), "rank_L_min_max": [min(r[2] for r in results), max(r[2] for r in results)],
                  "mask_fraction_mean": float(np.mean([r[3] for r in results])), "clocks": clocks, "roofline": None, "cpu_baseline": None,
                  "e2e": {"value": None, "unit": "frames/s", "h2d_bytes_per_step": None, "d2h_bytes_per_step": None,
                          "what": "not measured for this workload"}, "gpu_launches": None})


# ------------------------------------------------------------------------------------------------------------------
def graph_lsd(args, rank, world, sampler_cls, rows, cols, frames, video, label):
    """The reference's DEFAULT LSD() mode (graphs=getGraphSPAMS_all_groups: overlapping 3x3 windows at every pixel,
    /root/reference/inexact_alm_lsd.py:203-235, :49-57): spill pass -> tile-local dual BCD (prox.cu) -> dual update, fp64 Gram.  One
    step = one whole decomposition (init + ALM iterations + L + mask), device-resident float32 input.  --gpus N > 1: pixel-column
    shards with the prox on whole frames after an all-to-all (this leg of the bench was written after the round's GPU budget was
    spent: the driver it calls is verified on 2 B200s by scripts/check_sharded_graph.py, the bench wrapper itself has not run)."""
    import ctypes
    import torch
    import background_subtraction_b200 as B
    from background_subtraction_b200 import _cabi as C, synth
    m = rows * cols
    stream = torch.cuda.current_stream()
    if world > 1:
        # pixel-column shards; the prox runs on whole frames after an all-to-all (dist.ShardedLSD._prox_on_frames), so the dominant
        # cost of this mode splits by frames.  Same choreography as scripts/check_sharded_graph.py (verified on 2 B200s).
        import torch.distributed as dist
        from background_subtraction_b200 import dist as bdist
        c0, c1 = bdist.shard_columns(cols, world, rank)
        cl = c1 - c0
        full = synth.preprocess_u8(video)
        D = torch.from_numpy(np.ascontiguousarray(full.reshape(frames, cols, rows)[:, c0:c1, :].reshape(frames, rows * cl))).cuda()
        solver = bdist.CudaStepSolver(rows, cl, frames, m, graph_cols=cols)
        dec = solver.dec
        driver = bdist.ShardedLSD(solver, bdist.TorchComm())

        def one():
            solver.load(D)
            driver.solve()
            return driver.finish(2.0, want_mask=True)
    else:
        D = torch.from_numpy(synth.preprocess_u8(video)).cuda()              # float32 [frames][m]
        dec = B.Decomposition(B.make_config(m, frames, C.PROX_GRAPH_LINF, rows, cols))
        dec.set_graph_windows(None)

        def one():
            dec.load(D)
            dec.run()
            C.check(dec.lib.bsub_finalize(dec.h, dec.stream()))
            mask = torch.empty((frames, m), dtype=torch.uint8, device="cuda")
            C.check(dec.lib.bsub_mask_stats_local(dec.h, 0, dec.stream()))
            C.check(dec.lib.bsub_mask_stats_local(dec.h, 1, dec.stream()))
            C.check(dec.lib.bsub_mask_dev(dec.h, 2.0, ctypes.c_void_p(mask.data_ptr()), dec.stream()))
            return mask

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 0)):          # --warmup 0 is allowed here: a 1080p x 300 graph solve takes minutes
        mask = one()
    barrier()
    sampler = sampler_cls(torch.cuda.current_device())
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        mask = one()
    e1.record(stream)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    tmax = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    frac = torch.tensor([float(mask.float().sum().item()), float(mask.numel())], dtype=torch.float64, device="cuda")
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(frac)
    ms = float(tmax.item()) / args.steps
    if rank != 0:
        return None
    st = dec.status()
    log = dec.log()
    cfg = {"workload": label, "rows": rows, "cols": cols, "frames": frames,
           "prox": "overlapping 3x3 windows at every pixel, l_inf (the reference's default LSD() graph mode)", "delta": 10,
           "sharding": "one GPU" if world == 1 else "pixel columns over %d GPUs, frames around the prox (all-to-all)" % world,
           "timing": "CUDA events on the launching stream around whole decompositions; device-resident float32 D"}
    return _line(args, world, frames / (ms * 1e-3), ms, cfg,
                 {"data": "reference fixture (WaterSurface)" if label.startswith("watersurface") else "synthetic",
                  "alm_iters": int(st.iter), "converged": bool(st.converged), "rank_L": int(st.svp), "err": float(st.err),
                  "rank_sequence": [int(l["svp"]) for l in log], "mask_fraction": float(frac[0].item() / frac[1].item()),
                  "ms_per_alm_iteration": ms / max(int(st.iter), 1), "clocks": clocks, "roofline": None, "cpu_baseline": None,
                  "e2e": {"value": None, "unit": "frames/s", "h2d_bytes_per_step": None, "d2h_bytes_per_step": None,
                          "what": "not measured for this workload"}, "gpu_launches": None})
