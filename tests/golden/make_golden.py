"""tests/golden/make_golden.py -- generates the committed golden fixtures.

Run once in the build container (needs /root/reference; never runs on the GPU box):
    python tests/golden/make_golden.py

Everything written here is produced by the REFERENCE'S OWN CODE, imported in place and unmodified by
oracle/ref_harness.py (the only substitution is the `spams` stand-in documented there):

  watersurface_u8.npz   input fixture: ImData of the reference's data/WaterSurface.mat (uint8 128x160x48)
  highway_half_u8.npz   input fixture: frames 1-289 of the reference's input/ sequence, grayscale
                        (reference loader utils.py:68-86) resized x0.5 with the reference's own
                        resize_with_cv2 (utils.py:129-136, INTER_AREA) and rounded to uint8;
                        plus the per-frame block label maps / lambdas produced by the reference's
                        run_motion_saliency_check (motion_saliency_check.py:66-120)
  golden_cases.npz      full L, S, masks, per-iteration (svp, err) of small crops through
                        inexact_alm_lsd (flat and graph), inexact_alm_group_sparse_RPCA,
                        block_shrinkage_operator, foreground_mask, prox stand-ins
  golden_summary.json   iteration counts / rank sequences / err tails / norms / mask fractions of the
                        full-size reference runs (WaterSurface flat d=10, d=1; highway-half group-sparse)
"""
import json
import os
import re
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import alm_oracle as O  # noqa: E402
from oracle import ref_harness as R  # noqa: E402

ITER_RE = re.compile(r"Iteration:\s+(\d+) rank\(L\):\s+(\d+) \|\|S\|\|_0: (\S+) err: (\S+)")


def parse_log(text):
    out = [(int(a), int(b), float(d)) for a, b, _c, d in ITER_RE.findall(text)]
    return np.array([x[1] for x in out], dtype=np.int32), np.array([x[2] for x in out], dtype=np.float64)


def crop_D(cube_u8, r0, r1, c0, c1, t0, t1):
    """Reference preprocessing on a crop: normalizeImage + mean-subtract + F-reshape
    (inexact_alm_lsd.py:211-225)."""
    u = R.load()["utils"]
    x = np.asfortranarray(cube_u8[r0:r1, c0:c1, t0:t1].astype(np.float64))
    u.normalizeImage(x)
    x = x - np.mean(x)
    h, w, t = x.shape
    return x.reshape((h * w, t), order='F'), (h, w, t)


def synthetic_blocks(shape, rng):
    """Deterministic disjoint rectangular blocks per frame (0-3 per frame), as bool[m] F-order masks."""
    h, w, t = shape
    blocks, lambdas = [], []
    lam = 1.0 / (10 * np.sqrt(max(h * w, t)))
    for f in range(t):
        nb = int(rng.integers(0, 4))
        bl, ll = [], []
        used = np.zeros((h, w), dtype=bool)
        for _ in range(nb):
            bh, bw = int(rng.integers(3, h // 3)), int(rng.integers(3, w // 3))
            i0, j0 = int(rng.integers(0, h - bh)), int(rng.integers(0, w - bw))
            mk = np.zeros((h, w), dtype=bool)
            mk[i0:i0 + bh, j0:j0 + bw] = True
            mk &= ~used
            if mk.sum() == 0:
                continue
            used |= mk
            bl.append(mk.flatten(order='F'))
            ll.append(float(lam * rng.uniform(0.3, 1.0)))
        blocks.append(bl)
        lambdas.append(ll)
    return blocks, lambdas


def main():
    mods = R.load()
    U, LSDm, GS, LI, MS = (mods["utils"], mods["inexact_alm_lsd"], mods["group_sparse_RPCA"],
                           mods["lsd_improvement"], mods["motion_saliency_check"])
    cases = {}
    summary = {"generator": "tests/golden/make_golden.py", "numpy": np.__version__,
               "spams": "stand-in (oracle/prox_oracle.c)" if not R.real_spams() else "real"}

    # ---------------- input fixtures ----------------
    ws = R.load_watersurface()
    np.savez_compressed(os.path.join(HERE, "watersurface_u8.npz"), ImData=ws)

    hw_full = R.load_input_frames(0, 288)                      # float64 [240,320,289]
    hw_half = U.resize_with_cv2(hw_full, 0.5)                  # reference's own resize
    hw_half_u8 = np.asfortranarray(np.clip(np.rint(hw_half), 0, 255).astype(np.uint8))
    print("highway half", hw_half_u8.shape)

    # ---------------- small LSD cases (flat) ----------------
    for name, crop in {"flat_a": (40, 72, 60, 100, 0, 16),      # 32x40x16, cols%3=1
                       "flat_b": (10, 41, 20, 61, 8, 20)}.items():   # 31x41x12, rows%3=1 cols%3=2
        D, shp = crop_D(ws, *crop)
        groups = LI.get_proximal_flat_groups_nonoverlap(shp[:2], (3, 3))
        with R.quiet() as buf:
            L, S, it, conv = LSDm.inexact_alm_lsd(D, groups=groups)
        svp, err = parse_log(buf.getvalue())
        mask = U.foreground_mask(D, L, S)
        cases.update({f"{name}_crop": np.array(crop), f"{name}_L": L, f"{name}_S": S, f"{name}_iter": it,
                      f"{name}_conv": conv, f"{name}_svp": svp, f"{name}_err": err,
                      f"{name}_mask": np.packbits(mask.ravel(order='F')), f"{name}_groups": groups})
        print(name, shp, it, conv, svp.tolist(), err[-1])

    # ---------------- small LSD case (overlapping graph) ----------------
    crop = (50, 74, 70, 100, 0, 10)                              # 24x30x10
    D, shp = crop_D(ws, *crop)
    with R.quiet() as buf:
        graph = LSDm.getGraphSPAMS_all_groups(shp[:2], (3, 3))
        t0 = time.time()
        L, S, it, conv = LSDm.inexact_alm_lsd(D, graphs=graph)
    svp, err = parse_log(buf.getvalue())
    mask = U.foreground_mask(D, L, S)
    cases.update({"graph_a_crop": np.array(crop), "graph_a_L": L, "graph_a_S": S, "graph_a_iter": it,
                  "graph_a_conv": conv, "graph_a_svp": svp, "graph_a_err": err,
                  "graph_a_mask": np.packbits(mask.ravel(order='F'))})
    print("graph_a", shp, it, conv, svp.tolist(), err[-1], "%.1fs" % (time.time() - t0))

    # ---------------- small group-sparse case ----------------
    crop = (40, 72, 60, 100, 0, 16)
    D, shp = crop_D(ws, *crop)
    rng = np.random.default_rng(7)
    blocks, lambdas = synthetic_blocks(shp, rng)
    with R.quiet() as buf:
        L, S, it, conv = GS.inexact_alm_group_sparse_RPCA(D, blocks, lambdas, delta=10)
    svp, err = parse_log(buf.getvalue())
    labels, ptr = O.blocks_to_labels(blocks, D.shape[0])
    lam_flat = np.array([x for l in lambdas for x in l], dtype=np.float64)
    cases.update({"gs_a_crop": np.array(crop), "gs_a_L": L, "gs_a_S": S, "gs_a_iter": it, "gs_a_conv": conv,
                  "gs_a_svp": svp, "gs_a_err": err, "gs_a_labels": labels.astype(np.uint8), "gs_a_lam_ptr": ptr,
                  "gs_a_lam": lam_flat,
                  "gs_a_mask2": np.packbits(U.foreground_mask(D, L, S, 2).ravel(order='F')),
                  "gs_a_mask3": np.packbits(U.foreground_mask(D, L, S, 3).ravel(order='F'))})
    print("gs_a", shp, it, conv, svp.tolist(), err[-1] if len(err) else None)

    # operator-level goldens: block shrinkage + the prox stand-ins on a fixed random matrix
    rng = np.random.default_rng(11)
    G = np.asfortranarray(rng.standard_normal(D.shape) * 0.05)
    with np.errstate(divide='ignore', invalid='ignore'):
        cases["bs_G"] = G
        cases["bs_out"] = GS.block_shrinkage_operator(G, blocks, lambdas, 3.0, 0.02)
    groups = LI.get_proximal_flat_groups_nonoverlap(shp[:2], (3, 3))
    cases["pf_out"] = LSDm.prox_flat(G, 0.04, groups)
    with R.quiet():
        graph = LSDm.getGraphSPAMS_all_groups(shp[:2], (3, 3))
    cases["pg_out"] = LSDm.prox(G[:, :4], 0.04, graph)

    np.savez_compressed(os.path.join(HERE, "golden_cases.npz"), **cases)

    # ---------------- full WaterSurface, flat, delta = 10 and 1 ----------------
    D, shp = crop_D(ws, 0, 128, 0, 160, 0, 48)
    groups = LI.get_proximal_flat_groups_nonoverlap(shp[:2], (3, 3))
    for delta in (10, 1):
        with R.quiet() as buf:
            t0 = time.time()
            L, S, it, conv = LSDm.inexact_alm_lsd(D, groups=groups, delta=delta)
            dt = time.time() - t0
        svp, err = parse_log(buf.getvalue())
        mask = U.foreground_mask(D, L, S)
        summary[f"watersurface_flat_delta{delta}"] = dict(
            iters=it, converged=bool(conv), svp=svp.tolist(), err=err.tolist(), normL=float(np.linalg.norm(L)),
            normS=float(np.linalg.norm(S)), mask_fraction=float(mask.mean()), mask_count=int(mask.sum()),
            cpu_seconds=dt, norm_two=float(np.linalg.norm(D, 2)), norm_fro=float(np.linalg.norm(D)),
            norm_inf_rowsum=float(np.linalg.norm(D, np.inf)))
        print("ws flat delta", delta, it, conv, svp.tolist(), err[-1], "%.1fs" % dt)
        if delta == 10:
            np.savez_compressed(os.path.join(HERE, "watersurface_flat_mask.npz"),
                                mask=np.packbits(mask.ravel(order='F')))

    # ---------------- highway (half res) group-sparse flow ----------------
    x = np.asfortranarray(hw_half_u8.astype(np.float64))
    U.normalizeImage(x)
    mean = float(np.mean(x))
    xc = x - mean
    h, w, t = xc.shape
    D = np.asfortranarray(xc.reshape((h * w, t), order='F'))
    groups = LI.get_proximal_flat_groups_nonoverlap((h, w), (3, 3))
    with R.quiet() as buf:
        t0 = time.time()
        L1, S1, it1, conv1 = LSDm.inexact_alm_lsd(D, groups=groups)       # stage 1 (flat LSD)
        dt1 = time.time() - t0
    svp1, err1 = parse_log(buf.getvalue())
    mask1 = U.foreground_mask(D, L1, S1).reshape((h, w, t), order='F')
    # deterministic stand-in saliency cube (the reference's RobustPCA-based stage 2 is not installable):
    med = np.median(x, axis=2, keepdims=True)
    sal = np.abs(x - med)
    sal /= sal.sum()
    with R.quiet():
        blocks, lambdas = MS.run_motion_saliency_check(xc, mask1, sal)
    with R.quiet() as buf:
        t0 = time.time()
        L2, S2, it2, conv2 = GS.inexact_alm_group_sparse_RPCA(D, blocks, lambdas, delta=10)
        dt2 = time.time() - t0
    svp2, err2 = parse_log(buf.getvalue())
    labels, ptr = O.blocks_to_labels(blocks, D.shape[0])
    lam_flat = np.array([v for l in lambdas for v in l], dtype=np.float64)
    m2 = U.foreground_mask(D, L2, S2, 2)
    m3 = U.foreground_mask(D, L2, S2, 3)
    np.savez_compressed(os.path.join(HERE, "highway_half_u8.npz"), frames=hw_half_u8, labels=labels.astype(np.uint8),
                        lam_ptr=ptr, lam=lam_flat, gs_mask2=np.packbits(m2.ravel(order='F')),
                        gs_mask3=np.packbits(m3.ravel(order='F')),
                        lsd_mask=np.packbits(mask1.ravel(order='F')))
    summary["highway_half_lsd_flat"] = dict(iters=it1, converged=bool(conv1), svp=svp1.tolist(), err=err1.tolist(),
                                            normL=float(np.linalg.norm(L1)), normS=float(np.linalg.norm(S1)),
                                            mask_fraction=float(mask1.mean()), cpu_seconds=dt1, mean=mean)
    summary["highway_half_group_sparse"] = dict(iters=it2, converged=bool(conv2), svp=svp2.tolist(), err=err2.tolist(),
                                                normL=float(np.linalg.norm(L2)), normS=float(np.linalg.norm(S2)),
                                                mask2_fraction=float(m2.mean()), mask3_fraction=float(m3.mean()),
                                                blocks_total=int(len(lam_flat)), cpu_seconds=dt2)
    print("highway lsd", it1, conv1, svp1.tolist(), "%.1fs" % dt1)
    print("highway gs", it2, conv2, svp2.tolist(), "%.1fs" % dt2, "blocks", len(lam_flat))

    with open(os.path.join(HERE, "golden_summary.json"), "w") as f:
        json.dump(summary, f, indent=1)


if __name__ == "__main__":
    main()
