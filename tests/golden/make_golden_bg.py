"""tests/golden/make_golden_bg.py -- golden fixture for inexact_alm_lsd_with_background.

Run once in the build container (needs /root/reference):  python tests/golden/make_golden_bg.py

golden_bg.npz is produced by the REFERENCE'S OWN CODE imported in place by oracle/ref_harness.py
(lsd_improvement.inexact_alm_lsd_with_background, get_proximal_graph_group_centers, merge_masks; `spams` replaced by
the documented stand-in).  Inputs: the flat_a crop of WaterSurface (32x40x16) and a weight cube built the way
build_improved_LSD_graphs does (lsd_improvement.py:368-434): stage-1 mask = the reference mask of the flat_a golden,
"morph" mask = its 3x3 binary dilation (the reference uses skimage disk dilation + closing, which is not installed and
is outside the hot path), weights (1, 1.5), background marker -1.
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import ref_harness as R  # noqa: E402
from make_golden import crop_D, parse_log  # noqa: E402


def dilate3(mask):
    h, w, t = mask.shape
    p = np.zeros((h + 2, w + 2, t), dtype=bool)
    p[1:-1, 1:-1] = mask
    out = np.zeros_like(mask)
    for di in range(3):
        for dj in range(3):
            out |= p[di:di + h, dj:dj + w]
    return out


def main():
    mods = R.load()
    LI = mods["lsd_improvement"]
    ws = np.load(os.path.join(HERE, "watersurface_u8.npz"))["ImData"]
    g = np.load(os.path.join(HERE, "golden_cases.npz"))
    crop = tuple(int(x) for x in g["flat_a_crop"])
    D, shp = crop_D(ws, *crop)
    m = shp[0] * shp[1]
    S_mask = np.unpackbits(g["flat_a_mask"])[:m * shp[2]].astype(bool).reshape((m, shp[2]), order='F').reshape(shp, order='F')
    weight_mask = LI.merge_masks((S_mask, dilate3(S_mask)), (1, 1.5))
    graphs = [LI.get_proximal_graph_group_centers(weight_mask[:, :, i].shape, 1, group_centers=weight_mask[:, :, i])
              for i in range(shp[2])]
    background_masks = [(weight_mask[:, :, i] < 0).flatten(order='F') for i in range(shp[2])]
    t0 = time.time()
    with R.quiet() as buf:
        L, S, it, conv = LI.inexact_alm_lsd_with_background(D, graphs, background_masks)
    svp, err = parse_log(buf.getvalue())
    print("bg_a", shp, it, conv, svp.tolist(), err[-1], "groups/frame", [len(x['eta_g']) for x in graphs][:6],
          "%.1fs" % (time.time() - t0))
    np.savez_compressed(os.path.join(HERE, "golden_bg.npz"), bg_a_crop=np.array(crop), bg_a_weights=weight_mask.astype(np.float32),
                        bg_a_L=L, bg_a_S=S, bg_a_iter=it, bg_a_conv=conv, bg_a_svp=svp, bg_a_err=err)


if __name__ == "__main__":
    main()
