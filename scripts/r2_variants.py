"""Round-2 tuning probe: one resident 1080p x 300 clip, the solve timed under several environment switches of the shrink pass
(read at launch time by libbsub_b200.so): BSUB_FLAT_XPOSE, BSUB_FLAT_POLICY.  Prints ms per solve (best of 3) per variant."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import background_subtraction_b200 as B
from background_subtraction_b200 import _cabi as C, synth

rows, cols, n = 1080, 1920, 300
video, _ = synth.make_clip(rows, cols, n, seed=0, n_rect=6)
D = synth.preprocess_u8(video)
cfg = B.make_config(rows * cols, n, C.PROX_FLAT_LINF, rows, cols)
dec = B.Decomposition(cfg)
dec.set_flat_groups(B.get_proximal_flat_groups_nonoverlap((rows, cols), (3, 3)))
dec.load(D)
variants = [{"BSUB_FLAT_XPOSE": x, "BSUB_FLAT_POLICY": p} for x in ("0", "1", "2") for p in ("3", "0", "1", "2")]
ref_mask = None
for v in variants:
    os.environ.update(v)
    best = 1e9
    for rep in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        dec.run(); st = dec.status()
        torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
    mask = dec.mask(2)
    if ref_mask is None:
        ref_mask = mask
    print(v, "ms %.2f" % (best * 1e3), "iters", st.iter, "conv", st.converged, "mask agreement with first variant %.6f" % float((mask == ref_mask).mean()), flush=True)
