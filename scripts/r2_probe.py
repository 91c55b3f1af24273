"""Round-2 probe: one 1080p x 300 solve with per-iteration counters (eig fast path, steps), used under ncu for launch lists."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import background_subtraction_b200 as B
from background_subtraction_b200 import _cabi as C, synth

rows, cols, n = (int(v) for v in (sys.argv[1:4] if len(sys.argv) > 3 else (1080, 1920, 300)))
video, _ = synth.make_clip(rows, cols, n, seed=0, n_rect=6)
D = synth.preprocess_u8(video)
max_iter = int(sys.argv[4]) if len(sys.argv) > 4 else 500
cfg = B.make_config(rows * cols, n, C.PROX_FLAT_LINF, rows, cols, max_iter=max_iter)
dec = B.Decomposition(cfg)
dec.set_flat_groups(B.get_proximal_flat_groups_nonoverlap((rows, cols), (3, 3)))
dec.load(D)
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    dec.run(); st = dec.status()
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print("rep", rep, "ms", round(dt * 1e3, 2), "iters", st.iter, "conv", st.converged, "counters", dec.counters(), flush=True)
print("svp", [l['svp'] for l in dec.log()], "sv", [l['sv'] for l in dec.log()])
import ctypes
out = (ctypes.c_int64 * 16)()
C.check(dec.lib.bsub_debug_eig_cycles(dec.h, out))
print("eig fast-path cycles of the last call [load, matvec, grams, jacobi, rotate, resid, chol+solve, certificate]:", [int(v) for v in list(out)[8:16]])
print(dec.debug_info())
