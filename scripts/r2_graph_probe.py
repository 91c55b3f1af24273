"""Round-2 probe: graph-mode (overlapping windows) LSD on the WaterSurface clip, max_iter ALM iterations; used under ncu to capture
prox_graph3_tile_kernel (the first launch is ALM iteration 1, the expensive one)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import background_subtraction_b200 as B
from background_subtraction_b200 import _cabi as C, synth

max_iter = int(sys.argv[1]) if len(sys.argv) > 1 else 500
cube = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "watersurface_u8.npz"))["ImData"]
rows, cols, n = cube.shape
video = np.ascontiguousarray(cube.transpose(2, 1, 0)).reshape(n, rows * cols)
D = torch.from_numpy(synth.preprocess_u8(video)).cuda()
dec = B.Decomposition(B.make_config(rows * cols, n, C.PROX_GRAPH_LINF, rows, cols, max_iter=max_iter))
dec.set_graph_windows(None)
for rep in range(2):
    dec.load(D)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    dec.run(); st = dec.status()
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print("rep", rep, "ms", round(dt * 1e3, 2), "iters", st.iter, "conv", st.converged, flush=True)
print("svp", [l['svp'] for l in dec.log()])
