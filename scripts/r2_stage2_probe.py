"""Stage-2 timing probe: executeSaliencyRPCA + computeSCube on a 320 x 240 x 200 clip (BASELINE config 5 clip size)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import background_subtraction_b200 as B
from background_subtraction_b200 import flow
from test_oracle_flow import saliency_slices
video = torch.from_numpy(saliency_slices(11, 240, 320, 200)).cuda()
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    xl, xs, yl, ys = B.executeSaliencyRPCA(video, 1)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    cube = B.computeSCube(xs, ys, return_device=True)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    print("rep", rep, "executeSaliencyRPCA ms %.2f (560 slices)" % ((t1 - t0) * 1e3), "computeSCube ms %.2f" % ((t2 - t1) * 1e3), flush=True)
xt = video.permute(2, 1, 0).contiguous()
L, S, info = flow.inexact_alm_rpca_batch(xt, tol_l1=xt.shape[1] * xt.shape[2] * 1e-4, return_info=True)
print("xt slices", xt.shape, "iters min/max", info["iters"].min(), info["iters"].max(), "rank", np.unique(info["rank"]))
