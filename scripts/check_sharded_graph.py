"""Overlapping-window LSD sharded over GPUs (frame re-sharding around the prox, dist.ShardedLSD._prox_on_frames) against the
single-GPU solve of the same clip (run on 2 B200s in round 2: profiles/r2x_check_sharded_2gpu.log).  The all-to-all choreography
is also covered by tests/test_dist_gloo.py on CPU ranks and the split entry points by tests/test_gpu_parity.py on one GPU.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 scripts/check_sharded_graph.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from background_subtraction_b200 import dist as bdist, synth  # noqa: E402


class Solo:
    world, rank = 1, 0

    def all_reduce_sum(self, t):
        pass

    def all_reduce_max(self, t):
        pass


def solve(D, rows, cols, frames, c0, c1, comm):
    cl = c1 - c0
    shard = np.ascontiguousarray(D.reshape(frames, cols, rows)[:, c0:c1, :].reshape(frames, rows * cl))
    s = bdist.CudaStepSolver(rows, cl, frames, rows * cols, graph_cols=cols)
    s.load(shard)
    drv = bdist.ShardedLSD(s, comm)
    drv.solve()
    drv.finish(2.0, want_mask=False)
    torch.cuda.synchronize()
    return s.status(), s.dec.download('S', dtype=np.float32), [l['svp'] for l in s.dec.log()]


def main():
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    rows, cols, frames = 96, 129, 30
    video, _ = synth.make_clip(rows, cols, frames, seed=5, n_rect=3)
    D = synth.preprocess_u8(video)
    c0, c1 = bdist.shard_columns(cols, world, rank)
    st, S, svp = solve(D, rows, cols, frames, c0, c1, bdist.TorchComm())
    ok = True
    if rank == 0:
        st1, S1, svp1 = solve(D, rows, cols, frames, 0, cols, Solo())
        S1 = S1.reshape(cols, rows, frames)[c0:c1].reshape(-1, frames)
        rel = float(np.linalg.norm(S - S1) / max(np.linalg.norm(S1), 1e-30))
        print("graph %dx%dx%d sharded: iter=%d err=%.3e svp=%s | single: iter=%d err=%.3e | relF(S) on rank 0 columns %.3e"
              % (rows, cols, frames, st.iter, st.err, svp, st1.iter, st1.err, rel), flush=True)
        ok = st.iter == st1.iter and svp == svp1 and rel <= 1e-4 and bool(st.converged)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    return 0 if int(flag.item()) == 1 else 1


if __name__ == "__main__":
    sys.exit(main())
