"""Sharded solve (one process per GPU, NCCL) against the single-GPU solve of the same clip.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 scripts/check_sharded.py

Every rank solves its column shard through ShardedLSD; rank 0 also solves the whole clip alone and compares iteration
count, stop residual, rank and the foreground mask of its own columns.  Exit code 0 = agreement.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from background_subtraction_b200 import dist as bdist, synth  # noqa: E402


def solve(rows, cols, c0, c1, frames, m_global, D_full, comm):
    cl = c1 - c0
    shard = np.ascontiguousarray(D_full.reshape(frames, cols, rows)[:, c0:c1, :].reshape(frames, rows * cl))
    s = bdist.CudaStepSolver(rows, cl, frames, m_global)
    s.load(shard)
    drv = bdist.ShardedLSD(s, comm)                      # a fence per iteration is mandatory for the device solver (default)
    drv.solve()
    mask = drv.finish(2.0, want_mask=True)
    torch.cuda.synchronize()
    st = s.status()
    return st, mask.cpu().numpy().reshape(frames, cl, rows), s.dec.debug_info()


def main():
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    class Solo:
        world, rank = 1, 0

        def all_reduce_sum(self, t):
            pass

        def all_reduce_max(self, t):
            pass

    ok = True
    # the second clip is tiny: the host, not the device, paces the loop, so the ranks see the device-written stop flag at
    # different loop indices unless the stop decision is tied to the fenced iteration (dist.py)
    for rows, cols, frames, reps in ((240, 321, 96, 1), (24, 33, 20, 8)):
        video, _ = synth.make_clip(rows, cols, frames, seed=11, n_rect=3)
        D = synth.preprocess_u8(video)                                   # float32 [frames][m]
        m = rows * cols
        c0, c1 = bdist.shard_columns(cols, world, rank)
        for _rep in range(reps):
            st, mask, info = solve(rows, cols, c0, c1, frames, m, D, bdist.TorchComm())
        if rank == 0:
            st1, mask1, _ = solve(rows, cols, 0, cols, frames, m, D, Solo())
            same = float((mask1[:, c0:c1, :] == mask).mean())
            print("%dx%dx%d sharded: iter=%d err=%.3e svp=%d use_i8=%d | single: iter=%d err=%.3e svp=%d | mask agreement on rank 0 columns %.6f"
                  % (rows, cols, frames, st.iter, st.err, st.svp, info["use_i8"], st1.iter, st1.err, st1.svp, same), flush=True)
            ok = ok and st.iter == st1.iter and st.svp == st1.svp and same >= 0.999 and bool(st.converged)
    # group-sparse solver (l2 blocks per frame): blocks straddle the shard border, one frame has none (the whole frame is then the
    # complement group); the per-(frame, group) sums of squares are all-reduced between the two halves of the shrink pass
    rows, cols, frames = 120, 161, 64
    video, _ = synth.make_clip(rows, cols, frames, seed=7, n_rect=3)
    D = synth.preprocess_u8(video)
    m = rows * cols
    labels = np.zeros((frames, cols, rows), dtype=np.uint8)
    ptr, lam = [0], []
    for f in range(frames):
        if f != 5:
            c = cols // 2 - 6 + (f % 5)
            labels[f, c:c + 12, 20:50] = 1
            labels[f, 10:30, 70:100] = 2
            lam += [0.02 + 0.0001 * f, 0.03]
        ptr.append(len(lam))
    ptr, lam = np.asarray(ptr, dtype=np.int32), np.asarray(lam + [0.0])
    c0, c1 = bdist.shard_columns(cols, world, rank)

    def solve_gs(a0, a1, comm):
        cl = a1 - a0
        shard = np.ascontiguousarray(D.reshape(frames, cols, rows)[:, a0:a1, :].reshape(frames, rows * cl))
        s = bdist.CudaStepSolver(rows, cl, frames, m, blocks=(np.ascontiguousarray(labels[:, a0:a1, :]).reshape(frames, rows * cl), ptr, lam))
        s.load(shard)
        drv = bdist.ShardedLSD(s, comm)
        drv.solve()
        drv.finish(2.0, want_mask=False)
        torch.cuda.synchronize()
        return s.status(), s.dec.download('S', dtype=np.float32)
    st, S = solve_gs(c0, c1, bdist.TorchComm())
    if rank == 0:
        st1, S1 = solve_gs(0, cols, Solo())
        S1 = S1.reshape(cols, rows, frames)[c0:c1].reshape(-1, frames)
        rel = float(np.linalg.norm(S - S1) / max(np.linalg.norm(S1), 1e-30))
        print("group-sparse %dx%dx%d sharded: iter=%d err=%.3e svp=%d done=%d | single: iter=%d err=%.3e svp=%d | relF(S) on rank 0 columns %.3e"
              % (rows, cols, frames, st.iter, st.err, st.svp, st.done, st1.iter, st1.err, st1.svp, rel), flush=True)
        ok = ok and st.iter == st1.iter and st.svp == st1.svp and st.done == st1.done and rel <= 1e-4
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    return 0 if int(flag.item()) == 1 else 1


if __name__ == "__main__":
    sys.exit(main())
