"""Seeded inputs shared by the CPU (oracle) and GPU tests of the stages around the decomposition."""
import numpy as np


def random_masks(seed, h, w, t, density=(0.2, 0.6)):
    rng = np.random.default_rng(seed)
    return np.stack([rng.random((h, w)) < rng.uniform(*density) for _ in range(t)], axis=2)


def blob_video(seed, h, w, t, n_blobs=5):
    """Binary video with a few moving rectangles / rings (rings enclose a small blob: bbox-nested components), speckle noise,
    and a saliency cube that is high on some of the objects -- the shape of what run_motion_saliency_check sees."""
    rng = np.random.default_rng(seed)
    mask = np.zeros((h, w, t), dtype=bool)
    cube = rng.random((h, w, t)) * 1e-6
    objs = []
    for b in range(n_blobs):
        bh, bw = rng.integers(h // 8, h // 3), rng.integers(w // 8, w // 3)
        objs.append((rng.integers(0, h - bh), rng.integers(0, w - bw), bh, bw, rng.integers(-1, 2), rng.integers(-1, 2), b % 2 == 0,
                     rng.uniform(0.5, 4.0) if b % 3 else 0.05))
    for f in range(t):
        for (i0, j0, bh, bw, vi, vj, ring, sal) in objs:
            i = int(np.clip(i0 + vi * f, 0, h - bh)); j = int(np.clip(j0 + vj * f, 0, w - bw))
            if ring and bh >= 7 and bw >= 7:
                mask[i:i + bh, j:j + bw, f] = True
                mask[i + 1:i + bh - 1, j + 1:j + bw - 1, f] = False
                mask[i + 3:i + bh - 3, j + 3:j + bw - 3, f] = True            # nested inside the ring's bounding box
            else:
                mask[i:i + bh, j:j + bw, f] = True
            cube[i:i + bh, j:j + bw, f] += sal * 1e-5 * (1 + 0.1 * rng.random((bh, bw)))
        mask[:, :, f] |= rng.random((h, w)) < 0.01                            # speckle: many tiny components
    return mask, cube / cube.sum()
