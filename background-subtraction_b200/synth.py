"""Deterministic synthetic surveillance clips for the benchmark configurations:
low-rank background (two smooth random images with slow global illumination changes, rank ~3) + a few opaque
moving rectangles (the foreground, kept as ground truth) + sensor noise, quantised to uint8."""
import numpy as np


def _smooth(rng, rows, cols, sigma):
    from scipy.ndimage import gaussian_filter
    img = gaussian_filter(rng.standard_normal((rows, cols)).astype(np.float32), sigma, mode="wrap")
    img -= img.min()
    img /= max(float(img.max()), 1e-12)
    return img


def make_clip(rows, cols, frames, seed=0, n_rect=6, noise=2.0 / 255.0, return_gt=False, period=None):
    """uint8 video [frames][cols][rows] i.e. frame-major with the reference's pixel order p = j*rows + i
    (each frame is the F-order flattening of a rows x cols image).  Returns (video_u8 [n][m], gt_mask [n][m] or None)."""
    rng = np.random.default_rng(seed)
    period = period or frames
    sigma = max(2.0, 25.0 * min(rows, cols) / 1080.0)
    b0 = 0.25 + 0.5 * _smooth(rng, rows, cols, sigma)
    b1 = _smooth(rng, rows, cols, sigma) - 0.5
    rects = []
    for _ in range(n_rect):
        rh = int(rng.integers(max(3, rows // 18), max(4, rows // 9)))
        rw = int(rng.integers(max(3, cols // 19), max(4, cols // 9)))
        y0, x0 = rng.uniform(0, rows), rng.uniform(0, cols)
        vy, vx = rng.uniform(-1, 1) * rows / 300.0, rng.uniform(-1, 1) * cols / 150.0
        off = float(rng.choice([-1.0, 1.0]) * rng.uniform(0.3, 0.5))
        rects.append((rh, rw, y0, x0, vy, vx, off))
    m = rows * cols
    video = np.empty((frames, m), dtype=np.uint8)
    gt = np.zeros((frames, m), dtype=bool) if return_gt else None
    for t in range(frames):
        img = b0 * (1.0 + 0.05 * np.sin(2 * np.pi * t / period)) + 0.03 * b1 * np.cos(4 * np.pi * t / period)
        fg = np.zeros((rows, cols), dtype=bool)
        for rh, rw, y0, x0, vy, vx, off in rects:
            y = int(y0 + vy * t) % rows
            x = int(x0 + vx * t) % cols
            ys = (np.arange(rh) + y) % rows
            xs = (np.arange(rw) + x) % cols
            img[np.ix_(ys, xs)] = np.clip(b0[np.ix_(ys, xs)] + off, 0.02, 0.98)
            fg[np.ix_(ys, xs)] = True
        img = img + noise * rng.standard_normal((rows, cols), dtype=np.float32)
        q = np.clip(np.rint(img * 255.0), 0, 255).astype(np.uint8)
        video[t] = q.ravel(order="F")
        if return_gt:
            gt[t] = fg.ravel(order="F")
    return video, gt


def preprocess_u8(video_u8):
    """Host float32 D [n][m] with the LSD() pre-processing (inexact_alm_lsd.py:211-225) applied in fp64."""
    lo, hi = float(video_u8.min()), float(video_u8.max())
    mean = float(video_u8.mean(dtype=np.float64))
    scale = 1.0 / (hi - lo) if hi > lo else 0.0
    mean_n = (mean - lo) * scale
    out = np.empty(video_u8.shape, dtype=np.float32)
    step = max(1, (1 << 24) // video_u8.shape[1])
    for f0 in range(0, video_u8.shape[0], step):
        out[f0:f0 + step] = ((video_u8[f0:f0 + step].astype(np.float64) - lo) * scale - mean_n).astype(np.float32)
    return out
