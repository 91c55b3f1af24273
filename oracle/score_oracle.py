"""oracle/score_oracle.py -- TEST INFRASTRUCTURE ONLY (never imported by the product package).

Vectorised restatement of the acceptance metric of the reference, /root/reference/compute_score.py:24-100
(true_positive / false_positive / false_negative per frame inside the ROI, CDnet label set {0, 50, 255}, then
precision / recall / F per frame with the reference's 0/0 -> 1 conventions and float32 results).  Used to check that
the CUDA masks and the reference-algorithm masks score the same (north_star: F-measure within 0.001).
Pinned against the reference functions themselves in tests/test_oracle.py when /root/reference is present.
"""
import numpy as np

KNOWN_VALUES = (0, 50, 255)          # compute_score.py:5


def confusion_counts(sparse_mat, gt_mat, roi_mask):
    """(tp, fp, fn) int arrays over frames; sparse_mat bool [h,w,t], gt_mat uint8 [h,w,t], roi_mask uint8 [h,w]."""
    roi = (np.asarray(roi_mask) == 255)[:, :, None]
    gt = np.asarray(gt_mat)
    area = np.isin(gt, KNOWN_VALUES) & roi
    obj = area & (gt == 255)
    bg = area & (gt != 255)
    sp = np.asarray(sparse_mat).astype(bool)
    tp = (obj & sp).sum(axis=(0, 1))
    fp = (bg & sp).sum(axis=(0, 1))
    fn = (obj & ~sp).sum(axis=(0, 1))
    return tp, fp, fn


def _ratio(num, other):
    out = np.ones(num.shape, dtype=np.float32)                 # 0/0 -> 1 (compute_score.py:68-71, 81-84)
    nz = ~((num == 0) & (other == 0))
    out[nz] = (num[nz] / (num[nz] + other[nz])).astype(np.float32)
    return out


def compute_precision(tp, fp):
    return _ratio(np.asarray(tp), np.asarray(fp))


def compute_recall(tp, fn):
    return _ratio(np.asarray(tp), np.asarray(fn))


def compute_fscore(tp, fp, fn):
    rc, pr = compute_recall(tp, fn), compute_precision(tp, fp)
    out = np.ones(rc.shape, dtype=np.float32)                  # rc = pr = 0 -> 1 (compute_score.py:95-96)
    nz = ~((rc == 0) & (pr == 0))
    out[nz] = (2 * rc[nz] * pr[nz] / (rc[nz] + pr[nz])).astype(np.float32)
    return out


def mean_fscore(mask_cube, gt_cube_bool):
    """Mean per-frame F of a boolean mask cube against a boolean ground truth (ROI = whole frame)."""
    gt = np.where(gt_cube_bool, 255, 0).astype(np.uint8)
    roi = np.full(gt.shape[:2], 255, dtype=np.uint8)
    tp, fp, fn = confusion_counts(mask_cube, gt, roi)
    return float(np.mean(compute_fscore(tp, fp, fn)))
