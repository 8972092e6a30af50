"""Oracle for the ProbMap decode stage (TEST INFRASTRUCTURE, see oracle/__init__.py).

Restates, in numpy/scipy, what the reference does on the host for every person:

* ``_prepare_oks_kernels``        mmpose/codecs/utils/post_processing.py:13-39
* ``get_heatmap_expected_value``  mmpose/codecs/utils/post_processing.py:308-381
* ``_get_subpixel_maximums``      mmpose/codecs/utils/post_processing.py:384-430
* ``ProbMap.decode``              mmpose/codecs/probmap.py:170-220 (gaussian branch)
* per-instance loop               mmpose/models/heads/base_head.py:64-77

Pinned: ``tests/test_oracle_decode.py`` compares these functions bit-for-bit with
outputs of the genuine reference file, captured by ``oracle/gen_golden.py``.
"""
from __future__ import annotations

import numpy as np
from scipy import ndimage

# COCO per-keypoint OKS sigmas, post_processing.py:16 (already divided by 100).
COCO_SIGMAS = np.array(
    [2.6, 2.5, 2.5, 3.5, 3.5, 7.9, 7.9, 7.2, 7.2, 6.2, 6.2, 10.7, 10.7, 8.7, 8.7, 8.9, 8.9]
) / 100


def oks_variances(num_keypoints: int, height: int, width: int):
    """Per-keypoint Gaussian variance ``s`` and integer radius ``ceil(3 s)``.

    post_processing.py:14,21-24: s = clip((2*sigma)^2 * sqrt(H/1.25 * W/1.25) * 2, 0.55, 3).
    """
    area = np.sqrt(height / 1.25 * width / 1.25)
    s = np.clip((COCO_SIGMAS[:num_keypoints] * 2) ** 2 * area * 2, 0.55, 3.0)
    radius = np.ceil(s * 3).astype(int)
    return s, radius


def oks_kernel_1d(s: float, radius: int) -> np.ndarray:
    """1-D factor g of the separable OKS kernel: outer(g, g) == the 2-D kernel (sum 1)."""
    d = np.arange(-radius, radius + 1, dtype=np.float64)
    g = np.exp(-(d**2) / (2 * s))
    return g / g.sum()


def oks_kernels_2d(num_keypoints: int, height: int, width: int):
    """The (1, d, d) float64 kernels exactly as post_processing.py:19-37 builds them."""
    s, radius = oks_variances(num_keypoints, height, width)
    out = []
    for k in range(num_keypoints):
        d = np.arange(2 * radius[k] + 1) - radius[k]
        gx, gy = np.meshgrid(d, d)
        dist = np.sqrt(gx**2 + gy**2)
        kern = np.exp(-(dist**2) / (2 * s[k]))
        out.append((kern / kern.sum()).reshape(1, 2 * radius[k] + 1, 2 * radius[k] + 1))
    return out


def subpixel_refine(conv: np.ndarray, locs: np.ndarray) -> np.ndarray:
    """One quadratic step around strictly-interior integer peaks (post_processing.py:384-430).

    ``conv`` (N, H, W) float32 convolved maps, ``locs`` (N, 2) float32 integer (x, y).
    """
    n, h, w = conv.shape
    x = locs[:, 0].astype(np.int32)
    y = locs[:, 1].astype(np.int32)
    out = locs.copy()
    ok = (x > 0) & (x < w - 1) & (y > 0) & (y < h - 1)
    if ok.any():
        xv, yv = x[ok], y[ok]
        c = conv[ok, yv, xv]
        r, l = conv[ok, yv, xv + 1], conv[ok, yv, xv - 1]
        d, u = conv[ok, yv + 1, xv], conv[ok, yv - 1, xv]
        dx = (r - l) / 2.0
        dy = (d - u) / 2.0
        dxx = r + l - 2 * c
        dyy = d + u - 2 * c
        dxx = np.where(dxx != 0, dxx, 1e-6)
        dyy = np.where(dyy != 0, dyy, 1e-6)
        out[ok, 0] += -dx / dxx
        out[ok, 1] += -dy / dyy
    return out


def expected_value_decode(heatmaps: np.ndarray, return_conv: bool = False):
    """One person: (K, H, W) float32 -> locs (K, 2) float32 heatmap px, vals (K,) float32.

    Follows post_processing.py:344-365 step by step, including the scipy 2-D
    convolution (float64 accumulate, float32 result, ``mode='reflect'``) so that it
    costs what the reference costs on a CPU.
    """
    assert isinstance(heatmaps, np.ndarray) and heatmaps.ndim == 3, "expects (K, H, W)"
    k_, h, w = heatmaps.shape
    kernels = oks_kernels_2d(k_, h, w)
    conv = np.zeros_like(heatmaps)
    for k in range(k_):
        conv[k] = ndimage.convolve(heatmaps[k][None], kernels[k], mode="reflect")[0]
    flat = conv.reshape(k_, h * w)
    ys, xs = np.unravel_index(np.argmax(flat, axis=1), (h, w))  # first maximum wins
    locs = np.stack((xs, ys), axis=-1).astype(np.float32)
    locs = subpixel_refine(conv, locs)
    vals = heatmaps[np.arange(k_), ys, xs]  # un-convolved map at the integer peak
    if return_conv:
        return locs, vals, conv
    return locs, vals


def separable_conv_f64(heatmaps: np.ndarray) -> np.ndarray:
    """Batched separable restatement of the OKS convolution: (..., K, H, W) float32 in,
    float64 accumulate, float32 out.  ``np.pad(mode='symmetric')`` == scipy ``reflect``.
    Used for batches too large for the scipy path; agreement with the 2-D path is
    asserted in tests/test_oracle_decode.py.
    """
    *lead, k_, h, w = heatmaps.shape
    hm = heatmaps.reshape(-1, k_, h, w).astype(np.float64)
    s, radius = oks_variances(k_, h, w)
    out = np.empty(hm.shape, dtype=np.float32)
    for k in range(k_):
        r = int(radius[k])
        g = oks_kernel_1d(s[k], r)
        p = np.pad(hm[:, k], ((0, 0), (r, r), (r, r)), mode="symmetric")
        rows = np.zeros((hm.shape[0], h + 2 * r, w))
        for i in range(2 * r + 1):
            rows += g[i] * p[:, :, i : i + w]
        acc = np.zeros((hm.shape[0], h, w))
        for i in range(2 * r + 1):
            acc += g[i] * rows[:, i : i + h, :]
        out[:, k] = acc.astype(np.float32)
    return out.reshape(*lead, k_, h, w)


def expected_value_decode_batch(heatmaps: np.ndarray):
    """(B, K, H, W) float32 -> locs (B, K, 2), vals (B, K) via the separable path."""
    b, k_, h, w = heatmaps.shape
    conv = separable_conv_f64(heatmaps).reshape(b * k_, h, w)
    idx = np.argmax(conv.reshape(b * k_, -1), axis=1)
    ys, xs = np.unravel_index(idx, (h, w))
    locs = subpixel_refine(conv, np.stack((xs, ys), -1).astype(np.float32))
    vals = heatmaps.reshape(b * k_, h, w)[np.arange(b * k_), ys, xs]
    return locs.reshape(b, k_, 2), vals.reshape(b, k_)


def probmap_decode(heatmaps: np.ndarray, input_size=(192, 256), heatmap_size=(48, 64)):
    """``ProbMap.decode`` (probmap.py:184-220): keypoints (1, K, 2) float64 in input px,
    scores (1, K) float32.  Note the reference divides by (W-1, H-1) and multiplies by
    the full input size (probmap.py:218).
    """
    w, h = heatmap_size
    locs, vals = expected_value_decode(heatmaps.copy())
    keypoints = locs[None] / [w - 1, h - 1] * input_size
    return keypoints, vals[None]


def decode_instances(batch_heatmaps: np.ndarray, input_size=(192, 256), heatmap_size=(48, 64)):
    """The reference's per-instance loop (base_head.py:64-77) over a (B, K, H, W) array."""
    kpts, scores = [], []
    for hm in batch_heatmaps:
        k, s = probmap_decode(hm, input_size, heatmap_size)
        kpts.append(k)
        scores.append(s)
    return kpts, scores


# ----------------------------------------------------------------------------------
# sparsemax + flip-TTA on raw logits (numpy twin of model_oracle's torch code; used to
# check the fused CUDA decode kernel, which takes logits, on identical bits)
# ----------------------------------------------------------------------------------


def sparsemax_rows(z: np.ndarray) -> np.ndarray:
    """Sparsemax over the last axis, float32, sort + cumsum form.

    PyPI ``sparsemax`` (call site probmap_head.py:11,251,642); Martins & Astudillo 2016
    Alg. 1.  Parity unpinned (package not in the reference tree).
    """
    z = z.astype(np.float32)
    z = z - z.max(axis=-1, keepdims=True)
    zs = -np.sort(-z, axis=-1)
    n = z.shape[-1]
    rng = np.arange(1, n + 1, dtype=np.float32)
    csum = np.cumsum(zs, axis=-1, dtype=np.float32)
    in_support = (1 + rng * zs) > csum
    k = np.max(in_support * rng, axis=-1, keepdims=True)
    tau = (np.sum(in_support * zs, axis=-1, keepdims=True, dtype=np.float32) - 1) / k
    return np.maximum(z - tau, 0).astype(np.float32)


def heatmaps_from_logits(logits: np.ndarray, temperature: float = 0.5, normalize: float = 1.0):
    """probmap_head.py:639-646: clamp(sparsemax(logits / T) * normalize, 0, 1)."""
    b, k_, h, w = logits.shape
    p = sparsemax_rows(logits.reshape(b, k_, h * w).astype(np.float32) / np.float32(temperature))
    p = np.clip(p * np.float32(normalize), 0, 1)
    return p.reshape(b, k_, h, w).astype(np.float32)


def tta_merge(p: np.ndarray, p_flip: np.ndarray, flip_indices) -> np.ndarray:
    """tta.py:35-39 + probmap_head.py:763: 0.5 * (P + mirror(Pf)[:, flip_indices])."""
    return ((p + p_flip[..., ::-1][:, list(flip_indices)]) * np.float32(0.5)).astype(np.float32)


COCO_FLIP_INDICES = [0, 2, 1, 4, 3, 6, 5, 8, 7, 10, 9, 12, 11, 14, 13, 16, 15]
