"""Oracle for the UDPHeatmap (DARK-UDP) decode stage of the ViTPose td-hm configs (TEST
INFRASTRUCTURE, see oracle/__init__.py).  SURVEY.md section 8(f) rank 3.

Restates, in numpy + cv2 (the reference itself calls cv2.GaussianBlur):

* ``get_heatmap_maximum``         mmpose/codecs/utils/post_processing.py:178-217
* ``gaussian_blur``               mmpose/codecs/utils/post_processing.py:220-249
* ``refine_keypoints_dark_udp``   mmpose/codecs/utils/refinement.py:102-160
* ``UDPHeatmap.decode``           mmpose/codecs/udp_heatmap.py:146-196 (gaussian branch, rescale :194-195)
* ``HeatmapHead.predict`` merge   mmpose/models/heads/heatmap_heads/heatmap_head.py:245-258,
  ``flip_heatmaps``               mmpose/models/utils/tta.py:35-39 (flip_mode="heatmap", shift_heatmap=False)

Pinned: ``tests/test_oracle_udp.py`` compares these functions bit-for-bit with outputs of the genuine
reference files, captured by ``oracle/gen_golden_udp.py`` (tests/golden/udp_kat.npz).
"""
from __future__ import annotations

import numpy as np

K, H, W = 17, 64, 48
COCO_FLIP_INDICES = [0, 2, 1, 4, 3, 6, 5, 8, 7, 10, 9, 12, 11, 14, 13, 16, 15]


def heatmap_maximum(heatmaps: np.ndarray):
    """post_processing.py:178-217 for (K, H, W): first arg max per map, (-1, -1) where the maximum is <= 0."""
    k, h, w = heatmaps.shape
    flat = heatmaps.reshape(k, -1)
    y, x = np.unravel_index(np.argmax(flat, axis=1), (h, w))
    locs = np.stack((x, y), axis=-1).astype(np.float32)
    vals = np.amax(flat, axis=1)
    locs[vals <= 0.0] = -1
    return locs, vals


def gaussian_kernel_1d(ksize: int = 11) -> np.ndarray:
    """The float32 taps cv2.GaussianBlur(ksize, sigma=0) uses for CV_32F images: sigma = 0.3 ((ksize - 1) / 2 - 1) + 0.8,
    exp(-x^2 / (2 sigma^2)) normalised in double, stored as float (OpenCV getGaussianKernel)."""
    sigma = 0.3 * ((ksize - 1) * 0.5 - 1) + 0.8
    x = np.arange(ksize, dtype=np.float64) - (ksize - 1) / 2
    g = np.exp(-(x * x) / (2 * sigma * sigma))
    return (g / g.sum()).astype(np.float32)


def gaussian_blur(heatmaps: np.ndarray, kernel: int = 11) -> np.ndarray:
    """post_processing.py:220-249, in place like the reference: zero-padded 11 x 11 Gaussian, rescaled to the old maximum."""
    import cv2

    border = (kernel - 1) // 2
    k, h, w = heatmaps.shape
    for i in range(k):
        origin_max = np.max(heatmaps[i])
        dr = np.zeros((h + 2 * border, w + 2 * border), dtype=np.float32)
        dr[border:-border, border:-border] = heatmaps[i].copy()
        dr = cv2.GaussianBlur(dr, (kernel, kernel), 0)
        heatmaps[i] = dr[border:-border, border:-border].copy()
        heatmaps[i] *= origin_max / (np.max(heatmaps[i]) + 1e-12)
    return heatmaps


def gaussian_blur_exact(img: np.ndarray, kernel: int = 11) -> np.ndarray:
    """cv2.GaussianBlur(zero-padded img, (kernel, kernel), 0) cropped back, WITHOUT OpenCV, bit for bit: the float
    arithmetic of OpenCV's CV_32F filter engine (AVX2 / FMA3 build) is, per output pixel,
      row filter     s = x[0] k[0];  s = fma(x[j], k[j], s), j = 1 .. kernel - 1          (general form)
      column filter  s = r[c] k[c];  s = fma(r[c + d] + r[c - d], k[c + d], s), d = 1 .. radius   (symmetric form)
    (found by trying the four combinations against cv2, tests/test_oracle_udp.py keeps it pinned).  This is the
    arithmetic csrc/decode_udp.cu implements; fma is emulated in float64 (the product of two float32 is exact there)."""
    k = gaussian_kernel_1d(kernel)
    b = (kernel - 1) // 2
    h, w = img.shape

    def fma(a, t, c):
        return (a.astype(np.float64) * np.float64(t) + c.astype(np.float64)).astype(np.float32)

    xp = np.zeros((h, w + 2 * b), np.float32)
    xp[:, b:-b] = img
    rows = (xp[:, 0:w] * k[0]).astype(np.float32)
    for j in range(1, kernel):
        rows = fma(xp[:, j:j + w], k[j], rows)
    rp = np.zeros((h + 2 * b, w), np.float32)
    rp[b:-b] = rows
    out = (rp[b:b + h] * k[b]).astype(np.float32)
    for d in range(1, b + 1):
        out = fma((rp[b + d:b + d + h] + rp[b - d:b - d + h]).astype(np.float32), k[b + d], out)
    return out


def refine_dark_udp(keypoints: np.ndarray, heatmaps: np.ndarray, blur_kernel_size: int = 11) -> np.ndarray:
    """refinement.py:102-160 (keypoints (N, K, 2) float32 in place, heatmaps (K, H, W) modified in place)."""
    n_inst, k = keypoints.shape[:2]
    h, w = heatmaps.shape[1:]
    heatmaps = gaussian_blur(heatmaps, blur_kernel_size)
    np.clip(heatmaps, 1e-3, 50.0, heatmaps)
    np.log(heatmaps, heatmaps)
    pad = np.pad(heatmaps, ((0, 0), (1, 1), (1, 1)), mode="edge").flatten()
    for n in range(n_inst):
        index = keypoints[n, :, 0] + 1 + (keypoints[n, :, 1] + 1) * (w + 2)
        index += (w + 2) * (h + 2) * np.arange(0, k)
        index = index.astype(int).reshape(-1, 1)
        i_ = pad[index]
        ix1 = pad[index + 1]
        iy1 = pad[index + w + 2]
        ix1y1 = pad[index + w + 3]
        ix1_y1_ = pad[index - w - 3]
        ix1_ = pad[index - 1]
        iy1_ = pad[index - 2 - w]
        dx = 0.5 * (ix1 - ix1_)
        dy = 0.5 * (iy1 - iy1_)
        derivative = np.concatenate([dx, dy], axis=1).reshape(k, 2, 1)
        dxx = ix1 - 2 * i_ + ix1_
        dyy = iy1 - 2 * i_ + iy1_
        dxy = 0.5 * (ix1y1 - ix1 - iy1 + i_ + i_ - ix1_ - iy1_ + ix1_y1_)
        hessian = np.concatenate([dxx, dxy, dxy, dyy], axis=1).reshape(k, 2, 2)
        hessian = np.linalg.pinv(hessian + np.finfo(np.float32).eps * np.eye(2))
        keypoints[n] -= np.einsum("imn,ink->imk", hessian, derivative).squeeze()
    return keypoints


def hessian_min_eig(heatmaps: np.ndarray, blur_kernel_size: int = 11) -> np.ndarray:
    """Per map, the smallest |eigenvalue| of the DARK-UDP Hessian at the peak (a trained sigma = 2 Gaussian gives
    ~1/8 after the blur).  Where it is tiny (flat or clipped maps) the refinement step amplifies the float rounding of
    the blur without bound - tests use this to say where a pixel tolerance is meaningful."""
    hm = heatmaps.copy()
    locs, _ = heatmap_maximum(hm)
    hm = gaussian_blur(hm, blur_kernel_size)
    np.clip(hm, 1e-3, 50.0, hm)
    np.log(hm, hm)
    pad = np.pad(hm, ((0, 0), (1, 1), (1, 1)), mode="edge")
    out = np.zeros(len(hm))
    for k, (x, y) in enumerate(locs.astype(int)):
        if x < 0:
            continue
        p = pad[k].astype(np.float64)
        x, y = x + 1, y + 1
        dxx = p[y, x + 1] - 2 * p[y, x] + p[y, x - 1]
        dyy = p[y + 1, x] - 2 * p[y, x] + p[y - 1, x]
        dxy = 0.5 * (p[y + 1, x + 1] - p[y, x + 1] - p[y + 1, x] + 2 * p[y, x] - p[y, x - 1] - p[y - 1, x] + p[y - 1, x - 1])
        out[k] = np.abs(np.linalg.eigvalsh(np.array([[dxx, dxy], [dxy, dyy]]))).min()
    return out


def udp_decode(heatmaps: np.ndarray, input_size=(192, 256), blur_kernel_size: int = 11):
    """UDPHeatmap.decode (udp_heatmap.py:146-196, gaussian): (K, H, W) float32 -> keypoints (1, K, 2) float64 in
    input-image pixels, scores (1, K) float32."""
    hm = heatmaps.copy()
    kpts, scores = heatmap_maximum(hm)
    kpts, scores = kpts[None], scores[None]
    kpts = refine_dark_udp(kpts, hm, blur_kernel_size)
    h, w = heatmaps.shape[1:]
    kpts = kpts / [w - 1, h - 1] * input_size
    return kpts, scores


def decode_instances(batch_heatmaps: np.ndarray, input_size=(192, 256)):
    """BaseHead.decode's per-instance loop (base_head.py:64-77)."""
    out = [udp_decode(hm, input_size) for hm in batch_heatmaps]
    return [o[0] for o in out], [o[1] for o in out]


def merge_flip(heatmaps: np.ndarray, heatmaps_flip: np.ndarray, flip_indices=COCO_FLIP_INDICES) -> np.ndarray:
    """heatmap_head.py:245-256 with tta.py:35-39: (hm + flip(hm_f, -1)[:, flip_indices]) * 0.5 in float32."""
    return (heatmaps + heatmaps_flip[..., ::-1][:, flip_indices]) * np.float32(0.5)


# ---- seeded input families (shared by the golden generator and the tests) ----
def gaussian_heatmaps(batch: int, seed: int = 0, sigma: float = 2.0, noise: float = 0.01):
    """Trained-ViTPose-like maps: a * exp(-r^2 / 2 sigma^2) (the training target shape, sigma = 2) + N(0, noise);
    centres uniform over the map including a margin outside it, a in [0.3, 1]."""
    rng = np.random.default_rng(seed)
    cx = rng.uniform(-2, W + 1, (batch, K, 1, 1))
    cy = rng.uniform(-2, H + 1, (batch, K, 1, 1))
    a = rng.uniform(0.3, 1.0, (batch, K, 1, 1))
    yy, xx = np.mgrid[0:H, 0:W]
    r2 = (xx[None, None] - cx) ** 2 + (yy[None, None] - cy) ** 2
    return (a * np.exp(-r2 / (2 * sigma**2)) + rng.normal(0, noise, (batch, K, H, W))).astype(np.float32)


def no_response_heatmaps():
    """(2, K, H, W): maps whose maximum is <= 0 (all-zero, all-negative, zero with negative dips) next to ordinary ones -
    get_heatmap_maximum marks them (-1, -1) and refine_keypoints_dark_udp then reads samples that wrap into the previous
    keypoint's padded plane (refinement.py:130-138 on index 0; keypoint K - 1 for keypoint 0)."""
    hm = gaussian_heatmaps(2, seed=5)
    rng = np.random.default_rng(6)
    hm[0, 0] = 0.0
    hm[0, 3] = -0.25
    hm[0, 4] = -np.abs(rng.normal(0, 0.1, (H, W))).astype(np.float32)
    hm[0, 9] = 0.0
    hm[0, 9, 5:9, 7:11] = -1.0
    hm[1, 16] = 0.0
    hm[1, 0] = -1e-3
    hm[1, 1] = 0.0  # two no-response maps in a row: the neighbour is one as well
    return hm


def special_heatmaps():
    """(1, K, H, W) edge cases: peaks at corners / borders, an exact 2-pixel tie, one dominant pixel, two blobs."""
    hm = np.full((1, K, H, W), 1e-4, np.float32)
    p = hm[0]
    yy, xx = np.mgrid[0:H, 0:W]
    for k, (y, x) in enumerate([(0, 0), (0, W - 1), (H - 1, 0), (H - 1, W - 1), (0, 20), (30, 0), (H - 1, 7), (31, W - 1), (1, 1), (H - 2, W - 2)]):
        p[k] += (0.9 * np.exp(-((xx - x) ** 2 + (yy - y) ** 2) / 8.0)).astype(np.float32)
    p[10, 20, 10] = p[10, 20, 11] = 0.5
    p[11, 40, 30] = 1.0
    p[12] += (0.8 * np.exp(-((xx - 12) ** 2 + (yy - 12) ** 2) / 8.0) + 0.6 * np.exp(-((xx - 30) ** 2 + (yy - 50) ** 2) / 8.0)).astype(np.float32)
    p[13] += (0.5 * np.exp(-((xx - 24.4) ** 2 + (yy - 31.7) ** 2) / 8.0)).astype(np.float32)
    p[14] += (60.0 * np.exp(-((xx - 5.5) ** 2 + (yy - 5.5) ** 2) / 8.0)).astype(np.float32)  # beyond the clip at 50
    p[15] += (0.002 * np.exp(-((xx - 40) ** 2 + (yy - 10) ** 2) / 8.0)).astype(np.float32)    # around the clip at 1e-3
    p[16] += (0.7 * np.exp(-((xx - 23) ** 2 + (yy - 63) ** 2) / 2.0)).astype(np.float32)
    return hm
