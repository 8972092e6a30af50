"""TEST INFRASTRUCTURE (see oracle/__init__.py): CPU restatement of the reference's crop front-end,
SURVEY.md section 8(f) rank 1 - from a full frame and person boxes to the uint8 model inputs.

Reference (paths relative to the reference checkout):
  * ``GetBBoxCenterScale.transform``  mmpose/datasets/transforms/common_transforms.py:62-94
  * ``bbox_xyxy2cs``                  mmpose/structures/bbox/transforms.py:44-72
  * ``TopdownAffine.transform``       mmpose/datasets/transforms/topdown_transforms.py:70-150
    (``_fix_aspect_ratio`` :55-68; the fork recomputes centre / scale from ``bbox_xyxy_wrt_input`` with
    ``input_padding`` :93-98; ``use_udp=True`` -> ``get_udp_warp_matrix``; ``cv2.warpAffine(img, warp_mat,
    (w, h), flags=cv2.INTER_LINEAR)`` :126)
  * ``get_udp_warp_matrix``           mmpose/structures/bbox/transforms.py:315-359
  * ``PackPoseInputs``                mmpose/datasets/transforms/formatting.py (HWC -> CHW, still BGR uint8)

``cv2.warpAffine`` is a third-party dependency (opencv-python, 4.13.0 in this image; the reference does not
pin it).  ``warp_affine_u8`` restates its published fixed-point algorithm (imgwarp.cpp: inverse matrix in
double, AB_BITS = 10 coordinates rounded with cvRound, 1/32-pixel fractions, 15-bit bilinear weights,
``(sum + 2^14) >> 15``, BORDER_CONSTANT 0).  PINNED: bit-for-bit against cv2.warpAffine itself on the seeded
frames / boxes of ``oracle/gen_golden_crops.py`` (tests/golden/crop_kat.npz), matrices against the genuine
``get_udp_warp_matrix`` / ``bbox_xyxy2cs`` loaded from the reference file.
"""
from __future__ import annotations

import math

import numpy as np

INPUT_SIZE = (192, 256)  # (w, h), config :48
INPUT_PADDING = 1.25     # config :6


def bbox_xyxy2cs(bbox: np.ndarray, padding: float = 1.0):
    """structures/bbox/transforms.py:61-72 (float32 in, float32 out)."""
    dim = bbox.ndim
    if dim == 1:
        bbox = bbox[None, :]
    scale = (bbox[..., 2:] - bbox[..., :2]) * padding
    center = (bbox[..., 2:] + bbox[..., :2]) * 0.5
    if dim == 1:
        center, scale = center[0], scale[0]
    return center, scale


def fix_aspect_ratio(bbox_scale: np.ndarray, aspect_ratio: float) -> np.ndarray:
    """TopdownAffine._fix_aspect_ratio, topdown_transforms.py:55-68."""
    w, h = np.hsplit(bbox_scale, [1])
    return np.where(w > h * aspect_ratio, np.hstack([w, w / aspect_ratio]), np.hstack([h * aspect_ratio, h]))


def udp_warp_matrix(center: np.ndarray, scale: np.ndarray, rot: float, output_size) -> np.ndarray:
    """get_udp_warp_matrix, structures/bbox/transforms.py:343-359 (same expression order / dtypes)."""
    input_size = center * 2
    rot_rad = np.deg2rad(rot)
    warp_mat = np.zeros((2, 3), dtype=np.float32)
    scale_x = (output_size[0] - 1) / scale[0]
    scale_y = (output_size[1] - 1) / scale[1]
    warp_mat[0, 0] = math.cos(rot_rad) * scale_x
    warp_mat[0, 1] = -math.sin(rot_rad) * scale_x
    warp_mat[0, 2] = scale_x * (-0.5 * input_size[0] * math.cos(rot_rad) + 0.5 * input_size[1] * math.sin(rot_rad) + 0.5 * scale[0])
    warp_mat[1, 0] = math.sin(rot_rad) * scale_y
    warp_mat[1, 1] = math.cos(rot_rad) * scale_y
    warp_mat[1, 2] = scale_y * (-0.5 * input_size[0] * math.sin(rot_rad) - 0.5 * input_size[1] * math.cos(rot_rad) + 0.5 * scale[1])
    return warp_mat


def topdown_geometry(bbox_xyxy: np.ndarray, input_size=INPUT_SIZE, padding: float = INPUT_PADDING):
    """One box (4,) float32 -> (center (2,), scale (2,), warp matrix (2, 3) float32) exactly as
    GetBBoxCenterScale + TopdownAffine(use_udp=True) produce them (topdown_transforms.py:93-118)."""
    w, h = input_size
    c, s = bbox_xyxy2cs(np.asarray(bbox_xyxy)[None], padding=padding)
    s = fix_aspect_ratio(s.reshape(1, 2), aspect_ratio=w / h)
    center, scale = c.reshape(1, 2)[0], s[0]
    return center, scale, udp_warp_matrix(center, scale, 0.0, output_size=(w, h))


def warp_affine_u8(img: np.ndarray, m: np.ndarray, out_wh) -> np.ndarray:
    """cv2.warpAffine(img, m, out_wh, flags=cv2.INTER_LINEAR) for uint8 HWC images, BORDER_CONSTANT 0."""
    w, h = out_wh
    hh, ww = img.shape[:2]
    m = np.asarray(m, np.float64).copy().reshape(6)
    d = m[0] * m[4] - m[1] * m[3]
    d = 1.0 / d if d != 0 else 0.0
    a11, a22 = m[4] * d, m[0] * d
    m[0] = a11; m[1] *= -d; m[3] *= -d; m[4] = a22
    b1 = -m[0] * m[2] - m[1] * m[5]
    b2 = -m[3] * m[2] - m[4] * m[5]
    m[2] = b1; m[5] = b2
    xs, ys = np.arange(w, dtype=np.float64), np.arange(h, dtype=np.float64)
    adelta = np.rint(m[0] * xs * 1024).astype(np.int64)  # cvRound: round half to even
    bdelta = np.rint(m[3] * xs * 1024).astype(np.int64)
    x0 = np.rint((m[1] * ys + m[2]) * 1024).astype(np.int64) + 16  # round_delta = AB_SCALE / INTER_TAB_SIZE / 2
    y0 = np.rint((m[4] * ys + m[5]) * 1024).astype(np.int64) + 16
    xq = (x0[:, None] + adelta[None, :]) >> 5
    yq = (y0[:, None] + bdelta[None, :]) >> 5
    sx, sy = np.clip(xq >> 5, -32768, 32767), np.clip(yq >> 5, -32768, 32767)
    fx, fy = xq & 31, yq & 31
    src = img.astype(np.int64).reshape(hh, ww, -1)

    def tap(yy, xx):
        ok = (yy >= 0) & (yy < hh) & (xx >= 0) & (xx < ww)
        return src[np.clip(yy, 0, hh - 1), np.clip(xx, 0, ww - 1)] * ok[..., None]

    acc = (tap(sy, sx) * ((32 - fx) * (32 - fy) * 32)[..., None] + tap(sy, sx + 1) * (fx * (32 - fy) * 32)[..., None]
           + tap(sy + 1, sx) * ((32 - fx) * fy * 32)[..., None] + tap(sy + 1, sx + 1) * (fx * fy * 32)[..., None])
    return ((acc + (1 << 14)) >> 15).astype(np.uint8).reshape(h, w, *img.shape[2:])


def topdown_crops(frame_bgr: np.ndarray, bboxes_xyxy: np.ndarray, input_size=INPUT_SIZE, padding: float = INPUT_PADDING):
    """Frame (H, W, 3) uint8 BGR + boxes (N, 4) -> (crops (N, 3, h, w) uint8 BGR CHW, centers (N, 2), scales (N, 2),
    matrices (N, 2, 3)): the val pipeline of the config (:106-111) up to PackPoseInputs."""
    crops, cs, ss, ms = [], [], [], []
    for bbox in np.asarray(bboxes_xyxy, np.float32):
        c, s, m = topdown_geometry(bbox, input_size, padding)
        crops.append(warp_affine_u8(frame_bgr, m, input_size).transpose(2, 0, 1))
        cs.append(c); ss.append(s); ms.append(m)
    return np.stack(crops), np.stack(cs), np.stack(ss), np.stack(ms)


def synthetic_frame(seed: int, height: int = 480, width: int = 640) -> np.ndarray:
    """Smooth blobs + noise, uint8 BGR (H, W, 3)."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:height, 0:width]
    img = np.zeros((height, width, 3), np.float64)
    for c in range(3):
        for _ in range(6):
            cx, cy, s, a = rng.uniform(0, width), rng.uniform(0, height), rng.uniform(20, 120), rng.uniform(40, 160)
            img[..., c] += a * np.exp(-((xx - cx) ** 2 + (yy - cy) ** 2) / (2 * s * s))
    img += rng.uniform(0, 40, img.shape)
    return np.clip(img, 0, 255).astype(np.uint8)


def synthetic_boxes(seed: int, n: int, height: int = 480, width: int = 640) -> np.ndarray:
    """Person-like boxes incl. ones that leave the frame (border handling) and tiny / huge ones."""
    rng = np.random.default_rng(seed)
    x1 = rng.uniform(-60, width * 0.8, n)
    y1 = rng.uniform(-60, height * 0.8, n)
    bw = rng.uniform(8, width * 0.9, n)
    bh = rng.uniform(8, height * 1.1, n)
    return np.stack([x1, y1, x1 + bw, y1 + bh], 1).astype(np.float32)
