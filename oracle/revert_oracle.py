"""Oracle for the heatmap read-back path (TEST INFRASTRUCTURE, see oracle/__init__.py): SURVEY.md 8(f) rank 4.

Restates, with numpy + cv2 (the reference itself calls cv2.getAffineTransform / cv2.warpAffine):

* ``get_warp_matrix``     mmpose/structures/bbox/transforms.py:362-425
* ``revert_heatmap``      mmpose/structures/utils.py:146-175
* ``merge_data_samples``  mmpose/structures/utils.py:51-118, the ``pred_fields.heatmaps`` part: per-image padding so that every
  person's padded box fits, per-person inverse warp of the (H, W, K) heatmap into the padded image, element-wise max.

Pinned: ``tests/test_oracle_revert.py`` compares bit-for-bit with ``oracle/gen_golden_revert.py``'s run of the genuine
``get_warp_matrix`` (loaded from the reference file) followed by the verbatim ``utils.py`` lines.
"""
from __future__ import annotations

import numpy as np


def _rotate_point(pt, angle_rad):  # transforms.py:475-488
    sn, cs = np.sin(angle_rad), np.cos(angle_rad)
    return np.array([[cs, -sn], [sn, cs]]) @ pt


def _get_3rd_point(a, b):  # transforms.py:491-507
    direction = a - b
    return b + np.r_[-direction[1], direction[0]]


def get_warp_matrix(center, scale, rot, output_size, shift=(0.0, 0.0), inv=False, fix_aspect_ratio=True):
    """transforms.py:362-425."""
    import cv2

    shift = np.array(shift)
    src_w, src_h = scale[:2]
    dst_w, dst_h = output_size[:2]
    rot_rad = np.deg2rad(rot)
    src_dir = _rotate_point(np.array([src_w * -0.5, 0.0]), rot_rad)
    dst_dir = np.array([dst_w * -0.5, 0.0])
    src = np.zeros((3, 2), dtype=np.float32)
    src[0, :] = center + scale * shift
    src[1, :] = center + src_dir + scale * shift
    dst = np.zeros((3, 2), dtype=np.float32)
    dst[0, :] = [dst_w * 0.5, dst_h * 0.5]
    dst[1, :] = np.array([dst_w * 0.5, dst_h * 0.5]) + dst_dir
    if fix_aspect_ratio:
        src[2, :] = _get_3rd_point(src[0, :], src[1, :])
        dst[2, :] = _get_3rd_point(dst[0, :], dst[1, :])
    else:
        src_dir_2 = _rotate_point(np.array([0.0, src_h * -0.5]), rot_rad)
        dst_dir_2 = np.array([0.0, dst_h * -0.5])
        src[2, :] = center + src_dir_2 + scale * shift
        dst[2, :] = np.array([dst_w * 0.5, dst_h * 0.5]) + dst_dir_2
    if inv:
        return cv2.getAffineTransform(np.float32(dst), np.float32(src))
    return cv2.getAffineTransform(np.float32(src), np.float32(dst))


def revert_heatmap(heatmap: np.ndarray, input_center, input_scale, img_shape) -> np.ndarray:
    """utils.py:146-175: (K, H, W) float32 -> (K, img_h, img_w)."""
    import cv2

    hm = heatmap.transpose(1, 2, 0)
    hm_h, hm_w = hm.shape[:2]
    img_h, img_w = img_shape
    warp_mat = get_warp_matrix(input_center.reshape((2,)), input_scale.reshape((2,)), rot=0, output_size=(hm_w, hm_h), inv=True)
    hm = cv2.warpAffine(hm, warp_mat, (img_w, img_h), flags=cv2.INTER_LINEAR)
    return hm.transpose(2, 0, 1)


def image_padding(centers, scales, ori_shape):
    """utils.py:71-88: [left, top, right, bottom] so that every person's box (+10 px) fits."""
    pad = np.array([0, 0, 0, 0])
    for c, s in zip(centers, scales):
        pad = np.maximum(pad, [int(max(s[0] / 2 - c[0] + 10, 0)), int(max(s[1] / 2 - c[1] + 10, 0)),
                               int(max(c[0] + s[0] / 2 - ori_shape[1] + 10, 0)), int(max(c[1] + s[1] / 2 - ori_shape[0] + 10, 0))])
    return pad


def merged_padded_heatmaps(heatmaps, centers, scales, ori_shape):
    """utils.py:69-118: what ``merge_data_samples`` stores in ``pred_fields.heatmaps``."""
    pad = image_padding(centers, scales, ori_shape)
    shape = (ori_shape[0] + pad[1] + pad[3], ori_shape[1] + pad[0] + pad[2])
    out = [revert_heatmap(hm, c + np.array([pad[0], pad[1]]), s, shape) for hm, c, s in zip(heatmaps, centers, scales)]
    return np.max(out, axis=0), pad


def synthetic_people(seed: int, n: int, img_h: int, img_w: int):
    """Seeded person boxes (centre, aspect-fixed scale, some reaching over the image border) and sparse ProbMap-like
    heatmaps (K, 64, 48) with a few positive pixels plus a smooth low-amplitude background."""
    rng = np.random.default_rng(seed)
    centers = np.stack([rng.uniform(0, img_w, n), rng.uniform(0, img_h, n)], -1).astype(np.float32)
    hh = rng.uniform(0.15, 0.9, n) * img_h
    scales = np.stack([hh * 0.75, hh], -1).astype(np.float32)
    hms = np.zeros((n, 17, 64, 48), np.float32)
    yy, xx = np.mgrid[0:64, 0:48]
    for p in range(n):
        for k in range(17):
            cx, cy = rng.uniform(0, 47), rng.uniform(0, 63)
            hms[p, k] = (0.8 * np.exp(-((xx - cx) ** 2 + (yy - cy) ** 2) / 6.0)).astype(np.float32)
            hms[p, k][hms[p, k] < 0.05] = 0
    return hms, centers, scales
