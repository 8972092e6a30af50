"""``merge_data_samples`` / ``revert_heatmap`` of mmpose/structures/utils.py (:19-175) for the top-down demo callers
(``demo/topdown_demo_with_mmdet.py`` merges the per-person samples of one frame before visualising them): the geometry
stays on the host exactly as the reference computes it, the per-person ``cv2.warpAffine`` of the (H, W, K) heatmaps and the
``np.max`` over persons run as one CUDA kernel (``pp_revert_heatmaps``)."""
from __future__ import annotations

import warnings
from typing import List

import numpy as np
import torch

from .. import ops
from .structures import InstanceData, PixelData, PoseDataSample


def _get_affine_transform(src: np.ndarray, dst: np.ndarray) -> np.ndarray:
    """cv2.getAffineTransform (float32 points in, float64 2 x 3 out); numpy solve of the same 6 x 6 system without OpenCV."""
    try:
        import cv2
        return cv2.getAffineTransform(np.float32(src), np.float32(dst))
    except ImportError:
        a = np.zeros((6, 6))
        b = np.zeros(6)
        for i in range(3):
            a[i, 0:2], a[i, 2] = src[i], 1
            a[i + 3, 3:5], a[i + 3, 5] = src[i], 1
            b[i], b[i + 3] = dst[i, 0], dst[i, 1]
        return np.linalg.solve(a, b).reshape(2, 3)


def get_warp_matrix(center, scale, rot, output_size, shift=(0.0, 0.0), inv=False, fix_aspect_ratio=True) -> np.ndarray:
    """mmpose/structures/bbox/transforms.py:362-425."""
    assert len(center) == 2 and len(scale) == 2 and len(output_size) == 2 and len(shift) == 2
    shift = np.array(shift)
    src_w, src_h = scale[:2]
    dst_w, dst_h = output_size[:2]
    rot_rad = np.deg2rad(rot)
    sn, cs = np.sin(rot_rad), np.cos(rot_rad)
    rot_mat = np.array([[cs, -sn], [sn, cs]])
    src_dir = rot_mat @ np.array([src_w * -0.5, 0.0])
    dst_dir = np.array([dst_w * -0.5, 0.0])
    src = np.zeros((3, 2), dtype=np.float32)
    src[0, :] = center + scale * shift
    src[1, :] = center + src_dir + scale * shift
    dst = np.zeros((3, 2), dtype=np.float32)
    dst[0, :] = [dst_w * 0.5, dst_h * 0.5]
    dst[1, :] = np.array([dst_w * 0.5, dst_h * 0.5]) + dst_dir

    def third(a, b):
        d = a - b
        return b + np.r_[-d[1], d[0]]

    if fix_aspect_ratio:
        src[2, :] = third(src[0, :], src[1, :])
        dst[2, :] = third(dst[0, :], dst[1, :])
    else:
        src[2, :] = center + rot_mat @ np.array([0.0, src_h * -0.5]) + scale * shift
        dst[2, :] = np.array([dst_w * 0.5, dst_h * 0.5]) + np.array([0.0, dst_h * -0.5])
    return _get_affine_transform(dst, src) if inv else _get_affine_transform(src, dst)


def flip_heatmaps(heatmaps: torch.Tensor, flip_indices=None, flip_mode: str = "heatmap", shift_heatmap: bool = True):
    """mmpose/models/utils/tta.py:9-67 on device tensors, every ``flip_mode``.  The fused kernels cover
    ``flip_mode="heatmap", shift_heatmap=False`` (the shipped test_cfg); any other setting takes this tensor path."""
    if flip_mode == "heatmap":
        heatmaps = heatmaps.flip(-1)
        if flip_indices is not None:
            assert len(flip_indices) == heatmaps.shape[1]
            heatmaps = heatmaps[:, list(flip_indices)]
    elif flip_mode in ("udp_combined", "offset"):
        group, neg = (3, 1) if flip_mode == "udp_combined" else (2, 0)
        b, c, h, w = heatmaps.shape
        heatmaps = heatmaps.reshape(b, c // group, -1, h, w).flip(-1)
        if flip_indices is not None:
            assert len(flip_indices) == c // group
            heatmaps = heatmaps[:, list(flip_indices)]
        heatmaps = heatmaps.clone()
        heatmaps[:, :, neg] = -heatmaps[:, :, neg]
        heatmaps = heatmaps.reshape(b, c, h, w)
    else:
        raise ValueError(f'Invalid flip_mode value "{flip_mode}"')
    if shift_heatmap:
        heatmaps = heatmaps.clone()
        heatmaps[..., 1:] = heatmaps[..., :-1].clone()
    return heatmaps


def _as_cuda(heatmap) -> torch.Tensor:
    t = heatmap if torch.is_tensor(heatmap) else torch.from_numpy(np.ascontiguousarray(heatmap, np.float32))
    return t.detach().float().cuda().contiguous() if not t.is_cuda else t.detach().float().contiguous()


def revert_heatmap(heatmap, input_center, input_scale, img_shape) -> np.ndarray:
    """structures/utils.py:146-175: (K, H, W) heatmap -> (K, img_h, img_w) numpy, warped back onto the image."""
    t = _as_cuda(heatmap)
    assert t.dim() == 3, "expects a (K, H, W) heatmap"
    hm_h, hm_w = t.shape[1:]
    mat = get_warp_matrix(np.asarray(input_center).reshape((2,)), np.asarray(input_scale).reshape((2,)), rot=0,
                          output_size=(hm_w, hm_h), inv=True)
    return ops.revert_heatmaps(t[None], mat[None], img_shape).cpu().numpy()


def merge_data_samples(data_samples: List[PoseDataSample]) -> PoseDataSample:
    """structures/utils.py:19-127 for the fields of this path: metainfo of the first sample with stacked
    ``input_center`` / ``input_scale``, concatenated ``gt_instances`` / ``pred_instances``, and ``pred_fields.heatmaps`` =
    the max over persons of the heatmaps warped into the padded image (:51-118)."""
    if not isinstance(data_samples, (list, tuple)) or not all(isinstance(d, PoseDataSample) for d in data_samples):
        raise ValueError("Invalid input type, should be a list of " ":obj:`PoseDataSample`")
    if len(data_samples) == 0:
        warnings.warn("Try to merge an empty list of data samples.")
        return PoseDataSample()
    meta = dict(data_samples[0].metainfo)
    centers = [np.asarray(d.metainfo["input_center"]) for d in data_samples]
    scales = [np.asarray(d.metainfo["input_scale"]) for d in data_samples]
    meta["input_center"], meta["input_scale"] = np.array(centers), np.array(scales)
    merged = PoseDataSample(metainfo=meta)

    def cat(name):
        parts = [getattr(d, name) for d in data_samples]
        keys = parts[0].keys()
        out = InstanceData()
        for k in keys:
            vals = [p[k] for p in parts]
            out.set_field(torch.cat(vals, 0) if torch.is_tensor(vals[0]) else np.concatenate(vals, 0), k)
        return out

    if "gt_instances" in data_samples[0]:
        merged.gt_instances = cat("gt_instances")
    if "pred_instances" in data_samples[0]:
        merged.pred_instances = cat("pred_instances")
    if "pred_fields" in data_samples[0] and "heatmaps" in data_samples[0].pred_fields:
        ori_shape = data_samples[0].metainfo["ori_shape"]
        pad = np.array([0, 0, 0, 0])
        for c, s in zip(centers, scales):  # utils.py:71-88: [left, top, right, bottom]
            pad = np.maximum(pad, [int(max(s[0] / 2 - c[0] + 10, 0)), int(max(s[1] / 2 - c[1] + 10, 0)),
                                   int(max(c[0] + s[0] / 2 - ori_shape[1] + 10, 0)),
                                   int(max(c[1] + s[1] / 2 - ori_shape[0] + 10, 0))])
        padded_shape = (ori_shape[0] + pad[1] + pad[3], ori_shape[1] + pad[0] + pad[2])
        hms = torch.stack([_as_cuda(d.pred_fields.heatmaps) for d in data_samples])
        hm_h, hm_w = hms.shape[2:]
        mats = np.stack([get_warp_matrix((c + np.array([pad[0], pad[1]])).reshape((2,)), s.reshape((2,)), rot=0,
                                         output_size=(hm_w, hm_h), inv=True) for c, s in zip(centers, scales)])
        fields = PixelData()
        fields.set_field(ops.revert_heatmaps(hms, mats, padded_shape).cpu().numpy(), "heatmaps")
        merged.pred_fields = fields
    return merged
