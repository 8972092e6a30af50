"""``inference_topdown`` for the B200 path: the caller contract of mmpose/apis/inference.py:133-200 with
the per-person CPU pipeline (GetBBoxCenterScale -> TopdownAffine -> PackPoseInputs -> pseudo_collate ->
H2D) replaced by ONE frame upload and one GPU warp launch (``pp_crop_warp``), then ``pp_engine_infer``.

Host-side geometry mirrors the reference line by line so that centres, scales and matrices are the
same float32 values:
  * ``bbox_xyxy2cs``            mmpose/structures/bbox/transforms.py:44-72
  * ``GetBBoxCenterScale``      mmpose/datasets/transforms/common_transforms.py:62-94
  * ``TopdownAffine``           mmpose/datasets/transforms/topdown_transforms.py:70-150 (``use_udp=True``)
  * ``get_udp_warp_matrix``     mmpose/structures/bbox/transforms.py:315-359
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence, Tuple, Union

import numpy as np
import torch

from .structures import InstanceData, PoseDataSample


def bbox_xywh2xyxy(bbox_xywh: np.ndarray) -> np.ndarray:
    """structures/bbox/transforms.py:26-41."""
    bbox_xyxy = bbox_xywh.copy()
    bbox_xyxy[:, 2] = bbox_xyxy[:, 2] + bbox_xyxy[:, 0]
    bbox_xyxy[:, 3] = bbox_xyxy[:, 3] + bbox_xyxy[:, 1]
    return bbox_xyxy


def bbox_xyxy2cs(bbox: np.ndarray, padding: float = 1.0) -> Tuple[np.ndarray, np.ndarray]:
    dim = bbox.ndim
    if dim == 1:
        bbox = bbox[None, :]
    scale = (bbox[..., 2:] - bbox[..., :2]) * padding
    center = (bbox[..., 2:] + bbox[..., :2]) * 0.5
    if dim == 1:
        center = center[0]
        scale = scale[0]
    return center, scale


def get_udp_warp_matrix(center: np.ndarray, scale: np.ndarray, rot: float, output_size: Tuple[int, int]) -> np.ndarray:
    assert len(center) == 2
    assert len(scale) == 2
    assert len(output_size) == 2
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


class TopdownAffine:
    """Geometry of ``TopdownAffine(input_size, use_udp=True, input_padding)``; the pixel work (cv2.warpAffine)
    is done for all persons at once on the GPU by :func:`inference_topdown`."""

    def __init__(self, input_size: Tuple[int, int], input_padding: float = 1.25, use_udp: bool = False) -> None:
        assert len(input_size) == 2 and all(isinstance(i, int) for i in input_size), f"Invalid input_size {input_size}"
        if not use_udp:
            raise NotImplementedError("the ProbPose configs use use_udp=True; the non-UDP matrix is not built")
        self.input_size = input_size
        self.use_udp = use_udp
        self.input_padding = input_padding

    @staticmethod
    def _fix_aspect_ratio(bbox_scale: np.ndarray, aspect_ratio: float):
        w, h = np.hsplit(bbox_scale, [1])
        return np.where(w > h * aspect_ratio, np.hstack([w, w / aspect_ratio]), np.hstack([h * aspect_ratio, h]))

    def geometry(self, bbox: np.ndarray):
        """bbox (1, 4) xyxy -> (center (2,), scale (2,), warp_mat (2, 3) float32), topdown_transforms.py:93-118."""
        w, h = self.input_size
        _c, _s = bbox_xyxy2cs(bbox, padding=self.input_padding)
        bbox_center, bbox_scale = _c.reshape(1, 2), _s.reshape(1, 2)
        bbox_scale = self._fix_aspect_ratio(bbox_scale, aspect_ratio=w / h)
        center, scale = bbox_center[0], bbox_scale[0]
        return center, scale, get_udp_warp_matrix(center, scale, 0.0, output_size=(w, h))


def inference_topdown(model, img: Union[np.ndarray, torch.Tensor], bboxes: Optional[Union[List, np.ndarray]] = None,
                      bbox_format: str = "xyxy", flip_indices: Optional[Sequence[int]] = None) -> List[PoseDataSample]:
    """Inference image with a top-down pose estimator (apis/inference.py:133-200).

    ``img``: the loaded frame, uint8 BGR (H, W, 3) (numpy, or a CUDA tensor already on the model's device).
    ``bboxes``: (N, 4); ``None`` / empty -> the whole image.  Returns one ``PoseDataSample`` per box with
    image-space ``pred_instances`` exactly as the reference (topdown.py:128-194)."""
    from .. import ops
    from . import COCO_FLIP_INDICES

    if isinstance(img, str):  # LoadImage (mmcv.imfrombytes, cv2 backend, color): BGR uint8 (apis/inference.py:176-183)
        import cv2
        path, img = img, cv2.imread(img, cv2.IMREAD_COLOR)
        if img is None:
            raise FileNotFoundError(f"cannot read image {path!r}")
    h, w = img.shape[:2]
    if bboxes is None or len(bboxes) == 0:
        bboxes = np.array([[0, 0, w, h]], dtype=np.float32)
    else:
        if isinstance(bboxes, list):
            bboxes = np.array(bboxes)
        assert bbox_format in {"xyxy", "xywh"}, f'Invalid bbox_format "{bbox_format}".'
        if bbox_format == "xywh":
            bboxes = bbox_xywh2xyxy(bboxes)
    codec = model.head.decoder
    in_w, in_h = int(codec.input_size[0]), int(codec.input_size[1])
    affine = TopdownAffine(input_size=(in_w, in_h), use_udp=True, input_padding=getattr(model, "input_padding", 1.25))
    fi = list(flip_indices if flip_indices is not None else COCO_FLIP_INDICES)
    samples, mats = [], []
    for bbox in bboxes:
        bbox = np.asarray(bbox)[None, :4]  # shape (1, 4), inference.py:185
        center, scale, m = affine.geometry(bbox)
        # ori_shape / img_shape: what LoadImage + PackPoseInputs record (merge_data_samples reads ori_shape)
        ds = PoseDataSample(metainfo=dict(input_size=(in_w, in_h), input_center=center, input_scale=scale, flip_indices=fi,
                                          ori_shape=(h, w), img_shape=(h, w)))
        ds.gt_instances = InstanceData(bboxes=bbox, bbox_scores=np.ones(1, dtype=np.float32))
        samples.append(ds)
        mats.append(m)
    dev = model._device()
    if isinstance(img, np.ndarray):
        if img.dtype != np.uint8 or img.ndim != 3 or img.shape[2] != 3:
            raise ValueError(f"img must be uint8 (H, W, 3) BGR, got {img.dtype} {img.shape}")
        frame = torch.from_numpy(np.ascontiguousarray(img)).to(dev, non_blocking=True)
    else:
        frame = img.to(dev).contiguous()
    with torch.no_grad():
        crops = ops.crop_warp(frame, torch.from_numpy(np.stack(mats)).to(dev), out_hw=(in_h, in_w))
        return model.test_step(dict(inputs=crops, data_samples=samples))
