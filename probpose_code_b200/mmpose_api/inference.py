"""``inference_topdown`` for the B200 path: the caller contract of mmpose/apis/inference.py:133-200 with
the per-person CPU pipeline (GetBBoxCenterScale -> TopdownAffine -> PackPoseInputs -> pseudo_collate ->
H2D) replaced by ONE frame upload and one GPU warp launch (``pp_crop_warp``), then ``pp_engine_infer``.

Host-side geometry is computed for all boxes of a frame at once, with the reference's float32 operations per element,
so that centres, scales and matrices are the same float32 values as those of:
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
    """(N, 4+) boxes ``x, y, w, h`` -> ``x0, y0, x1, y1`` (what structures/bbox/transforms.py:26-41 returns)."""
    out = np.array(bbox_xywh, copy=True)
    out[:, 2:4] += out[:, 0:2]
    return out


def bbox_xyxy2cs(bbox: np.ndarray, padding: float = 1.0) -> Tuple[np.ndarray, np.ndarray]:
    """Boxes (4,) or (N, 4) -> centre and padded extent, the values of structures/bbox/transforms.py:44-72
    (``(x1 + x0) * 0.5`` and ``(x1 - x0) * padding`` in the boxes' dtype)."""
    lo, hi = bbox[..., 0:2], bbox[..., 2:4]
    return (hi + lo) * 0.5, (hi - lo) * padding


def fix_aspect(scales: np.ndarray, aspect: float) -> np.ndarray:
    """Grow (N, 2) extents to width / height = ``aspect`` (``TopdownAffine._fix_aspect_ratio``,
    topdown_transforms.py:70-91): the wider-than-``aspect`` boxes keep their width, the others their height."""
    w, h = scales[:, 0], scales[:, 1]
    wide = w > h * aspect
    return np.stack([np.where(wide, w, h * aspect), np.where(wide, w / aspect, h)], axis=1)


def udp_warp_matrices(centers: np.ndarray, scales: np.ndarray, rot: float, output_size: Tuple[int, int]) -> np.ndarray:
    """(N, 2) centres and extents -> the (N, 2, 3) float32 matrices ``TopdownAffine(use_udp=True)`` hands to
    ``cv2.warpAffine``, for all persons of a frame in one pass.  Every entry is formed by the float32 operations, in the
    order, of ``get_udp_warp_matrix`` (structures/bbox/transforms.py:315-359) - the crops are compared with the
    reference's bit for bit, and a last-bit difference in a matrix moves pixels (tests/golden/crop_kat.npz)."""
    ang = float(np.deg2rad(rot))
    c, s = math.cos(ang), math.sin(ang)
    gx, gy = (output_size[0] - 1) / scales[:, 0], (output_size[1] - 1) / scales[:, 1]  # python int / float32 column: float32
    ex, ey = centers[:, 0] * 2, centers[:, 1] * 2                                    # "input_size = center * 2"
    m = np.zeros((centers.shape[0], 2, 3), dtype=np.float32)
    m[:, 0, 0], m[:, 0, 1] = c * gx, -s * gx
    m[:, 1, 0], m[:, 1, 1] = s * gy, c * gy
    m[:, 0, 2] = gx * (-0.5 * ex * c + 0.5 * ey * s + 0.5 * scales[:, 0])
    m[:, 1, 2] = gy * (-0.5 * ex * s - 0.5 * ey * c + 0.5 * scales[:, 1])
    return m


def get_udp_warp_matrix(center: np.ndarray, scale: np.ndarray, rot: float, output_size: Tuple[int, int]) -> np.ndarray:
    """One person's matrix (the reference's signature); see :func:`udp_warp_matrices`."""
    assert len(center) == 2 and len(scale) == 2 and len(output_size) == 2
    return udp_warp_matrices(np.asarray(center)[None], np.asarray(scale)[None], rot, output_size)[0]


class TopdownAffine:
    """Geometry of ``TopdownAffine(input_size, use_udp=True, input_padding)`` for all boxes of a frame at once; the pixel
    work (cv2.warpAffine) is done for all persons by one GPU launch in :func:`inference_topdown`."""

    def __init__(self, input_size: Tuple[int, int], input_padding: float = 1.25, use_udp: bool = False) -> None:
        assert len(input_size) == 2 and all(isinstance(i, int) for i in input_size), f"Invalid input_size {input_size}"
        if not use_udp:
            raise NotImplementedError("the ProbPose configs use use_udp=True; the non-UDP matrix is not built")
        self.input_size = input_size
        self.use_udp = use_udp
        self.input_padding = input_padding

    def batch_geometry(self, bboxes: np.ndarray):
        """(N, 4) xyxy -> centres (N, 2), aspect-fixed extents (N, 2), matrices (N, 2, 3) float32
        (GetBBoxCenterScale + topdown_transforms.py:93-118 for every box)."""
        w, h = self.input_size
        centers, scales = bbox_xyxy2cs(bboxes, padding=self.input_padding)
        scales = fix_aspect(scales, w / h)
        return centers, scales, udp_warp_matrices(centers, scales, 0.0, (w, h))

    def geometry(self, bbox: np.ndarray):
        """One box, (1, 4) xyxy -> (center (2,), scale (2,), warp_mat (2, 3) float32)."""
        c, s, m = self.batch_geometry(np.asarray(bbox).reshape(1, -1)[:, :4])
        return c[0], s[0], m[0]


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
    boxes = np.asarray(bboxes)[:, :4]
    centers, scales, mats = affine.batch_geometry(boxes)  # every person of the frame in one pass
    samples = []
    for i in range(boxes.shape[0]):
        # ori_shape / img_shape: what LoadImage + PackPoseInputs record (merge_data_samples reads ori_shape)
        ds = PoseDataSample(metainfo=dict(input_size=(in_w, in_h), input_center=centers[i], input_scale=scales[i],
                                          flip_indices=fi, ori_shape=(h, w), img_shape=(h, w)))
        ds.gt_instances = InstanceData(bboxes=boxes[i:i + 1], bbox_scores=np.ones(1, dtype=np.float32))
        samples.append(ds)
    dev = model._device()
    if isinstance(img, np.ndarray):
        if img.dtype != np.uint8 or img.ndim != 3 or img.shape[2] != 3:
            raise ValueError(f"img must be uint8 (H, W, 3) BGR, got {img.dtype} {img.shape}")
        frame = torch.from_numpy(np.ascontiguousarray(img)).to(dev, non_blocking=True)
    else:
        frame = img.to(dev).contiguous()
    with torch.no_grad():
        crops = ops.crop_warp(frame, torch.from_numpy(mats).to(dev), out_hw=(in_h, in_w))
        return model.test_step(dict(inputs=crops, data_samples=samples))
