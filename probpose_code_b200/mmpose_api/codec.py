"""``ProbMap`` and ``UDPHeatmap`` codecs, decode side, on the GPU (mirror mmpose/codecs/probmap.py,
mmpose/codecs/udp_heatmap.py and mmpose/codecs/base.py).  ``batch_decode`` is overridden, so ``BaseHead.decode`` takes its
batched branch (base_head.py:57-62) and nothing round-trips through per-person numpy."""
from __future__ import annotations

from typing import List, Tuple

import numpy as np
import torch

from .. import ops
from .registry import KEYPOINT_CODECS, register


class BaseKeypointCodec:
    """mmpose/codecs/base.py:9-77."""

    auxiliary_encode_keys = set()
    field_mapping_table = dict()
    instance_mapping_table = dict()
    label_mapping_table = dict()

    def encode(self, keypoints, keypoints_visible=None) -> dict:
        raise NotImplementedError

    def decode(self, encoded):
        raise NotImplementedError

    def batch_decode(self, batch_encoded):
        raise NotImplementedError()

    @property
    def support_batch_decoding(self) -> bool:
        return type(self).batch_decode is not BaseKeypointCodec.batch_decode


def _locs_to_input_space(locs: np.ndarray, heatmap_size, input_size) -> np.ndarray:
    """``locs / [W - 1, H - 1] * input_size`` in float64 - the values numpy's mixed float32 / python-int arithmetic
    produces - evaluated on rows of 2 K numbers (numpy's inner loop over a last axis of length 2, with casting buffers,
    costs 30 us for a batch of 64; this is host time after the device has finished)."""
    k2 = locs.shape[-1] * (locs.shape[-2] if locs.ndim >= 2 else 1)
    key = (tuple(heatmap_size), tuple(input_size), k2)
    rows = _LOC_ROWS.get(key)
    if rows is None:
        w, h = heatmap_size
        rows = _LOC_ROWS[key] = (np.tile([w - 1.0, h - 1.0], k2 // 2), np.tile(np.asarray(input_size, dtype=np.float64), k2 // 2))
    out = locs.astype(np.float64).reshape(-1, k2)
    out /= rows[0]
    out *= rows[1]
    return out.reshape(locs.shape)


_LOC_ROWS: dict = {}


@register(KEYPOINT_CODECS, ["ProbMap"])
class ProbMap(BaseKeypointCodec):
    """Same constructor as the reference (probmap.py:71-96).  Only the ``"gaussian"`` heatmap
    type decodes here (the shipped config); ``encode`` builds training targets and is out of
    scope for the inference hot path."""

    label_mapping_table = dict(keypoint_weights="keypoint_weights")
    field_mapping_table = dict(heatmaps="heatmaps")

    def __init__(self, input_size: Tuple[int, int], heatmap_size: Tuple[int, int], heatmap_type: str = "gaussian",
                 sigma: float = 2.0, radius_factor: float = 0.0546875, blur_kernel_size: int = 11,
                 increase_sigma_with_padding=False) -> None:
        super().__init__()
        self.input_size = input_size
        self.heatmap_size = heatmap_size
        self.radius_factor = radius_factor
        self.heatmap_type = heatmap_type
        self.blur_kernel_size = blur_kernel_size
        self.scale_factor = ((np.array(input_size) - 1) / (np.array(heatmap_size) - 1)).astype(np.float32)
        self.increase_sigma_with_padding = increase_sigma_with_padding
        self.sigma = sigma
        if self.heatmap_type not in {"gaussian", "combined"}:
            raise ValueError(f"{self.__class__.__name__} got invalid `heatmap_type` value"
                             f"{self.heatmap_type}. Should be one of " '{"gaussian", "combined"}')

    def encode(self, keypoints, keypoints_visible=None, id_similarity=0.0, keypoints_visibility=None) -> dict:
        raise NotImplementedError("ProbMap.encode builds training targets; probpose_code_b200 covers inference only")

    # -- records -> the reference's return types -------------------------------------------
    def keypoints_from_locs(self, locs: np.ndarray) -> np.ndarray:
        """probmap.py:218: ``keypoints / [W - 1, H - 1] * input_size`` (float64, as the
        reference's python-list arithmetic produces)."""
        return _locs_to_input_space(locs, self.heatmap_size, self.input_size)

    def _decode_device(self, heatmaps: torch.Tensor) -> np.ndarray:
        if self.heatmap_type != "gaussian":
            raise NotImplementedError('only heatmap_type="gaussian" is implemented on the GPU')
        w, h = self.heatmap_size
        assert heatmaps.dim() == 4 and tuple(heatmaps.shape[-2:]) == (h, w), (
            f"heatmaps must be (B, K, {h}, {w}), got {tuple(heatmaps.shape)}")
        rec = ops.decode(heatmaps.float().contiguous(), input_is_logits=False)
        return rec.cpu().numpy()

    def decode(self, encoded: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
        """(K, H, W) float32 heatmaps -> keypoints (1, K, 2) float64 in input pixels, scores
        (1, K) float32 (probmap.py:170-220)."""
        assert isinstance(encoded, np.ndarray) and encoded.ndim == 3, "expects heatmaps in shape (K, H, W)"
        rec = self._decode_device(torch.from_numpy(np.ascontiguousarray(encoded, np.float32)).cuda()[None])
        return self.keypoints_from_locs(rec[:, :, :2]), rec[:, :, 2]

    def batch_decode(self, batch_encoded: torch.Tensor) -> Tuple[List[np.ndarray], List[np.ndarray]]:
        """(B, K, H, W) CUDA tensor -> per-person lists, one kernel launch for the batch."""
        rec = self._decode_device(batch_encoded)
        kpts = self.keypoints_from_locs(rec[:, :, :2])
        return [k[None] for k in kpts], [s[None] for s in rec[:, :, 2]]


@register(KEYPOINT_CODECS, ["UDPHeatmap"])
class UDPHeatmap(BaseKeypointCodec):
    """Same constructor as the reference (udp_heatmap.py:70-98).  Only the ``"gaussian"`` heatmap type decodes
    here (the ViTPose td-hm configs); ``encode`` builds training targets and is out of scope."""

    label_mapping_table = dict(keypoint_weights="keypoint_weights")
    field_mapping_table = dict(heatmaps="heatmaps")

    def __init__(self, input_size: Tuple[int, int], heatmap_size: Tuple[int, int], heatmap_type: str = "gaussian",
                 sigma: float = 2.0, radius_factor: float = 0.0546875, blur_kernel_size: int = 11) -> None:
        super().__init__()
        self.input_size = input_size
        self.heatmap_size = heatmap_size
        self.sigma = sigma
        self.radius_factor = radius_factor
        self.heatmap_type = heatmap_type
        self.blur_kernel_size = blur_kernel_size
        self.scale_factor = ((np.array(input_size) - 1) / (np.array(heatmap_size) - 1)).astype(np.float32)
        if self.heatmap_type not in {"gaussian", "combined"}:
            raise ValueError(f"{self.__class__.__name__} got invalid `heatmap_type` value"
                             f"{self.heatmap_type}. Should be one of " '{"gaussian", "combined"}')

    def encode(self, keypoints, keypoints_visible=None) -> dict:
        raise NotImplementedError("UDPHeatmap.encode builds training targets; probpose_code_b200 covers inference only")

    def keypoints_from_locs(self, locs: np.ndarray) -> np.ndarray:
        """udp_heatmap.py:194-195: ``keypoints / [W - 1, H - 1] * input_size`` (float64)."""
        return _locs_to_input_space(locs, self.heatmap_size, self.input_size)

    def _decode_device(self, heatmaps: torch.Tensor) -> np.ndarray:
        if self.heatmap_type != "gaussian":
            raise NotImplementedError('only heatmap_type="gaussian" is implemented on the GPU')
        w, h = self.heatmap_size
        assert heatmaps.dim() == 4 and tuple(heatmaps.shape[-2:]) == (h, w), (
            f"heatmaps must be (B, K, {h}, {w}), got {tuple(heatmaps.shape)}")
        return ops.decode_udp(heatmaps.float().contiguous(), blur_kernel_size=self.blur_kernel_size).cpu().numpy()

    def decode(self, encoded: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
        """(K, H, W) float32 heatmaps -> keypoints (1, K, 2) float64 in input pixels, scores (1, K) float32
        (udp_heatmap.py:146-196)."""
        assert isinstance(encoded, np.ndarray) and encoded.ndim == 3, "expects heatmaps in shape (K, H, W)"
        rec = self._decode_device(torch.from_numpy(np.ascontiguousarray(encoded, np.float32)).cuda()[None])
        return self.keypoints_from_locs(rec[:, :, :2]), rec[:, :, 2]

    def batch_decode(self, batch_encoded: torch.Tensor) -> Tuple[List[np.ndarray], List[np.ndarray]]:
        rec = self._decode_device(batch_encoded)
        kpts = self.keypoints_from_locs(rec[:, :, :2])
        return [k[None] for k in kpts], [s[None] for s in rec[:, :, 2]]
