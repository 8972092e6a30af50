"""``TopdownPoseEstimator`` + ``PoseDataPreprocessor`` for the predict path
(mmpose/models/pose_estimators/{base,topdown}.py, models/data_preprocessors/data_preprocessor.py).

When backbone and head are this package's ``VisionTransformer`` and ``ProbMapHead`` the whole
``predict`` is ONE engine call (``pp_engine_infer``): uint8 crops in, (B, K, 7) records out, with
preprocessing, both flip passes, sparsemax, TTA merge and decode on the device.  Any other
combination falls back to the reference's generic module-by-module flow (still CUDA modules -
there is no CPU path)."""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np
import torch
from torch import nn

from ._engine_cache import EngineCache
from .backbone import VisionTransformer
from .head import HeatmapHead, ProbMapHead
from .registry import MODELS, register
from .structures import PixelData


@register(MODELS, ["PoseDataPreprocessor"])
class PoseDataPreprocessor(nn.Module):
    """BGR->RGB, float, ``(x - mean) / std``, stack (data_preprocessor.py:79-104 + mmengine
    ``ImgDataPreprocessor``).  Used on the generic path; the fused path hands the uint8 crops to
    the engine, which applies the same arithmetic inside the patch-extraction kernel."""

    def __init__(self, mean: Sequence[float] = None, std: Sequence[float] = None, pad_size_divisor: int = 1,
                 pad_value=0, bgr_to_rgb: bool = False, rgb_to_bgr: bool = False, non_blocking: bool = False,
                 batch_augments=None):
        super().__init__()
        assert not (bgr_to_rgb and rgb_to_bgr), "`bgr2rgb` and `rgb2bgr` cannot be set to True at the same time"
        assert (mean is None) == (std is None), "mean and std should be both None or tuple"
        self.channel_conversion = bgr_to_rgb or rgb_to_bgr
        self.mean_std = None if mean is None else (tuple(mean), tuple(std))
        if mean is not None:
            self.register_buffer("mean", torch.tensor(mean).view(-1, 1, 1), False)
            self.register_buffer("std", torch.tensor(std).view(-1, 1, 1), False)

    @staticmethod
    def stack(inputs) -> torch.Tensor:
        return torch.stack(list(inputs)) if not isinstance(inputs, torch.Tensor) else inputs

    def forward(self, data: dict, training: bool = False) -> dict:
        x = self.stack(data["inputs"]).to(self.mean.device if self.mean_std else None)
        if self.channel_conversion:
            x = x[:, [2, 1, 0]]
        x = x.float()
        if self.mean_std:
            x = (x - self.mean) / self.std
        return {"inputs": x, "data_samples": data.get("data_samples")}


@register(MODELS, ["TopdownPoseEstimator"])
class TopdownPoseEstimator(nn.Module):
    _version = 2

    def __init__(self, backbone: dict, neck: Optional[dict] = None, head: Optional[dict] = None,
                 train_cfg: Optional[dict] = None, test_cfg: Optional[dict] = None,
                 data_preprocessor: Optional[dict] = None, init_cfg=None, metainfo: Optional[dict] = None,
                 freeze_backbone: bool = False, precision: str = None):
        super().__init__()
        self.metainfo = metainfo
        self.train_cfg = train_cfg if train_cfg else {}
        self.test_cfg = test_cfg if test_cfg else {}
        if precision is not None:
            backbone = dict(backbone, precision=precision)
            head = dict(head, precision=precision) if head is not None else None
        self.data_preprocessor = MODELS.build(data_preprocessor or dict(type="PoseDataPreprocessor"))
        self.backbone = MODELS.build(backbone)
        if neck is not None:
            self.neck = MODELS.build(neck)
        if head is not None:
            self.head = MODELS.build(head)
            self.head.test_cfg = self.test_cfg.copy()
        self._fused = None
        self._stage = None  # (pinned host crops, device crops, pinned host records) of the fused path, grown on demand
        self.eval()

    with_neck = property(lambda self: hasattr(self, "neck") and self.neck is not None)
    with_head = property(lambda self: hasattr(self, "head") and self.head is not None)

    # ---- fused engine -----------------------------------------------------------------
    def _fusable(self) -> bool:
        return (isinstance(self.backbone, VisionTransformer) and self.with_head and isinstance(self.head, (ProbMapHead, HeatmapHead))
                and not self.with_neck and self.head.decoder is not None and self.head.fused_decoder()
                and self.head.fused_test_cfg(self.test_cfg)
                and self.backbone._cache.precision == self.head._cache.precision)

    def _fused_engine(self, images: int, device):
        if self._fused is None:
            kw = dict(self.backbone._cache.kwargs)
            hk = self.head._cache.kwargs
            kw.update({k: hk[k] for k in ("num_keypoints", "deconv_channels", "temperature", "normalize", "head_kind",
                                          "blur_kernel_size") if k in hk})
            pre = self.data_preprocessor
            if getattr(pre, "mean_std", None):
                kw.update(mean=pre.mean_std[0], std=pre.mean_std[1])
            self._fused = EngineCache(kw, self.backbone._cache.precision)
        tensors = dict(self.backbone.engine_tensors())
        tensors.update(self.head.engine_tensors())
        return self._fused.get(tensors, images, device)

    def _device(self):
        return self.backbone.pos_embed.device

    def _staging(self, batch: int, shape, rec_shape, device):
        """Pinned host staging for the step's crops and records plus the device-side crop buffer (one allocation per
        capacity, reused by every call): per-person host tensors are gathered straight into pinned memory (no
        ``torch.stack`` into pageable memory followed by a staged pageable copy), go to the device with ONE asynchronous
        copy, and the records come back through pinned memory as well."""
        st = self._stage
        if st is None or st[0].shape[0] < batch or tuple(st[0].shape[1:]) != tuple(shape) or st[1].device != device \
                or tuple(st[2].shape[1:]) != tuple(rec_shape):
            cap = max(16, 1 << (max(batch, 1) - 1).bit_length())
            st = (torch.empty((cap, *shape), dtype=torch.uint8).pin_memory(),
                  torch.empty((cap, *shape), dtype=torch.uint8, device=device),
                  torch.empty((cap, *rec_shape), dtype=torch.float32).pin_memory())
            self._stage = st
        return st

    def _upload(self, inputs, device) -> Optional[torch.Tensor]:
        """uint8 crops on the host (a list of per-person (3, H, W) tensors as mmengine's ``pseudo_collate`` hands them
        over, or one stacked tensor) -> device tensor through the pinned staging buffer.  Returns None when the inputs
        are not host uint8 crops (the caller takes the generic route)."""
        if isinstance(inputs, torch.Tensor):
            if inputs.is_cuda or inputs.dtype != torch.uint8 or inputs.dim() != 4:
                return None
            n, shape = inputs.shape[0], inputs.shape[1:]
            if inputs.is_pinned():  # already page-locked: copy straight from the caller's buffer
                _, dev_buf, _ = self._staging(n, shape, self._rec_shape(), device)
                dev_buf[:n].copy_(inputs, non_blocking=True)
                return dev_buf[:n]
            host, dev_buf, _ = self._staging(n, shape, self._rec_shape(), device)
            host[:n].copy_(inputs)
        else:
            inputs = list(inputs)
            if not inputs or any((not isinstance(t, torch.Tensor)) or t.is_cuda or t.dtype != torch.uint8 or t.dim() != 3
                                 for t in inputs):
                return None
            n, shape = len(inputs), inputs[0].shape
            host, dev_buf, _ = self._staging(n, shape, self._rec_shape(), device)
            step = max(1, (n + 3) // 4)  # quarters: the upload of one runs under the host gather of the next
            for lo in range(0, n, step):
                hi = min(n, lo + step)
                torch.stack(inputs[lo:hi], out=host[lo:hi])  # one multi-threaded gather per quarter, straight into pinned memory
                dev_buf[lo:hi].copy_(host[lo:hi], non_blocking=True)
            return dev_buf[:n]
        dev_buf[:n].copy_(host[:n], non_blocking=True)
        return dev_buf[:n]

    def _rec_shape(self):
        return (self.head.out_channels, 7 if isinstance(self.head, ProbMapHead) else 3)

    # ---- mmengine BaseModel surface -----------------------------------------------------
    @torch.no_grad()
    def test_step(self, data: dict) -> list:
        """``BaseModel.test_step``: ``data = dict(inputs=[uint8 BGR (3,H,W)...], data_samples=[...])``."""
        pre = self.data_preprocessor
        fused = self._fusable() and getattr(pre, "channel_conversion", False) and getattr(pre, "mean_std", None)
        if fused:
            dev = self._upload(data["inputs"], self._device())  # host uint8 crops: pinned staging, one async copy
            if dev is not None:
                return self._predict_fused(dev, data["data_samples"], pinned_records=True)
        inputs = PoseDataPreprocessor.stack(data["inputs"])
        if fused and inputs.dtype == torch.uint8:
            return self._predict_fused(inputs.to(self._device(), non_blocking=True).contiguous(), data["data_samples"])
        data = pre(data, False)
        return self.forward(data["inputs"], data["data_samples"], mode="predict")

    def forward(self, inputs: torch.Tensor, data_samples=None, mode: str = "tensor"):
        """pose_estimators/base.py:123-168."""
        if isinstance(inputs, list):
            inputs = torch.stack(inputs)
        if mode == "loss":
            return self.loss(inputs, data_samples)
        elif mode == "predict":
            if self.metainfo is not None:
                for data_sample in data_samples:
                    data_sample.set_metainfo(self.metainfo)
            return self.predict(inputs, data_samples)
        elif mode == "tensor":
            return self._forward(inputs)
        raise RuntimeError(f'Invalid mode "{mode}". ' "Only supports loss, predict and tensor mode.")

    def loss(self, inputs, data_samples):
        raise NotImplementedError("training is out of scope of the B200 inference path")

    def extract_feat(self, inputs: torch.Tensor):
        x = self.backbone(inputs)
        if self.with_neck:
            x = self.neck(x)
        return x

    def _forward(self, inputs: torch.Tensor, data_samples=None):
        x = self.extract_feat(inputs)
        if self.with_head:
            x = self.head.forward(x)
        return x

    @torch.no_grad()
    def predict(self, inputs: torch.Tensor, data_samples: list) -> list:
        """topdown.py:86-126."""
        assert self.with_head, "The model must have head to perform prediction."
        if self._fusable():
            return self._predict_fused(inputs.float().contiguous(), data_samples)
        if self.test_cfg.get("flip_test", False):
            feats = [self.extract_feat(inputs), self.extract_feat(inputs.flip(-1))]
        else:
            feats = self.extract_feat(inputs)
        preds = self.head.predict(feats, data_samples, test_cfg=self.test_cfg)
        if isinstance(preds, tuple):
            batch_pred_instances, batch_pred_fields = preds
        else:
            batch_pred_instances, batch_pred_fields = preds, None
        return self.add_pred_to_datasample(batch_pred_instances, batch_pred_fields, data_samples)

    def _predict_fused(self, inputs: torch.Tensor, data_samples: list, pinned_records: bool = False) -> list:
        """One ``pp_engine_infer`` call for the batch (uint8 BGR or normalised fp32 crops).  ``pinned_records``: read
        the records back through the pinned staging buffer (asynchronous copy + one stream synchronisation)."""
        cfg = self.test_cfg
        flip = bool(cfg.get("flip_test", False))
        want_hm = bool(cfg.get("output_heatmaps", False))
        flip_indices = data_samples[0].metainfo["flip_indices"] if flip else None
        eng = self._fused_engine(inputs.shape[0] * (2 if flip else 1), inputs.device)
        out = eng.infer(inputs, flip_test=flip, flip_indices=flip_indices, return_heatmaps=want_hm)
        records, heatmaps = out if want_hm else (out, None)
        # operand range guard: after the first call of an engine (weights + a real batch have been seen) and every
        # 256 calls afterwards - the check synchronises the device, so not on every call
        self._calls = getattr(self, "_calls", 0) + 1
        if getattr(self, "_guard_engine", None) is not eng or self._calls % 256 == 0:
            eng.raise_on_overflow()
            self._guard_engine = eng
        host = None
        if pinned_records and self._stage is not None and self._stage[2].shape[0] >= records.shape[0]:
            host = self._stage[2][:records.shape[0]]
            host.copy_(records, non_blocking=True)  # queued behind the decode kernel; waited for below
        fields = [PixelData(heatmaps=hm) for hm in heatmaps] if want_hm else None
        if cfg.get("output_keypoint_indices", None) is not None:
            if host is not None:
                torch.cuda.current_stream(records.device).synchronize()
                records = host
            return self.add_pred_to_datasample(self.head.pack_records(records), fields, data_samples)
        # The device is busy for milliseconds: build everything that does not need the results now - the per-person
        # containers (views of fresh batch arrays), the bbox fields, the geometry of topdown.py:165-167 - and only
        # then wait for the records and fill the batch arrays in place.
        if records.shape[0] == 0:
            return data_samples
        arrays, preds = self.head.alloc_records(records.shape[0])
        self.add_pred_to_datasample(preds, fields, data_samples, mapped=True)
        geo = np.concatenate([a for d in data_samples for m in (d.metainfo,)
                              for a in (m["input_size"], m["input_scale"], m["input_center"])]).reshape(-1, 3, 2)
        # float64 copies (exact; the keypoints are float64) repeated per keypoint: rows of 2 K numbers keep numpy out
        # of its length-2 inner loops and casting buffers after the device has finished
        kk = self.head.out_channels
        half = np.tile((0.5 * geo[:, 1]).astype(np.float64), kk)
        size, scale, center = (np.tile(geo[:, i].astype(np.float64), kk) for i in range(3))

        def to_image(k):  # topdown.py:165-167 for the whole batch at once (same float arithmetic per element)
            return (k.reshape(k.shape[0], -1) / size * scale + center - half).reshape(k.shape)

        if host is not None:
            torch.cuda.current_stream(records.device).synchronize()
            rec = host.numpy()
        else:
            rec = records.detach().cpu().numpy()
        self.head.fill_records(arrays, rec, to_image=to_image)
        return data_samples

    def add_pred_to_datasample(self, batch_pred_instances: list, batch_pred_fields: Optional[list],
                               batch_data_samples: list, mapped: bool = False) -> list:
        """Attach the predictions to the data samples with the field semantics of topdown.py:128-194 (the evaluator
        and the visualiser read these names): keypoints go from input space to image space
        (``k / input_size * input_scale + input_center - 0.5 * input_scale``), ``keypoints_visible`` defaults to the
        scores, ``output_keypoint_indices`` selects keypoints in every ``keypoint*`` field and in the per-keypoint
        prediction fields, the ground-truth boxes are carried over.  ``mapped``: the keypoints are already in image
        space (the fused path maps the whole batch in one vectorised step)."""
        if len(batch_pred_instances) != len(batch_data_samples):
            raise AssertionError("one prediction per data sample")
        keep = self.test_cfg.get("output_keypoint_indices", None)
        fields = list(batch_pred_fields) if batch_pred_fields else []
        fields += [None] * (len(batch_data_samples) - len(fields))
        for sample, inst, fld in zip(batch_data_samples, batch_pred_instances, fields):
            if inst is None:
                continue
            if not mapped:
                meta = sample.metainfo
                size, scale, center = meta["input_size"], meta["input_scale"], meta["input_center"]
                inst.keypoints[..., :2] = inst.keypoints[..., :2] / size * scale + center - 0.5 * scale
            if "keypoints_visible" not in inst:
                inst.keypoints_visible = inst.keypoint_scores
            if keep is not None:
                n_kpt = inst.keypoints.shape[1]
                for name, value in inst.all_items():
                    if name.startswith("keypoint"):
                        inst.set_field(value[:, keep], name)
            gt = sample.gt_instances
            inst.bboxes, inst.bbox_scores = gt.bboxes, gt.bbox_scores
            sample.pred_instances = inst
            if fld is not None:
                if keep is not None:
                    for name, value in fld.all_items():
                        if value.shape[0] == n_kpt:
                            fld.set_field(value[keep], name)
                sample.pred_fields = fld
        return batch_data_samples
