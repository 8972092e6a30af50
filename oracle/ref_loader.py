"""Load GENUINE reference modules from /root/reference under import stubs (TEST INFRASTRUCTURE).

``import mmpose`` fails in the build container because ``mmpose/__init__.py`` imports mmcv / mmengine
(not installed, no network).  The files on the hot path themselves need very little of those packages,
so this module installs *stubs* for the missing third-party names in ``sys.modules`` and then lets
Python's normal import machinery execute the reference's own, unmodified source files:

    mmpose/models/heads/hybrid_heads/probmap_head.py   (ProbMapHead: builders, forward, predict)
    mmpose/models/heads/heatmap_heads/heatmap_head.py  (HeatmapHead)
    mmpose/models/heads/base_head.py                   (BaseHead.decode)
    mmpose/models/utils/tta.py                         (flip_heatmaps)
    mmpose/models/pose_estimators/{base,topdown}.py    (TopdownPoseEstimator.predict / add_pred_to_datasample)
    mmpose/codecs/{base,probmap,udp_heatmap}.py + codecs/utils/*.py (ProbMap / UDPHeatmap decode)
    mmpose/utils/tensor_utils.py                       (to_numpy)

What is stubbed (and therefore NOT pinned by anything generated through this loader):

* ``mmcv.cnn.build_conv_layer`` / ``build_upsample_layer``: mmcv==2.1.0's registries map ``type="Conv2d"``
  to ``torch.nn.Conv2d`` and ``type="deconv"`` to ``torch.nn.ConvTranspose2d`` with the remaining keys as
  keyword arguments - restated here in three lines.
* ``sparsemax.Sparsemax`` (PyPI ``sparsemax``, unpinned in requirements/build.txt, not vendored): the sort /
  cumsum algorithm of Martins & Astudillo 2016, Alg. 1 (``oracle.model_oracle.sparsemax``).  PARITY UNPINNED.
* ``mmengine`` containers (``InstanceData``, ``PixelData``, ``BaseModule``, ``BaseModel``), the registries
  (``MODELS.build`` for the five loss configs returns a placeholder) and ``mmpose.structures.PoseDataSample``:
  attribute bags with ``set_field`` / ``__getitem__`` / ``metainfo``.
* The ViT backbone (mmpretrain==1.2.0) is not in the reference tree at all: the caller passes a backbone module.

Only runs where ``/root/reference`` exists (the build container); the GPU box uses the committed fixtures.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

import numpy as np
import torch
from torch import nn

REF_ROOT = "/root/reference"


class _Bag:
    """mmengine BaseDataElement stand-in: attribute bag with metainfo, set_field, item access."""

    def __init__(self, *, metainfo=None, **kw):
        object.__setattr__(self, "_meta", dict(metainfo or {}))
        object.__setattr__(self, "_data", {})
        for k, v in kw.items():
            self._data[k] = v

    @property
    def metainfo(self):
        return self._meta

    def set_metainfo(self, m):
        self._meta.update(m)

    def set_field(self, value, name, dtype=None, field_type="data"):
        (self._data if field_type == "data" else self._meta)[name] = value

    def __getattr__(self, name):
        d, m = object.__getattribute__(self, "_data"), object.__getattribute__(self, "_meta")
        if name in d:
            return d[name]
        if name in m:
            return m[name]
        raise AttributeError(name)

    def __setattr__(self, name, value):
        self._data[name] = value

    def __contains__(self, name):
        return name in self._data or name in self._meta

    def __getitem__(self, name):
        return self._data[name]

    def get(self, name, default=None):
        return self._data.get(name, self._meta.get(name, default))

    def keys(self):
        return list(self._data.keys())

    def all_keys(self):
        return list(self._data.keys()) + list(self._meta.keys())


class _Registry:
    def __init__(self, name):
        self.name, self._m = name, {}

    def register_module(self, name=None, force=False, module=None):
        def deco(cls):
            self._m[name or cls.__name__] = cls
            return cls

        return deco if module is None else deco(module)

    def get(self, key):
        return self._m.get(key)

    def build(self, cfg, **kw):
        if isinstance(cfg, nn.Module):
            return cfg
        cfg = dict(cfg)
        t = cfg.pop("type")
        if isinstance(t, str):
            if t not in self._m:
                return _Placeholder(t, cfg)  # loss modules etc.: accepted, never called at inference
            t = self._m[t]
        return t(**cfg)


class _Placeholder(nn.Module):
    def __init__(self, type_name, cfg):
        super().__init__()
        self.type_name, self.cfg = type_name, cfg


class _Sparsemax(nn.Module):
    """Stand-in for PyPI ``sparsemax.Sparsemax`` (see the module docstring: UNPINNED)."""

    def __init__(self, dim=-1):
        super().__init__()
        self.dim = dim

    def forward(self, x):
        from .model_oracle import sparsemax

        return sparsemax(x.transpose(self.dim, -1)).transpose(self.dim, -1)


def _build_conv_layer(cfg, *args, **kwargs):
    cfg = dict(cfg or dict(type="Conv2d"))
    t = cfg.pop("type")
    assert t in ("Conv2d", "Conv"), t
    return nn.Conv2d(*args, **kwargs, **cfg)


def _build_upsample_layer(cfg, *args, **kwargs):
    cfg = dict(cfg)
    t = cfg.pop("type")
    assert t == "deconv", t
    return nn.ConvTranspose2d(*args, **kwargs, **cfg)


class _BaseModule(nn.Module):
    def __init__(self, init_cfg=None):
        super().__init__()
        self.init_cfg = init_cfg

    def init_weights(self):
        pass


class _BaseModel(_BaseModule):
    def __init__(self, data_preprocessor=None, init_cfg=None):
        super().__init__(init_cfg)
        self.data_preprocessor = data_preprocessor


def _is_seq_of(seq, expected_type, seq_type=None):
    if not isinstance(seq, seq_type or (list, tuple)):
        return False
    return all(isinstance(x, expected_type) for x in seq)


def _is_method_overridden(method, base_class, derived_class):
    if not isinstance(derived_class, type):
        derived_class = derived_class.__class__
    return getattr(derived_class, method) is not getattr(base_class, method)


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def _pkg(name, path):
    """A package whose __init__.py is NOT executed but whose submodules import normally."""
    m = types.ModuleType(name)
    m.__path__ = [path]
    m.__package__ = name
    sys.modules[name] = m
    return m


_loaded = None


def load():
    """Install the stubs and import the genuine hot-path modules.  Returns a namespace with
    ``ProbMapHead, HeatmapHead, ProbMap, UDPHeatmap, flip_heatmaps, TopdownPoseEstimator, InstanceData,
    PixelData, PoseDataSample, MODELS, KEYPOINT_CODECS``."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not os.path.isdir(os.path.join(REF_ROOT, "mmpose")):
        raise FileNotFoundError(f"{REF_ROOT}/mmpose not found: the genuine reference only exists in the build container")

    models_reg, codecs_reg = _Registry("model"), _Registry("keypoint codec")

    class InstanceData(_Bag):
        pass

    class PixelData(_Bag):
        pass

    class PoseDataSample(_Bag):
        pass

    # ---- third-party stubs
    _mod("mmcv")
    _mod("mmcv.cnn", build_conv_layer=_build_conv_layer, build_upsample_layer=_build_upsample_layer)
    _mod("mmengine")
    _mod("mmengine.structures", InstanceData=InstanceData, PixelData=PixelData, BaseDataElement=_Bag)
    _mod("mmengine.model", BaseModule=_BaseModule, BaseModel=_BaseModel)
    _mod("mmengine.utils", is_seq_of=_is_seq_of, is_method_overridden=_is_method_overridden)
    _mod("mmengine.config", ConfigDict=dict)
    _mod("mmengine.dist", get_world_size=lambda: 1)
    _mod("mmengine.logging", print_log=lambda *a, **k: None)
    _mod("sparsemax", Sparsemax=_Sparsemax)

    # ---- mmpose package skeleton: __init__ files skipped, leaf files genuine
    r = os.path.join(REF_ROOT, "mmpose")
    _pkg("mmpose", r)
    _mod("mmpose.registry", MODELS=models_reg, KEYPOINT_CODECS=codecs_reg)
    _mod("mmpose.structures", PoseDataSample=PoseDataSample)
    _mod("mmpose.structures.keypoint", fix_bbox_aspect_ratio=None)
    _mod("mmpose.evaluation")
    _mod("mmpose.evaluation.functional", pose_pck_accuracy=None)
    _pkg("mmpose.utils", os.path.join(r, "utils"))  # tensor_utils.py, typing.py: genuine
    _pkg("mmpose.models", os.path.join(r, "models"))
    _mod("mmpose.models.utils", check_and_update_config=lambda neck, head: (neck, head)).__path__ = [
        os.path.join(r, "models", "utils")]  # tta.py: genuine
    _pkg("mmpose.models.heads", os.path.join(r, "models", "heads"))
    _pkg("mmpose.models.heads.hybrid_heads", os.path.join(r, "models", "heads", "hybrid_heads"))
    _pkg("mmpose.models.heads.heatmap_heads", os.path.join(r, "models", "heads", "heatmap_heads"))
    _pkg("mmpose.models.pose_estimators", os.path.join(r, "models", "pose_estimators"))
    _pkg("mmpose.datasets", os.path.join(r, "datasets"))
    _pkg("mmpose.datasets.datasets", os.path.join(r, "datasets", "datasets"))
    _mod("mmpose.datasets.datasets.utils", parse_pose_metainfo=lambda m: m)
    _pkg("mmpose.codecs", os.path.join(r, "codecs"))  # codecs/utils/__init__.py is self-contained: genuine

    ns = types.SimpleNamespace(InstanceData=InstanceData, PixelData=PixelData, PoseDataSample=PoseDataSample,
                               MODELS=models_reg, KEYPOINT_CODECS=codecs_reg)
    ns.probmap_codec = importlib.import_module("mmpose.codecs.probmap")
    ns.udp_codec = importlib.import_module("mmpose.codecs.udp_heatmap")
    ns.tta = importlib.import_module("mmpose.models.utils.tta")
    ns.probmap_head = importlib.import_module("mmpose.models.heads.hybrid_heads.probmap_head")
    ns.heatmap_head = importlib.import_module("mmpose.models.heads.heatmap_heads.heatmap_head")
    ns.topdown = importlib.import_module("mmpose.models.pose_estimators.topdown")
    for m in (ns.probmap_codec, ns.udp_codec, ns.tta, ns.probmap_head, ns.heatmap_head, ns.topdown):
        assert m.__file__.startswith(REF_ROOT), m.__file__
    ns.ProbMap, ns.UDPHeatmap = ns.probmap_codec.ProbMap, ns.udp_codec.UDPHeatmap
    ns.ProbMapHead, ns.HeatmapHead = ns.probmap_head.ProbMapHead, ns.heatmap_head.HeatmapHead
    ns.flip_heatmaps = ns.tta.flip_heatmaps
    ns.TopdownPoseEstimator = ns.topdown.TopdownPoseEstimator
    # the reference builds an ArgMaxProbMap "fast decoder" in the head's constructor (probmap_head.py:153-155,
    # training-time only); the class lives in codecs/argmax_probmap.py when present
    try:
        ns.argmax_codec = importlib.import_module("mmpose.codecs.argmax_probmap")
    except Exception:  # noqa: BLE001 - registry returns a placeholder for unknown types
        ns.argmax_codec = None
    _loaded = ns
    return ns


# The shipped configuration of the head (configs/body_2d_keypoint/topdown_probmap/coco/
# td-pm_ProbPose-small_8xb64-210e_coco-256x192.py:48,68-84), restated as plain dicts.
PROBMAP_CODEC_CFG = dict(type="ProbMap", input_size=(192, 256), heatmap_size=(48, 64), sigma=-1)


def probmap_head_cfg(in_channels=384, out_channels=17, deconv_out_channels=(256, 256), decoder=None):
    return dict(
        in_channels=in_channels, out_channels=out_channels, deconv_out_channels=tuple(deconv_out_channels),
        deconv_kernel_sizes=tuple(4 for _ in deconv_out_channels),
        keypoint_loss=dict(type="OKSHeatmapLoss"), probability_loss=dict(type="BCELoss"),
        visibility_loss=dict(type="BCELoss"), oks_loss=dict(type="MSELoss"), error_loss=dict(type="L1LogLoss"),
        normalize=1.0, decoder=dict(decoder or PROBMAP_CODEC_CFG))


def as_numpy(x):
    return x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else np.asarray(x)
