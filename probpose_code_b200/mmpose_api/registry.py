"""Registries the ProbPose config resolves its ``type=`` strings through.

Mirrors ``mmpose/registry.py:50`` (``MODELS``) and ``:92`` (``KEYPOINT_CODECS``).  When the
real mmengine + mmpose are importable the classes of this package are registered INTO those
registries (``force=True``), which is the drop-in: an unmodified
``configs/body_2d_keypoint/topdown_probmap/...py`` plus
``custom_imports = dict(imports=["probpose_code_b200.mmpose_api"])`` then builds the B200
modules.  When they are absent (this image) a small compatible ``Registry`` stands in.
"""
from __future__ import annotations

from typing import Callable, Dict, Optional


class Registry:
    """Subset of ``mmengine.registry.Registry``: ``register_module`` (decorator or call),
    ``get``, ``build(cfg)`` with ``cfg['type']`` a registered name or a class.  A scope
    prefix (``"mmpretrain.VisionTransformer"``) falls back to the bare name, like
    mmengine's cross-scope lookup does when the scope is this package."""

    def __init__(self, name: str):
        self.name = name
        self._modules: Dict[str, type] = {}

    def register_module(self, name: Optional[str] = None, force: bool = False, module: Optional[type] = None):
        def _do(cls):
            names = [name] if isinstance(name, str) else (list(name) if name else [cls.__name__])
            for n in names:
                if n in self._modules and not force and self._modules[n] is not cls:
                    raise KeyError(f"{n} is already registered in {self.name}")
                self._modules[n] = cls
            return cls

        if module is not None:
            return _do(module)
        return _do

    def get(self, key: str):
        if key in self._modules:
            return self._modules[key]
        if "." in key and key.split(".", 1)[1] in self._modules:
            return self._modules[key.split(".", 1)[1]]
        return None

    def build(self, cfg: dict, **default_args):
        if not isinstance(cfg, dict) or "type" not in cfg:
            raise TypeError(f"cfg must be a dict with a 'type' key, got {cfg!r}")
        args = dict(cfg)
        t = args.pop("type")
        cls = t if isinstance(t, type) else self.get(t)
        if cls is None:
            raise KeyError(f"{t} is not in the {self.name} registry")
        for k, v in default_args.items():
            args.setdefault(k, v)
        return cls(**args)

    def __contains__(self, key: str) -> bool:
        return self.get(key) is not None


HAVE_MMPOSE = False
try:  # the real thing, when installed next to this package
    from mmpose.registry import KEYPOINT_CODECS, MODELS  # type: ignore # noqa: F401

    HAVE_MMPOSE = True
except Exception:  # noqa: BLE001 - mmengine / mmcv / mmpose absent
    MODELS = Registry("model")
    KEYPOINT_CODECS = Registry("keypoint codec")


def register(registry, names) -> Callable:
    """Register under every name in ``names``, overriding an existing entry (drop-in)."""

    def _do(cls):
        for n in names:
            registry.register_module(name=n, force=True, module=cls)
        return cls

    return _do
