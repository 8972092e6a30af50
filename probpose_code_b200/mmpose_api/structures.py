"""Result containers.  With mmengine installed these ARE ``mmengine.structures.InstanceData`` /
``PixelData`` and ``mmpose.structures.PoseDataSample``; otherwise minimal stand-ins with the
methods the hot path touches (``set_field``, item / attribute access, ``all_items``,
``metainfo`` / ``set_metainfo``) - probmap_head.py:780-804, topdown.py:128-194."""
from __future__ import annotations

try:
    from mmengine.structures import InstanceData, PixelData  # type: ignore # noqa: F401
    from mmpose.structures import PoseDataSample  # type: ignore # noqa: F401
except Exception:  # noqa: BLE001

    class _Data:
        def __init__(self, *, metainfo=None, **fields):
            object.__setattr__(self, "_fields", {})
            object.__setattr__(self, "_meta", dict(metainfo or {}))
            for k, v in fields.items():
                self.set_field(v, k)

        def set_field(self, value, name, **_):
            self._fields[name] = value

        def set_metainfo(self, metainfo: dict):
            self._meta.update(metainfo)

        @property
        def metainfo(self) -> dict:
            return self._meta

        def __setattr__(self, name, value):
            self.set_field(value, name)

        def __getattr__(self, name):
            try:
                return object.__getattribute__(self, "_fields")[name]
            except KeyError:
                raise AttributeError(name) from None

        def __getitem__(self, name):
            return self._fields[name]

        def __contains__(self, name):
            return name in self._fields

        def keys(self):
            return list(self._fields)

        def all_items(self):
            return list(self._fields.items())

        def get(self, name, default=None):
            return self._fields.get(name, default)

    class InstanceData(_Data):
        pass

    class PixelData(_Data):
        pass

    class PoseDataSample(_Data):
        """Fields used on this path: ``gt_instances`` (bboxes, bbox_scores), ``pred_instances``,
        ``pred_fields``; metainfo ``input_center / input_scale / input_size / flip_indices``."""
