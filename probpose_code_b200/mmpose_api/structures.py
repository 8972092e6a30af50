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
        # one dict per container and no per-field call in the constructor: the fused path builds a batch of these per
        # call, on the host, after the device has finished (time the GPU cannot hide)
        __slots__ = ("_fields", "_meta")

        def __init__(self, *, metainfo=None, **fields):
            object.__setattr__(self, "_fields", fields)
            object.__setattr__(self, "_meta", dict(metainfo) if metainfo else {})

        def set_field(self, value, name, **_):
            self._fields[name] = value

        def set_metainfo(self, metainfo: dict):
            self._meta.update(metainfo)

        @property
        def metainfo(self) -> dict:
            return self._meta

        def __setattr__(self, name, value):
            self._fields[name] = value

        def __getattr__(self, name):
            try:
                return self._fields[name]
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

    def _bare(cls, fields: dict):
        obj = cls.__new__(cls)
        _Data._fields.__set__(obj, fields)
        _Data._meta.__set__(obj, {})
        return obj

    class PixelData(_Data):
        pass

    class PoseDataSample(_Data):
        """Fields used on this path: ``gt_instances`` (bboxes, bbox_scores), ``pred_instances``,
        ``pred_fields``; metainfo ``input_center / input_scale / input_size / flip_indices``."""


def instance_data(**fields) -> "InstanceData":
    """One ``InstanceData`` from ready-made fields (a batch of them is built per call on the fused path)."""
    bare = globals().get("_bare")
    return bare(InstanceData, fields) if bare is not None else InstanceData(**fields)
