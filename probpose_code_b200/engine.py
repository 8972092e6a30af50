"""Python handle on ``pp_engine`` (include/probpose_b200.h): owns the device workspace as one
torch uint8 tensor and passes raw pointers + the current CUDA stream across the C ABI.
There is no fallback: every method raises if the library or a CUDA device is missing."""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Sequence

import torch

from . import _lib
from ._lib import check, lib

PIXEL_MEAN = (123.675, 116.28, 103.53)  # RGB, td-pm_ProbPose-small config :53-55
PIXEL_STD = (58.395, 57.12, 57.375)
COCO_FLIP_INDICES = (0, 2, 1, 4, 3, 6, 5, 8, 7, 10, 9, 12, 11, 14, 13, 16, 15)


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


class Engine:
    """ViT backbone and / or ProbMapHead + fused decode on one GPU.

    ``depth=0`` builds a head-only engine, ``deconv_channels=0`` a backbone-only one.
    ``precision``: ``"fp16x3"`` (parity mode, fp32-grade tensor-core GEMMs), ``"bf16"`` /
    ``"fp16"`` (throughput modes) or ``"fp32_simt"`` (CUDA-core verification path).
    """

    def __init__(self, *, precision: str = "fp16x3", max_batch: int = 64, img_size=(256, 192), patch: int = 16,
                 patch_pad: int = 2, embed_dim: int = 384, depth: int = 12, heads: int = 12, ffn_dim: int = 1536,
                 num_keypoints: int = 17, deconv_channels: int = 256, ln_eps: float = 1e-6, bn_eps: float = 1e-5,
                 temperature: float = 0.5, normalize: float = 1.0, mean=PIXEL_MEAN, std=PIXEL_STD,
                 head_kind: str = "probmap", blur_kernel_size: int = 11, device: Optional[torch.device] = None):
        if precision not in _lib.PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(_lib.PRECISIONS)}, got {precision!r}")
        if head_kind not in ("probmap", "heatmap"):
            raise ValueError(f'head_kind must be "probmap" (ProbMapHead) or "heatmap" (HeatmapHead), got {head_kind!r}')
        if not torch.cuda.is_available():
            raise _lib.PPError("probpose_code_b200 needs a CUDA device (sm_100a): there is no CPU path")
        self.head_kind = head_kind
        self.record_floats = _lib.RECORD_FLOATS if head_kind == "probmap" else 3
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.precision = precision
        self.max_batch = int(max_batch)
        self.img_h, self.img_w = int(img_size[0]), int(img_size[1])
        self.embed_dim, self.depth, self.num_keypoints = embed_dim, depth, num_keypoints
        self.deconv_channels = deconv_channels
        self.gh = (self.img_h + 2 * patch_pad - patch) // patch + 1
        self.gw = (self.img_w + 2 * patch_pad - patch) // patch + 1
        self.cfg = _lib.EngineCfg(_lib.PRECISIONS[precision], self.max_batch, self.img_h, self.img_w, patch, patch_pad,
                                  embed_dim, depth, heads, ffn_dim, num_keypoints, deconv_channels, ln_eps, bn_eps,
                                  temperature, 1.0 if normalize is None else float(normalize),
                                  (C.c_float * 3)(*mean), (C.c_float * 3)(*std),
                                  _lib.HEAD_PROBMAP if head_kind == "probmap" else _lib.HEAD_HEATMAP, int(blur_kernel_size))
        nbytes = lib().pp_engine_workspace_bytes(C.byref(self.cfg))
        if nbytes == 0:
            check(-1, "pp_engine_workspace_bytes")
        with torch.cuda.device(self.device):
            self.workspace = torch.empty(nbytes + 1024, dtype=torch.uint8, device=self.device)
            base = (self.workspace.data_ptr() + 1023) & ~1023
            h = C.c_void_p()
            check(lib().pp_engine_create(C.byref(self.cfg), base, nbytes, C.byref(h)), "pp_engine_create")
        self._h = h
        self.workspace_bytes = nbytes
        self._finalized = False

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                lib().pp_engine_destroy(h)
            except Exception:  # noqa: BLE001 - interpreter shutdown
                pass

    # ---- weights ----------------------------------------------------------------------
    def load_state_dict(self, state_dict: Dict[str, torch.Tensor], prefixes: Sequence[str] = ("backbone.", "head.")):
        """Load every tensor whose MMPose name starts with one of ``prefixes`` and finalize.
        Names follow the reference checkpoints."""
        with torch.cuda.device(self.device):
            keep = []
            for name, t in state_dict.items():
                if not name.startswith(tuple(prefixes)) or name.endswith("num_batches_tracked"):
                    continue
                if name.startswith("head.loss") or ".loss_module" in name:
                    continue
                t = t.detach().to(device=self.device, dtype=torch.float32).contiguous()
                keep.append(t)  # alive until the async copies are enqueued on this stream
                check(lib().pp_engine_load(self._h, name.encode(), t.data_ptr(), t.numel(), _stream()),
                      f"pp_engine_load({name})")
            check(lib().pp_engine_finalize(self._h, _stream()), "pp_engine_finalize")
            torch.cuda.current_stream().synchronize()
        self._finalized = True
        return self

    # ---- forward ----------------------------------------------------------------------
    def _check_images(self, x: torch.Tensor, dtype) -> int:
        if not (isinstance(x, torch.Tensor) and x.is_cuda and x.dtype == dtype and x.is_contiguous()):
            raise ValueError(f"input must be a contiguous CUDA {dtype} tensor")
        if x.dim() != 4 or tuple(x.shape[1:]) != (3, self.img_h, self.img_w):
            raise ValueError(f"input must be (B, 3, {self.img_h}, {self.img_w}), got {tuple(x.shape)}")
        return x.shape[0]

    def backbone(self, x: torch.Tensor) -> torch.Tensor:
        """fp32 normalised RGB (B, 3, H, W) -> featmap fp32 (B, C, gh, gw)."""
        b = self._check_images(x, torch.float32)
        out = torch.empty((b, self.embed_dim, self.gh, self.gw), dtype=torch.float32, device=x.device)
        with torch.cuda.device(self.device):
            check(lib().pp_engine_backbone(self._h, x.data_ptr(), b, out.data_ptr(), _stream()), "pp_engine_backbone")
        return out

    def head(self, feat: torch.Tensor):
        """featmap fp32 (B, C, gh, gw) -> (heatmap logits (B, K, 4gh, 4gw), scalars (B, 4, K)); a HeatmapHead engine
        returns the heatmaps (B, K, 4gh, 4gw) alone."""
        if not (feat.is_cuda and feat.dtype == torch.float32 and feat.is_contiguous()):
            raise ValueError("feat must be a contiguous CUDA float32 tensor")
        if tuple(feat.shape[1:]) != (self.embed_dim, self.gh, self.gw):
            raise ValueError(f"feat must be (B, {self.embed_dim}, {self.gh}, {self.gw}), got {tuple(feat.shape)}")
        b = feat.shape[0]
        logits = torch.empty((b, self.num_keypoints, 4 * self.gh, 4 * self.gw), dtype=torch.float32, device=feat.device)
        scal = torch.empty((b, 4, self.num_keypoints), dtype=torch.float32, device=feat.device) if self.head_kind == "probmap" else None
        with torch.cuda.device(self.device):
            check(lib().pp_engine_head(self._h, feat.data_ptr(), b, logits.data_ptr(), None if scal is None else scal.data_ptr(),
                                       _stream()), "pp_engine_head")
        return (logits, scal) if scal is not None else logits

    def infer(self, crops: torch.Tensor, flip_test: bool = True, flip_indices: Sequence[int] = COCO_FLIP_INDICES,
              return_heatmaps: bool = False, out: Optional[torch.Tensor] = None):
        """End to end.  ``crops``: uint8 BGR (B, 3, H, W) (preprocessing fused) or fp32 normalised
        RGB.  Returns records (B, K, 7) fp32 [x_hm, y_hm, conf, prob, vis, oks, err / diag] and
        optionally the merged normalised heatmaps (B, K, 4gh, 4gw).  HeatmapHead engine: records (B, K, 3) =
        [x_hm, y_hm, score] (``pp_decode_udp``)."""
        is_u8 = crops.dtype == torch.uint8
        b = self._check_images(crops, torch.uint8 if is_u8 else torch.float32)
        k = self.num_keypoints
        rec = out if out is not None else torch.empty((b, k, self.record_floats), dtype=torch.float32, device=crops.device)
        merged = torch.empty((b, k, 4 * self.gh, 4 * self.gw), dtype=torch.float32, device=crops.device) if return_heatmaps else None
        fi = None
        if flip_test:
            assert len(flip_indices) == k, "flip_indices length must equal the number of keypoints"
            fi = (C.c_int32 * k)(*[int(i) for i in flip_indices])
        with torch.cuda.device(self.device):
            check(lib().pp_engine_infer(self._h, crops.data_ptr() if is_u8 else None, None if is_u8 else crops.data_ptr(), b,
                                        int(bool(flip_test)), fi, rec.data_ptr(),
                                        None if merged is None else merged.data_ptr(), _stream()), "pp_engine_infer")
        return (rec, merged) if return_heatmaps else rec

    def operand_overflow(self, clear: bool = True) -> bool:
        """True when a kernel on this device clamped an fp16 operand (|activation| > 1023.5 in fp16x3) since the last
        clear: the outputs computed meanwhile are not trustworthy (``pp_operand_overflow``; synchronises the device)."""
        n = C.c_int32(0)
        with torch.cuda.device(self.device):
            check(lib().pp_operand_overflow(int(clear), C.byref(n)), "pp_operand_overflow")
        return n.value > 0

    def raise_on_overflow(self) -> None:
        if self.operand_overflow():
            raise _lib.PPError(
                f"an activation exceeded the operand range of precision {self.precision!r} (|a| > 1023.5 for fp16x3 / 65504 for "
                "fp16) and was clamped: the results of this call are wrong. Build the model with precision='fp32_simt' "
                "(no range limit) for this checkpoint")

    @property
    def last_launch_count(self) -> int:
        return int(lib().pp_engine_last_launch_count(self._h))

    def set_graph(self, max_images: int) -> None:
        """CUDA-graph replay of the workspace-only middle of :meth:`infer` for calls of at most
        ``max_images`` images (flip counts twice); 0 = off, negative = no limit (``pp_engine_set_graph``)."""
        check(lib().pp_engine_set_graph(self._h, int(max_images)), "pp_engine_set_graph")

    @property
    def graph_replay_count(self) -> int:
        return int(lib().pp_engine_graph_replay_count(self._h))

    def profile_begin(self) -> None:
        check(lib().pp_engine_profile_begin(self._h), "pp_engine_profile_begin")

    def profile_end(self) -> dict:
        """Per-kernel-class device time (ms), launch counts and GEMM algorithmic FLOPs since
        :meth:`profile_begin` (synchronises the current stream)."""
        prof = _lib.Profile()
        with torch.cuda.device(self.device):
            check(lib().pp_engine_profile_end(self._h, C.byref(prof), _stream()), "pp_engine_profile_end")
        out = {name: dict(ms=prof.ms[i], launches=int(prof.launches[i])) for i, name in enumerate(_lib.KERNEL_CLASSES)}
        out["gemm_flops"] = prof.gemm_flops
        return out
