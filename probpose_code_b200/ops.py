"""Thin Python wrappers over the C ABI: torch tensors in/out only at this boundary
(``data_ptr()`` + shape + current CUDA stream).  No fallback paths."""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional, Sequence

import torch

from . import _lib
from ._lib import check, lib


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _need_cuda(t: torch.Tensor, name: str, dtype=torch.float32) -> None:
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor, got {type(t).__name__}")
    if not t.is_cuda:
        raise _lib.PPError(f"{name} must live on a CUDA device: probpose_code_b200 has no CPU path")
    if t.dtype != dtype:
        raise ValueError(f"{name} must be {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"{name} must be contiguous")


def decode(maps: torch.Tensor, maps_flip: Optional[torch.Tensor] = None,
           flip_indices: Optional[Sequence[int]] = None, scalars: Optional[torch.Tensor] = None,
           scalars_flip: Optional[torch.Tensor] = None, *, input_is_logits: bool, temperature: float = 0.5,
           normalize: float = 1.0, return_heatmaps: bool = False, out: Optional[torch.Tensor] = None):
    """Fused decode (``pp_decode``).  ``maps`` (B, K, H, W) fp32 CUDA.  Returns records
    (B, K, 7) fp32 [x_hm, y_hm, conf, prob, vis, oks, err/diag] and, optionally, the merged
    normalised heatmaps (B, K, H, W)."""
    _need_cuda(maps, "maps")
    if maps.dim() != 4:
        raise ValueError(f"maps must be (B, K, H, W), got {tuple(maps.shape)}")
    b, k, h, w = maps.shape
    for name, t in (("maps_flip", maps_flip),):
        if t is not None:
            _need_cuda(t, name)
            if t.shape != maps.shape:
                raise ValueError(f"{name} shape {tuple(t.shape)} != maps shape {tuple(maps.shape)}")
    for name, t in (("scalars", scalars), ("scalars_flip", scalars_flip)):
        if t is not None:
            _need_cuda(t, name)
            if tuple(t.shape) != (b, 4, k):
                raise ValueError(f"{name} must be (B, 4, K) = {(b, 4, k)}, got {tuple(t.shape)}")
    fi = None
    if maps_flip is not None:
        if flip_indices is None:
            raise ValueError("flip_indices are required with maps_flip")
        assert len(flip_indices) == k, "flip_indices length must equal the number of keypoints"
        fi = (C.c_int32 * k)(*[int(i) for i in flip_indices])
    cfg = _lib.DecodeCfg(k, h, w, int(input_is_logits), float(temperature), float(normalize),
                         float(math.sqrt(h * h + w * w)))
    rec = out if out is not None else torch.empty((b, k, _lib.RECORD_FLOATS), dtype=torch.float32, device=maps.device)
    if out is not None:
        _need_cuda(out, "out")
        if tuple(out.shape) != (b, k, _lib.RECORD_FLOATS):
            raise ValueError("out must be (B, K, 7)")
    merged = torch.empty_like(maps) if return_heatmaps else None
    with torch.cuda.device(maps.device):
        check(lib().pp_decode(C.byref(cfg), _ptr(maps), _ptr(maps_flip), fi, _ptr(scalars), _ptr(scalars_flip),
                              b, _ptr(rec), _ptr(merged), _stream()), "pp_decode")
    return (rec, merged) if return_heatmaps else rec


def decode_udp(maps: torch.Tensor, maps_flip: Optional[torch.Tensor] = None, flip_indices: Optional[Sequence[int]] = None, *,
               blur_kernel_size: int = 11, return_heatmaps: bool = False, out: Optional[torch.Tensor] = None):
    """``pp_decode_udp``: fused flip-TTA merge + UDPHeatmap (DARK-UDP) decode.  ``maps`` fp32 CUDA (B, K, 64, 48).
    Returns records fp32 (B, K, 3) = x, y in heatmap pixels, score (and the merged heatmaps if asked)."""
    _need_cuda(maps, "maps")
    if maps.dim() != 4:
        raise ValueError(f"maps must be (B, K, H, W), got {tuple(maps.shape)}")
    b, k, h, w = maps.shape
    fi = None
    if maps_flip is not None:
        _need_cuda(maps_flip, "maps_flip")
        if maps_flip.shape != maps.shape:
            raise ValueError("maps_flip must have the shape of maps")
        if flip_indices is None or len(flip_indices) != k:
            raise ValueError("flip_indices must list one partner per keypoint")  # heatmap_head.py:247, tta.py:37
        fi = (C.c_int32 * k)(*[int(i) for i in flip_indices])
    cfg = _lib.UdpCfg(k, h, w, int(blur_kernel_size))
    rec = out if out is not None else torch.empty((b, k, 3), dtype=torch.float32, device=maps.device)
    merged = torch.empty_like(maps) if return_heatmaps else None
    with torch.cuda.device(maps.device):
        check(lib().pp_decode_udp(C.byref(cfg), _ptr(maps), _ptr(maps_flip), fi, b, _ptr(rec), _ptr(merged), _stream()),
              "pp_decode_udp")
    return (rec, merged) if return_heatmaps else rec


def to_operand(x: torch.Tensor, precision: int) -> torch.Tensor:
    """fp32 (rows, k) CUDA -> GEMM operand buffer (uint8 tensor) in ``precision``."""
    _need_cuda(x, "x")
    rows, k = x.shape
    buf = torch.empty(lib().pp_operand_bytes(precision, rows, k), dtype=torch.uint8, device=x.device)
    with torch.cuda.device(x.device):
        check(lib().pp_operand_from_f32(precision, x.data_ptr(), rows, k, k, buf.data_ptr(), _stream()),
              "pp_operand_from_f32")
    return buf


def gemm(a_op: torch.Tensor, w_op: torch.Tensor, m: int, n: int, k: int, precision: int, *,
         scale: Optional[torch.Tensor] = None, shift: Optional[torch.Tensor] = None,
         residual: Optional[torch.Tensor] = None, act: int = _lib.ACT_NONE, out_kind: int = _lib.OUT_F32,
         out: Optional[torch.Tensor] = None, ldd: Optional[int] = None, plane: int = 0, up=None, tile_n: int = 0,
         res_mod: int = 0, taps: Optional[Sequence[int]] = None, in_pad=None, out_pad: bool = False,
         out_rows: Optional[int] = None, cta_pair: int = 0, shared_border: bool = False):
    """``pp_gemm``: D = epilogue(A . W^T).  ``a_op`` / ``w_op`` are operand buffers from
    :func:`to_operand`.  Returns the output tensor (fp32, or a uint8 operand buffer).

    ``taps``: row shifts of the implicit-GEMM A operand (logical width ``k / len(taps)``);
    ``in_pad=(h, w)``: the A rows enumerate a zero-padded ``(h + 2, w + 2)`` map; ``out_pad``:
    the output map carries a border too; ``shared_border``: both use the ``(h + 1, w + 1)`` layout (``in_pad`` /
    ``out_pad`` = 2 in the C ABI) instead; ``out_rows``: rows of a freshly allocated output."""
    dev = a_op.device
    if out_kind == _lib.OUT_F32:
        ldd = n if ldd is None else ldd
        rows = out_rows if out_rows is not None else (m if up is None else m * 4)
        if out is None:
            out = torch.zeros((rows, ldd), dtype=torch.float32, device=dev)
    elif out_kind == _lib.OUT_OPERAND:
        ldd = n if ldd is None else ldd
        rows = out_rows if out_rows is not None else (m if up is None else m * 4)
        if out is None:
            out = torch.zeros(lib().pp_operand_bytes(precision, rows, ldd), dtype=torch.uint8, device=dev)
    else:
        assert plane > 0 and m % plane == 0
        ldd = 0
        if out is None:
            out = torch.empty((m // plane, n, plane), dtype=torch.float32, device=dev)
    hin, win, py, px = up if up is not None else (0, 0, 0, 0)
    ntaps = 0 if taps is None else len(taps)
    shifts = (C.c_int32 * 9)(*([int(t) for t in taps] + [0] * (9 - ntaps))) if taps is not None else (C.c_int32 * 9)()
    ih, iw = in_pad if in_pad is not None else (0, 0)
    args = _lib.GemmArgs(precision, m, n, k, a_op.data_ptr(), w_op.data_ptr(), _ptr(scale), _ptr(shift),
                         _ptr(residual), act, out_kind, out.data_ptr(), ldd, plane, hin, win, py, px, tile_n, res_mod,
                         ntaps, shifts, int(in_pad is not None) * (2 if shared_border else 1), ih, iw,
                         int(bool(out_pad)) * (2 if shared_border else 1), int(cta_pair))
    with torch.cuda.device(dev):
        check(lib().pp_gemm(C.byref(args), _stream()), "pp_gemm")
    return out


def revert_heatmaps(heatmaps: torch.Tensor, warp_mats, img_shape) -> torch.Tensor:
    """``pp_revert_heatmaps``: heatmaps fp32 CUDA (P, K, H, W) + the (P, 2, 3) float64 heatmap -> image matrices of
    ``revert_heatmap`` -> the max over persons of the warped maps, fp32 (K, img_h, img_w) (structures/utils.py:117,146-175)."""
    _need_cuda(heatmaps, "heatmaps")
    if heatmaps.dim() != 4:
        raise ValueError(f"heatmaps must be (P, K, H, W), got {tuple(heatmaps.shape)}")
    p, k, h, w = heatmaps.shape
    import numpy as np
    mats = torch.as_tensor(np.asarray(warp_mats, dtype=np.float64)).reshape(-1, 2, 3)
    if mats.shape[0] != p:
        raise ValueError("one warp matrix per person")
    mats = mats.to(heatmaps.device).contiguous()
    img_h, img_w = int(img_shape[0]), int(img_shape[1])
    out = torch.empty((k, img_h, img_w), dtype=torch.float32, device=heatmaps.device)
    scratch = torch.empty((max(p, 1), 10), dtype=torch.float64, device=heatmaps.device)
    with torch.cuda.device(heatmaps.device):
        check(lib().pp_revert_heatmaps(heatmaps.data_ptr(), mats.data_ptr(), p, k, h, w, out.data_ptr(), img_h, img_w,
                                       scratch.data_ptr(), _stream()), "pp_revert_heatmaps")
    return out


def attention(qkv_op: torch.Tensor, batch: int, tokens: int, heads: int, head_dim: int, precision: int, impl: int = 0,
              out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``pp_attention``: softmax(Q K^T d_h^-0.5) V per (image, head) on the tensor cores.  ``qkv_op`` is the operand
    buffer (batch * tokens, 3 * heads * head_dim) of the qkv GEMM; returns the operand buffer (batch * tokens, heads * head_dim).
    ``impl``: 0 default, 1 mma.sync kernel, 2 tcgen05 kernel."""
    _need_cuda(qkv_op, "qkv_op", torch.uint8)
    if out is None:
        out = torch.zeros(lib().pp_operand_bytes(precision, batch * tokens, heads * head_dim), dtype=torch.uint8, device=qkv_op.device)
    with torch.cuda.device(qkv_op.device):
        check(lib().pp_attention(precision, qkv_op.data_ptr(), batch, tokens, heads, head_dim, out.data_ptr(), impl, _stream()),
              "pp_attention")
    return out


def from_operand(buf: torch.Tensor, rows: int, k: int, precision: int) -> torch.Tensor:
    """Operand buffer -> fp32 (rows, k) (tests / debugging; the inverse of :func:`to_operand`)."""
    if precision == _lib.PREC_FP16X3:
        h = buf.view(torch.float16).view(rows, 2 * k).float()
        return (h[:, :k] + h[:, k:]) / 64.0
    if precision == _lib.PREC_BF16:
        return buf.view(torch.bfloat16).view(rows, k).float()
    if precision == _lib.PREC_FP16:
        return buf.view(torch.float16).view(rows, k).float()
    return buf.view(torch.float32).view(rows, k).clone()


def crop_warp(frame: torch.Tensor, warp_mats: torch.Tensor, out_hw=(256, 192), out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``pp_crop_warp``: frame uint8 BGR (H, W, 3) CUDA + forward affine matrices fp32 (N, 2, 3) CUDA ->
    crops uint8 BGR (N, 3, h, w), bit-identical to ``cv2.warpAffine(..., flags=cv2.INTER_LINEAR)`` per person
    followed by HWC -> CHW (topdown_transforms.py:126, formatting.py)."""
    _need_cuda(frame, "frame", torch.uint8)
    _need_cuda(warp_mats, "warp_mats", torch.float32)
    if frame.dim() != 3 or frame.shape[2] != 3:
        raise ValueError(f"frame must be (H, W, 3) uint8 BGR, got {tuple(frame.shape)}")
    if warp_mats.dim() != 3 or tuple(warp_mats.shape[1:]) != (2, 3):
        raise ValueError(f"warp_mats must be (N, 2, 3), got {tuple(warp_mats.shape)}")
    n, (h, w) = warp_mats.shape[0], out_hw
    crops = out if out is not None else torch.empty((n, 3, h, w), dtype=torch.uint8, device=frame.device)
    if out is not None:
        _need_cuda(out, "out", torch.uint8)
        if tuple(out.shape) != (n, 3, h, w):
            raise ValueError(f"out must be {(n, 3, h, w)}")
    with torch.cuda.device(frame.device):
        check(lib().pp_crop_warp(frame.data_ptr(), frame.shape[0], frame.shape[1], frame.stride(0), warp_mats.data_ptr(), n,
                                 crops.data_ptr(), h, w, _stream()), "pp_crop_warp")
    return crops
