#!/usr/bin/env python
"""Launch a few tcgen05 GEMMs of one shape (for ncu captures).

    python tools/gemm_probe.py qkv|proj|fc1|fc2|square [precision] [iters]
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from probpose_code_b200 import _lib, ops  # noqa: E402

SHAPES = {"qkv": (24576, 1152, 384, "f32"), "proj": (24576, 384, 384, "res"), "fc1": (24576, 1536, 384, "gelu_op"),
          "fc2": (24576, 384, 1536, "res"), "square": (8192, 8192, 8192, "f32"), "branch1": (24576, 1536, 3456, "f32")}


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "qkv"
    prec = _lib.PRECISIONS[sys.argv[2] if len(sys.argv) > 2 else "fp16x3"]
    iters = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    m, n, k, kind = SHAPES[name]
    a = ops.to_operand(torch.randn(m, k, device="cuda"), prec)
    w = ops.to_operand(torch.randn(n, k, device="cuda") * 0.05, prec)
    shift = torch.randn(n, device="cuda")
    out = torch.zeros(m, n, device="cuda")
    for _ in range(iters):
        if kind == "f32":
            ops.gemm(a, w, m, n, k, prec, shift=shift, out=out)
        elif kind == "res":
            ops.gemm(a, w, m, n, k, prec, shift=shift, residual=out, out=out)
        else:
            ops.gemm(a, w, m, n, k, prec, shift=shift, act=_lib.ACT_GELU, out_kind=_lib.OUT_OPERAND)
    torch.cuda.synchronize()


if __name__ == "__main__":
    main()
