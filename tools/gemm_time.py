#!/usr/bin/env python
"""Back-to-back timing of the ViT-S GEMM shapes of a bench step (M = 24576 rows) for the library named by
PROBPOSE_B200_LIB: 4 rotating buffer sets (> L2), 40 launches between one CUDA event pair.

    python tools/gemm_time.py [label] [shape ...]
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from probpose_code_b200 import _lib, ops  # noqa: E402
from tools.gemm_probe import SHAPES  # noqa: E402


def main():
    label = sys.argv[1] if len(sys.argv) > 1 else "base"
    names = sys.argv[2:] or ["qkv", "proj", "fc1", "fc2"]
    prec = _lib.PREC_FP16X3
    tile_n, pair = int(os.environ.get("PP_TILE_N", 0)), int(os.environ.get("PP_CTA_PAIR", 0))  # 0 = the library's choice
    res = {"label": label}
    for name in names:
        m, n, k, kind = SHAPES[name]
        sets = []
        for i in range(4):
            a = ops.to_operand(torch.randn(m, k, device="cuda"), prec)
            out = torch.zeros(m, n, device="cuda") if kind != "gelu_op" else torch.zeros(
                _lib.lib().pp_operand_bytes(prec, m, n), dtype=torch.uint8, device="cuda")
            sets.append((a, out))
        w = ops.to_operand(torch.randn(n, k, device="cuda") * 0.05, prec)
        shift = torch.randn(n, device="cuda")

        def run(i):
            a, out = sets[i % 4]
            if kind == "f32":
                ops.gemm(a, w, m, n, k, prec, shift=shift, out=out, tile_n=tile_n, cta_pair=pair)
            elif kind == "res":
                ops.gemm(a, w, m, n, k, prec, shift=shift, residual=out, out=out, tile_n=tile_n, cta_pair=pair)
            else:
                ops.gemm(a, w, m, n, k, prec, shift=shift, act=_lib.ACT_GELU, out_kind=_lib.OUT_OPERAND, out=out, tile_n=tile_n, cta_pair=pair)

        for i in range(8):
            run(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        iters = 40
        for i in range(iters):
            run(i)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / iters * 1e3
        res[name] = dict(us=round(us, 2), mma_tflops=round(3 * 2.0 * m * n * k / us / 1e6, 1))
    print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
