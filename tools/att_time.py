"""Back-to-back timing of pp_attention (ViT-S shape of a bench step: 128 images, 12 heads, d_h 32) for the library
named by PROBPOSE_B200_LIB (kernel-variant experiments).  python tools/att_time.py [label]"""
import os
import sys
import json

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from probpose_code_b200 import _lib, ops  # noqa: E402


def main():
    label = sys.argv[1] if len(sys.argv) > 1 else "base"
    B, N, H, DH = 128, 192, 12, 32
    out = {"label": label}
    for prec_name in ("fp16x3", "fp16"):
        prec = _lib.PRECISIONS[prec_name]
        torch.manual_seed(0)
        bufs = [ops.to_operand(torch.randn(B * N, 3 * H * DH, device="cuda"), prec) for _ in range(3)]  # 3 x 113 MB > L2
        o = ops.attention(bufs[0], B, N, H, DH, prec, impl=2)
        for i in range(6):
            ops.attention(bufs[i % 3], B, N, H, DH, prec, impl=2, out=o)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        iters = 60
        for i in range(iters):
            ops.attention(bufs[i % 3], B, N, H, DH, prec, impl=2, out=o)
        e1.record()
        torch.cuda.synchronize()
        out[prec_name + "_us"] = round(e0.elapsed_time(e1) / iters * 1e3, 2)
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
