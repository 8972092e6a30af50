#!/usr/bin/env python
"""Launch the fused decode kernel a few times on one input family (for ncu captures).

    python tools/decode_probe.py planted|noise1|flat|pair (+ _alt: alternate plain / TTA launches) [batch] [tta] [iters]
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import cases  # noqa: E402  (input generators only)
from probpose_code_b200 import ops  # noqa: E402


def main():
    fam = sys.argv[1] if len(sys.argv) > 1 else "planted"
    fam_alt = fam.endswith("_alt")
    fam = fam.replace("_alt", "")
    batch = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    tta = (sys.argv[3] if len(sys.argv) > 3 else "0") not in ("0", "false")
    iters = int(sys.argv[4]) if len(sys.argv) > 4 else 3
    gen = {"planted": lambda s: cases.planted_peak_logits(batch, seed=s),
           "noise1": lambda s: cases.noise_logits(batch, s, 1.0),
           "flat": lambda s: cases.noise_logits(batch, s, 1e-3), "pair": None}[fam]
    pair = fam == "pair"
    if pair:
        a, b = cases.planted_peak_pair(batch, 1)
        z, zf2 = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    else:
        z = torch.from_numpy(gen(1)).cuda()
    zf = (zf2 if pair else torch.from_numpy(gen(2)).cuda()) if tta else None
    fi = [0, 2, 1, 4, 3, 6, 5, 8, 7, 10, 9, 12, 11, 14, 13, 16, 15]
    if fam_alt:
        zf = zf2 if pair else torch.from_numpy(gen(2)).cuda()
        for _ in range(iters):  # alternate plain / TTA launches (tools/ncu_traffic.py expects this order)
            ops.decode(z, input_is_logits=True)
            ops.decode(z, zf, fi, input_is_logits=True)
    else:
        for _ in range(iters):
            ops.decode(z, zf, fi if tta else None, input_is_logits=True)
    torch.cuda.synchronize()


if __name__ == "__main__":
    main()
