#!/usr/bin/env python
"""BASELINE config 5: ViTPose-base backbone (D 768, 12 layers, 12 heads of 64, FFN 3072), 256x192,
batch 128 on one B200 - tensor-pipe utilisation of the backbone alone, per precision mode.

    python tools/vitb_bench.py [batch] [precisions...]
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from probpose_code_b200 import synth  # noqa: E402
from probpose_code_b200.engine import Engine  # noqa: E402

GFLOP_PER_CROP = {"base": 34.2004, "small": 8.9465}


def main():
    batch = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    precs = sys.argv[2:] or ["fp16x3", "fp16", "bf16"]
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    peak = peaks.get("bf16_tflops_sustained", 1400.0)
    for arch_name, arch in (("base", synth.VIT_BASE), ("small", synth.VIT_SMALL)):
        sd = synth.make_state_dict(seed=2, arch=arch)
        for prec in precs:
            eng = Engine(precision=prec, max_batch=batch, embed_dim=arch["embed_dims"], heads=arch["num_heads"],
                         ffn_dim=arch["feedforward_channels"], deconv_channels=0)
            eng.load_state_dict(sd, prefixes=("backbone.",))
            xs = [torch.randn(batch, 3, 256, 192, device="cuda") for _ in range(3)]
            for i in range(3):
                eng.backbone(xs[i])
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            iters = 10
            a.record()
            for i in range(iters):
                eng.backbone(xs[i % 3])
            b.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b) / iters
            tf = GFLOP_PER_CROP[arch_name] * batch / ms
            mma = tf * (3 if prec == "fp16x3" else 1)
            print(json.dumps(dict(workload=f"ViT-{arch_name} backbone 256x192 batch={batch}", precision=prec, ms=ms,
                                  crops_per_s=batch / ms * 1e3, algorithmic_tflops=tf, frac_of_sustained_bf16=tf / peak,
                                  mma_tflops=mma, mma_frac=mma / peak)), flush=True)
            del eng
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
