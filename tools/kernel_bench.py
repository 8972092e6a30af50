#!/usr/bin/env python
"""Micro-benchmarks of the individual kernels (CUDA events, L2 flushed between iterations).
Writes one JSON line per case to stdout and gpurun_out/kernel_bench.jsonl.

    python tools/kernel_bench.py decode udp gemm
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from probpose_code_b200 import _lib, ops  # noqa: E402

PEAKS = {}
try:
    PEAKS = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
except Exception:  # noqa: BLE001
    pass
HBM = PEAKS.get("hbm_gbs", 6650.0)
TF = PEAKS.get("bf16_tflops", 1590.0)
OUT = os.path.join(ROOT, "gpurun_out", "kernel_bench.jsonl")
_flush = None


def flush_l2():
    global _flush
    if _flush is None:
        _flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    _flush.zero_()


def time_ms(fn, iters=20, warmup=3, flush=True):
    for _ in range(warmup):
        fn()
    ts = []
    for _ in range(iters):
        if flush:
            flush_l2()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def emit(rec):
    line = json.dumps(rec)
    print(line, flush=True)
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    with open(OUT, "a") as f:
        f.write(line + "\n")


def bench_decode():
    from oracle import cases
    fi = [0, 2, 1, 4, 3, 6, 5, 8, 7, 10, 9, 12, 11, 14, 13, 16, 15]
    for name, gen in (("planted", lambda b, s: cases.planted_peak_logits(b, seed=s)),
                      ("noise1e-3", lambda b, s: cases.noise_logits(b, s, 1e-3)),
                      ("noise1", lambda b, s: cases.noise_logits(b, s, 1.0))):
        for batch in (64, 256, 1024):
            if name == "planted":  # the flipped pass of a real model peaks at the mirrored place: consistent pair
                z, zf = (torch.from_numpy(a).cuda() for a in cases.planted_peak_pair(batch, 1))
            else:
                z = torch.from_numpy(gen(batch, 1)).cuda()
                zf = torch.from_numpy(gen(batch, 2)).cuda()
            for tta in (False, True):
                fn = (lambda: ops.decode(z, zf, fi, input_is_logits=True)) if tta else (lambda: ops.decode(z, input_is_logits=True))
                med, best = time_ms(fn)
                nbytes = batch * (17 * 3072 * 4 * (2 if tta else 1) + 17 * 7 * 4)
                emit(dict(kernel="decode", inputs=name, batch=batch, tta=tta, ms_median=med, ms_best=best,
                          gbs=nbytes / med / 1e6, frac_hbm=nbytes / med / 1e6 / HBM))


def bench_udp_and_revert():
    """the DARK-UDP decode and the heatmap read-back kernels."""
    from oracle import revert_oracle, udp_oracle
    fi = [0, 2, 1, 4, 3, 6, 5, 8, 7, 10, 9, 12, 11, 14, 13, 16, 15]
    for batch in (64, 256):
        z = torch.from_numpy(udp_oracle.gaussian_heatmaps(batch, seed=1)).cuda()
        zf = torch.from_numpy(udp_oracle.gaussian_heatmaps(batch, seed=2)).cuda()
        for tta in (False, True):
            fn = (lambda: ops.decode_udp(z, zf, fi)) if tta else (lambda: ops.decode_udp(z))
            med, best = time_ms(fn)
            nbytes = batch * (17 * 3072 * 4 * (2 if tta else 1) + 17 * 3 * 4)
            emit(dict(kernel="decode_udp", batch=batch, tta=tta, ms_median=med, ms_best=best, gbs=nbytes / med / 1e6,
                      frac_hbm=nbytes / med / 1e6 / HBM))
    for persons, ih, iw in ((4, 720, 1280), (16, 1080, 1920)):
        hms, centers, scales = revert_oracle.synthetic_people(5, persons, ih, iw)
        mats = [revert_oracle.get_warp_matrix(c, s, 0, (48, 64), inv=True) for c, s in zip(centers, scales)]
        t = torch.from_numpy(hms).cuda()
        med, best = time_ms(lambda: ops.revert_heatmaps(t, mats, (ih, iw)))
        nbytes = 17 * ih * iw * 4 + hms.nbytes
        emit(dict(kernel="revert_heatmaps", persons=persons, image=f"{ih}x{iw}", ms_median=med, ms_best=best,
                  gbs=nbytes / med / 1e6, frac_hbm=nbytes / med / 1e6 / HBM,
                  note="algorithmic bytes = one write of the merged (17, H, W) tensor + one read of the heatmaps"))


def bench_gemm():
    shapes = [("qkv", 24576, 1152, 384), ("proj", 24576, 384, 384), ("fc1", 24576, 1536, 384), ("fc2", 24576, 384, 1536),
              ("deconv1", 24576, 256, 1536), ("deconv2", 98304, 256, 1024), ("branch1", 24576, 1536, 3456),
              ("vitb_fc1", 24576, 3072, 768), ("square", 8192, 8192, 8192)]
    for prec_name in ("fp16x3", "bf16", "fp16"):
        prec = _lib.PRECISIONS[prec_name]
        for name, m, n, k in shapes:
            a = torch.randn(m, k, device="cuda")
            w = torch.randn(n, k, device="cuda") * 0.05
            ao, wo = ops.to_operand(a, prec), ops.to_operand(w, prec)
            del a, w
            out = torch.empty(m, n, device="cuda")
            for tile_n in (128, 192, 256):
                if n % tile_n and n > tile_n:
                    continue
                for pair in (1, 2):
                    try:
                        med, best = time_ms(lambda: ops.gemm(ao, wo, m, n, k, prec, out=out, tile_n=tile_n, cta_pair=pair), iters=10)
                    except Exception as e:  # noqa: BLE001
                        emit(dict(kernel="gemm", prec=prec_name, shape=name, tile_n=tile_n, pair=pair, error=str(e)))
                        continue
                    fl = 2.0 * m * n * k
                    emit(dict(kernel="gemm", prec=prec_name, shape=name, m=m, n=n, k=k, tile_n=tile_n, pair=pair, ms_median=med,
                              ms_best=best, tflops=fl / med / 1e9, frac_bf16_peak=fl / med / 1e9 / TF))
            del ao, wo, out


if __name__ == "__main__":
    which = sys.argv[1:] or ["decode", "gemm"]
    emit(dict(device=torch.cuda.get_device_name(0), peaks=dict(hbm_gbs=HBM, bf16_tflops=TF)))
    if "decode" in which:
        bench_decode()
    if "udp" in which:
        bench_udp_and_revert()
    if "gemm" in which:
        bench_gemm()
