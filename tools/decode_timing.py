#!/usr/bin/env python
"""Per-phase cycle totals of the sparse decode (needs a library built with PP_NVCC_EXTRA=-DPP_DECODE_TIMING).

    PP_NVCC_EXTRA=-DPP_DECODE_TIMING python -m probpose_code_b200.build --force
    python tools/decode_timing.py [batch]
"""
import ctypes as C
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from probpose_code_b200 import _lib, ops, synth  # noqa: E402
from probpose_code_b200.engine import Engine  # noqa: E402

NAMES = ["items", "wait", "scan1", "scan2", "lists", "setup", "eval", "after_scans"]


def read(reset=True):
    buf = (C.c_ulonglong * 16)()
    fn = _lib.lib().pp_debug_decode_timing
    fn.argtypes = [C.POINTER(C.c_ulonglong), C.c_int]
    assert fn(buf, int(reset)) == 0
    return list(buf)


def main():
    batch = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    dev = torch.device("cuda", 0)
    fi = [0, 2, 1, 4, 3, 6, 5, 8, 7, 10, 9, 12, 11, 14, 13, 16, 15]
    eng = Engine(precision="fp16x3", max_batch=64).load_state_dict(synth.make_state_dict(seed=0))
    mean = torch.tensor([123.675, 116.28, 103.53], device=dev).view(1, 3, 1, 1)
    std = torch.tensor([58.395, 57.12, 57.375], device=dev).view(1, 3, 1, 1)
    zs, zfs = [], []
    for c0 in range(0, batch, 64):
        crops = synth.make_crops(64, seed=5000 + c0).to(dev)
        x = ((crops[:, [2, 1, 0]].float() - mean) / std).contiguous()
        zs.append(eng.head(eng.backbone(x))[0])
        zfs.append(eng.head(eng.backbone(x.flip(-1).contiguous()))[0])
    z, zf = torch.cat(zs).contiguous(), torch.cat(zfs).contiguous()
    pz, pzf = synth.planted_logit_pair(batch, seed=7000, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for name, a, b in (("model_plain", z, None), ("model_tta", z, zf), ("planted_plain", pz, None), ("planted_tta", pz, pzf)):
        for _ in range(2):
            flush.zero_()
            ops.decode(a, b, fi if b is not None else None, input_is_logits=True)
        read()
        iters = 5
        for _ in range(iters):
            flush.zero_()
            ops.decode(a, b, fi if b is not None else None, input_is_logits=True)
        t = read()
        n = max(t[0], 1)
        print(json.dumps(dict(family=name, items=t[0] // iters, cycles_per_item={k: round(v / n) for k, v in zip(NAMES[1:], t[1:8])},
                              big_lists=dict(maps=t[8] // iters, cycles_each=round(t[9] / max(t[8], 1))),
                              tile=dict(maps=t[10] // iters, cycles_each=round(t[11] / max(t[10], 1))),
                              dense=dict(maps=t[12] // iters, cta_cycles_each=round(t[13] / max(t[12], 1))),
                              longest_map_after_scans=t[14], longest_cta=t[15])))


if __name__ == "__main__":
    main()
