#!/usr/bin/env python
"""The decode-kernel leg of bench.py alone (BASELINE config 3: batch 256, model logits and planted peaks, plain and TTA,
4 rotating sets > L2, back-to-back launches between one event pair).  One JSON line.

    python tools/decode_bench.py [batch] [iters]
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from probpose_code_b200 import synth  # noqa: E402
from probpose_code_b200.engine import Engine  # noqa: E402


def main():
    batch = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    dev = torch.device("cuda", 0)
    eng = Engine(precision="fp16x3", max_batch=64).load_state_dict(synth.make_state_dict(seed=0))
    fi = [0, 2, 1, 4, 3, 6, 5, 8, 7, 10, 9, 12, 11, 14, 13, 16, 15]
    out = bench.decode_leg(eng, dev, fi, batch=batch, iters=iters)
    hbm = bench.load_peaks()["hbm"]
    for v in out.values():
        v["frac"] = v["gbs"] / hbm
    print(json.dumps(dict(batch=batch, hbm_peak=hbm, **out)))


if __name__ == "__main__":
    main()
