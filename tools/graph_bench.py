"""Latency of pp_engine_infer per call, plain launches vs CUDA-graph replay (pp_engine_set_graph), over batch sizes.

    python tools/graph_bench.py [--out gpurun_out/graph_bench.jsonl]

Crops resident in HBM; per batch size: 10 warm-up calls, then `iters` back-to-back calls between one CUDA event pair
(device time per call) and the wall clock around the same loop with a final synchronize (host + device).
"""
import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import probpose_code_b200.mmpose_api as api  # noqa: E402
from probpose_code_b200 import synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/graph_bench.jsonl")
    ap.add_argument("--iters", type=int, default=100)
    ap.add_argument("--batches", default="1,2,4,8,16,32,64")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    model = api.MODELS.build(api.probpose_small_cfg(precision="fp16x3", flip_test=True))
    model.load_state_dict(synth.make_state_dict(seed=0))
    model.to(dev)
    eng = model._fused_engine(128, dev)
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    with open(args.out, "w") as f:
        for b in [int(x) for x in args.batches.split(",")]:
            crops = [synth.make_crops(b, seed=300 + i).to(dev) for i in range(4)]
            rec = torch.empty((b, 17, 7), dtype=torch.float32, device=dev)
            row = dict(batch=b, flip_test=True, precision="fp16x3", iters=args.iters)
            for name, mode in (("plain", 0), ("graph", -1)):
                eng.set_graph(mode)
                for i in range(10):
                    eng.infer(crops[i % 4], flip_test=True, out=rec)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                t0 = time.perf_counter()
                e0.record()
                for i in range(args.iters):
                    eng.infer(crops[i % 4], flip_test=True, out=rec)
                e1.record()
                t_issue = time.perf_counter() - t0
                torch.cuda.synchronize()
                wall = time.perf_counter() - t0
                row[name] = dict(device_ms=e0.elapsed_time(e1) / args.iters, wall_ms=wall / args.iters * 1e3,
                                 host_issue_ms=t_issue / args.iters * 1e3, launches=eng.last_launch_count)
            row["graph_replays"] = eng.graph_replay_count
            row["speedup_wall"] = row["plain"]["wall_ms"] / row["graph"]["wall_ms"]
            print(json.dumps(row), flush=True)
            f.write(json.dumps(row) + "\n")


if __name__ == "__main__":
    main()
