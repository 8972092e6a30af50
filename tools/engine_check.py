#!/usr/bin/env python
"""Stage-by-stage deviation of the CUDA engine from the fp32 oracle (prints a table; used to
set the tolerances in tests/test_engine_gpu.py).  Needs a GPU.

    python tools/engine_check.py [batch] [precisions...]
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import model_oracle  # noqa: E402
from probpose_code_b200 import synth  # noqa: E402
from probpose_code_b200.engine import Engine  # noqa: E402


def stats(name, got, ref):
    got, ref = np.asarray(got, np.float64), np.asarray(ref, np.float64)
    d = np.abs(got - ref)
    out = dict(stage=name, max_abs=float(d.max()), mean_abs=float(d.mean()), ref_max=float(np.abs(ref).max()),
               rel=float(d.max() / (np.abs(ref).max() + 1e-30)))
    print(json.dumps(out), flush=True)
    return out


def main():
    batch = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    precs = sys.argv[2:] or ["fp32_simt", "fp16x3", "fp16", "bf16"]
    torch.manual_seed(0)
    sd = synth.make_state_dict(seed=0)
    ref = model_oracle.ProbPoseRef().eval()
    missing = ref.load_state_dict(sd, strict=True)
    crops = synth.make_crops(batch, seed=1)
    x = ref.preprocess(crops)
    with torch.no_grad():
        t0 = time.time()
        feat = ref.backbone(x)[0]
        logits = ref.head.heatmap_logits(feat)
        scal = torch.stack([m(feat).flatten(1) for m in (ref.head.probability_layers, ref.head.visibility_layers,
                                                         ref.head.oks_layers, ref.head.error_layers)], 1)
        rec_ref = ref.predict(x, flip_test=True)
        rec_ref_nf = ref.predict(x, flip_test=False)
        print("oracle time %.1fs" % (time.time() - t0), flush=True)
    for prec in precs:
        print("==== precision", prec, flush=True)
        eng = Engine(precision=prec, max_batch=batch).load_state_dict(sd)
        f = eng.backbone(x.cuda())
        stats("backbone.feat", f.cpu().numpy(), feat.numpy())
        lg, sc = eng.head(feat.cuda().contiguous())
        stats("head.logits(oracle feat)", lg.cpu().numpy(), logits.numpy())
        stats("head.scalars(oracle feat)", sc.cpu().numpy(), scal.numpy())
        lg2, sc2 = eng.head(f)
        stats("head.logits(own feat)", lg2.cpu().numpy(), logits.numpy())
        for flip, rr in ((True, rec_ref), (False, rec_ref_nf)):
            for src, name in ((crops.cuda(), "u8"), (x.cuda(), "f32")):
                rec = eng.infer(src, flip_test=flip).cpu().numpy().astype(np.float64)
                kp = rec[..., :2] / [47, 63] * [192, 256]
                dk = np.abs(kp - rr[..., :2])
                print(json.dumps(dict(stage=f"infer flip={flip} in={name}", kpt_max_px=float(dk.max()),
                                      kpt_median_px=float(np.median(dk)), kpt_frac_gt_1e3=float((dk > 1e-3).mean()),
                                      conf_max=float(np.abs(rec[..., 2] - rr[..., 2]).max()),
                                      prob_max=float(np.abs(rec[..., 3] - rr[..., 3]).max()),
                                      vis_max=float(np.abs(rec[..., 4] - rr[..., 4]).max()),
                                      oks_max=float(np.abs(rec[..., 5] - rr[..., 5]).max()),
                                      err_max=float(np.abs(rec[..., 6] - rr[..., 6]).max()),
                                      launches=eng.last_launch_count)), flush=True)
        # timing (informational)
        big = synth.make_crops(min(64, eng.max_batch), seed=2).cuda()
        for _ in range(2):
            eng.infer(big[:batch], flip_test=True)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            eng.infer(big[:batch], flip_test=True)
        b.record()
        torch.cuda.synchronize()
        print(json.dumps(dict(stage="timing", batch=batch, ms=a.elapsed_time(b) / 5)), flush=True)
        del eng
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
