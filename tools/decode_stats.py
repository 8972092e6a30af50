#!/usr/bin/env python
"""Shape of the decode kernel's work on the two bench input families (model logits of the random-init head, planted
peaks): per map the number of sparsemax candidates, the support size after the merge, the area of the search box and
the pixel count of the one-hop boxes.  These decide which branch of decode.cu a map takes; printed as percentiles.

    python tools/decode_stats.py [persons]
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from probpose_code_b200 import ops, synth  # noqa: E402
from probpose_code_b200.engine import Engine  # noqa: E402

SIGMAS = [0.026, 0.025, 0.025, 0.035, 0.035, 0.079, 0.079, 0.072, 0.072, 0.062, 0.062, 0.107, 0.107, 0.087, 0.087, 0.089, 0.089]
H, W = 64, 48


def radius(k):
    s = (SIGMAS[k] * 2) ** 2 * np.sqrt(H / 1.25 * W / 1.25) * 2
    return int(np.ceil(min(max(s, 0.55), 3.0) * 3))


def extend(r, y0, y1, x0, x1):
    return (0 if y0 <= r - 1 else y0, H - 1 if y1 >= H - r else y1, 0 if x0 <= r - 1 else x0, W - 1 if x1 >= W - r else x1)


def stats(z, merged, temperature=0.5):
    rows = []
    B, K = z.shape[:2]
    for b in range(B):
        for k in range(K):
            r = radius(k)
            zz = z[b, k]
            cand = int((zz > zz.max() - temperature).sum())
            ys, xs = np.nonzero(merged[b, k])
            nnz = len(ys)
            y0, y1, x0, x1 = extend(r, ys.min(), ys.max(), xs.min(), xs.max())
            area = (y1 - y0 + 1) * (x1 - x0 + 1)
            boxes = []
            for i in range(nnz):
                near = (np.abs(ys - ys[i]) <= 2 * r) & (np.abs(xs - xs[i]) <= 2 * r)
                boxes.append(extend(r, ys[near].min(), ys[near].max(), xs[near].min(), xs[near].max()))
            keep = []
            for i, bx in enumerate(boxes):
                inside = any(j != i and o[0] <= bx[0] and o[1] >= bx[1] and o[2] <= bx[2] and o[3] >= bx[3] and (o != bx or j < i)
                             for j, o in enumerate(boxes))
                if not inside:
                    keep.append(bx)
            hop = sum((bx[1] - bx[0] + 1) * (bx[3] - bx[2] + 1) for bx in keep)
            rows.append((cand, nnz, area, hop, len(keep), r))
    return np.array(rows)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    dev = torch.device("cuda", 0)
    eng = Engine(precision="fp16x3", max_batch=64).load_state_dict(synth.make_state_dict(seed=0))
    mean = torch.tensor(bench.eng_mean(), device=dev).view(1, 3, 1, 1)
    std = torch.tensor(bench.eng_std(), device=dev).view(1, 3, 1, 1)
    crops = synth.make_crops(n, seed=5000).to(dev)
    x = ((crops[:, [2, 1, 0]].float() - mean) / std).contiguous()
    z = eng.head(eng.backbone(x))[0]
    zf = eng.head(eng.backbone(x.flip(-1).contiguous()))[0]
    fi = [0, 2, 1, 4, 3, 6, 5, 8, 7, 10, 9, 12, 11, 14, 13, 16, 15]
    pz, pzf = synth.planted_logit_pair(n, seed=7000, device=dev)
    out = {}
    for name, a, b in (("model", z, zf), ("planted", pz, pzf)):
        for tta in (False, True):
            _, merged = (ops.decode(a, b, fi, input_is_logits=True, return_heatmaps=True) if tta
                         else ops.decode(a, input_is_logits=True, return_heatmaps=True))
            st = stats(a.cpu().numpy(), merged.cpu().numpy(), temperature=0.5)
            q = lambda c: [float(v) for v in np.percentile(st[:, c], [10, 50, 90, 99, 100])]
            out[f"{name}_{'tta' if tta else 'plain'}"] = dict(candidates=q(0), nnz=q(1), box_area=q(2), hop_pixels=q(3), hop_boxes=q(4),
                                                              radius=q(5), frac_nnz_le32=float(np.mean(st[:, 1] <= 32)),
                                                              frac_area_le192=float(np.mean(st[:, 2] <= 192)),
                                                              frac_hop_le96=float(np.mean(st[:, 3] <= 96)),
                                                              frac_hop_le384=float(np.mean(st[:, 3] <= 384)))
    for k_, v_ in out.items():
        print(k_, json.dumps(v_))


if __name__ == "__main__":
    main()
