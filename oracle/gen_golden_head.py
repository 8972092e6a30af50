#!/usr/bin/env python
"""Generate tests/golden/head_kat.npz by running the GENUINE reference head / TTA / codec / estimator.

Runs only in the build container (needs /root/reference).  ``oracle/ref_loader.py`` executes the reference's
unmodified ``probmap_head.py``, ``base_head.py``, ``tta.py``, ``codecs/probmap.py`` (+ ``codecs/utils``),
``pose_estimators/{base,topdown}.py`` under import stubs for mmcv / mmengine / sparsemax (what is stubbed is
listed there; the sparsemax package and the mmpretrain ViT stay unpinned).  This script records, for seeded
weights and inputs:

``small_*``   ProbMapHead(in_channels=24, deconv_out_channels=(16, 16)) - every weight is stored in the file, so the
              pin is self-contained: ``forward`` on two feature batches (pre-sparsemax logits captured by a forward
              hook on ``final_layer``, normalised heatmaps, the four scalars), ``predict`` with and without
              flip_test (all seven InstanceData fields) and ``output_heatmaps``.
``full_*``    the shipped configuration (in_channels=384, deconv 256/256; 18.59 M parameters - regenerated from
              ``probpose_code_b200.synth.make_state_dict(seed=0)``, sha256 of the weights recorded) on one feature
              map + its flipped-pass partner: same outputs.
``est_*``     TopdownPoseEstimator.predict (genuine ``topdown.py:86-194``): flip orchestration on ``inputs.flip(-1)``,
              head.predict, ``add_pred_to_datasample`` to image space, with a tiny ViT-shaped backbone module
              (``oracle.model_oracle.VisionTransformerRef`` - the mmpretrain ViT is not in the reference tree).

    python oracle/gen_golden_head.py            # rewrites tests/golden/head_kat.npz
"""
import hashlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import decode_oracle, model_oracle, ref_loader  # noqa: E402
from probpose_code_b200 import synth  # noqa: E402

FLIP = list(decode_oracle.COCO_FLIP_INDICES)
FIELDS = ("keypoints", "keypoint_scores", "keypoints_conf", "keypoints_probs", "keypoints_visible",
          "keypoints_oks", "keypoints_error")
SMALL = dict(in_channels=24, deconv_out_channels=(16, 16))
TINY_VIT = dict(embed_dims=32, num_layers=2, num_heads=2, feedforward_channels=64)


def weights_sha256(sd):
    h = hashlib.sha256()
    for k in sorted(sd):
        h.update(k.encode())
        h.update(np.ascontiguousarray(sd[k].numpy()).tobytes())
    return h.hexdigest()


def small_head_state(seed=3):
    """Seeded weights for the small head, large enough that heatmaps are peaky and BN is not an identity."""
    full = synth.make_state_dict(seed=seed, arch=dict(embed_dims=SMALL["in_channels"], num_layers=0,
                                                       feedforward_channels=8),
                                 head_std=0.25, final_std=0.08, branch_std=0.08,
                                 deconv_channels=SMALL["deconv_out_channels"])
    return {k[5:]: v for k, v in full.items() if k.startswith("head.")}


def feats_like(batch, channels, seed):
    """Final-LN-like features, stored as fp16-exact fp32 so that the fixture stays small."""
    g = torch.Generator().manual_seed(seed)
    return torch.randn(batch, channels, 16, 12, generator=g).half().float()


def samples(ns, n):
    return [ns.PoseDataSample(metainfo=dict(flip_indices=FLIP)) for _ in range(n)]


def run_head(ns, head, feats, feats_flip, out, tag):
    head.eval()
    logits = []
    hook = head.final_layer.register_forward_hook(lambda m, i, o: logits.append(o.detach().clone()))
    with torch.no_grad():
        for name, f in (("a", feats), ("b", feats_flip)):
            res = head.forward((f,))
            for key, v in zip(("heatmaps", "prob", "vis", "oks", "err"), res):
                out[f"{tag}_fwd_{name}_{key}"] = ref_loader.as_numpy(v)
            out[f"{tag}_fwd_{name}_logits"] = ref_loader.as_numpy(logits.pop())
        hook.remove()
        b = feats.shape[0]
        for flip in (False, True):
            cfg = dict(flip_test=flip, flip_mode="heatmap", shift_heatmap=False, output_heatmaps=True)
            f = [(feats,), (feats_flip,)] if flip else (feats,)
            preds, fields = head.predict(f, samples(ns, b), test_cfg=cfg)
            for key in FIELDS:
                out[f"{tag}_pred_flip{int(flip)}_{key}"] = np.stack([np.asarray(p[key]) for p in preds])
            out[f"{tag}_pred_flip{int(flip)}_heatmaps"] = np.stack([ref_loader.as_numpy(f_.heatmaps) for f_ in fields])


def main():
    torch.manual_seed(0)
    torch.set_num_threads(1)  # one summation order for the convolutions
    ns = ref_loader.load()
    out = {}

    # ---- small head, weights in the file
    sd = small_head_state()
    head = ns.ProbMapHead(**ref_loader.probmap_head_cfg(**SMALL))
    missing = head.load_state_dict(sd, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    for k, v in sd.items():
        out["small_w_" + k] = v.numpy()
    fa, fb = feats_like(2, SMALL["in_channels"], 11), feats_like(2, SMALL["in_channels"], 12)
    out["small_feats_a"], out["small_feats_b"] = fa.half().numpy(), fb.half().numpy()
    run_head(ns, head, fa, fb, out, "small")

    # ---- the shipped configuration, weights regenerated from the seed (sha256 recorded)
    full = synth.make_state_dict(seed=0)
    sdh = {k[5:]: v for k, v in full.items() if k.startswith("head.")}
    head = ns.ProbMapHead(**ref_loader.probmap_head_cfg())
    head.load_state_dict(sdh, strict=True)  # the MMPose key layout of synth == the genuine module's
    out["full_weights_sha256"] = np.array(weights_sha256(sdh))
    fa, fb = feats_like(1, 384, 21), feats_like(1, 384, 22)
    out["full_feats_a"], out["full_feats_b"] = fa.half().numpy(), fb.half().numpy()
    run_head(ns, head, fa, fb, out, "full")

    # ---- estimator: genuine predict + add_pred_to_datasample around a tiny ViT-shaped backbone
    tiny = synth.make_state_dict(seed=5, arch=TINY_VIT, head_std=0.25, final_std=0.08, branch_std=0.08,
                                 deconv_channels=(16, 16))
    backbone = model_oracle.VisionTransformerRef(**TINY_VIT).eval()
    backbone.load_state_dict({k[9:]: v for k, v in tiny.items() if k.startswith("backbone.")})
    est = ns.TopdownPoseEstimator(
        backbone=backbone,
        head=dict(type="ProbMapHead", **ref_loader.probmap_head_cfg(in_channels=32, deconv_out_channels=(16, 16))),
        test_cfg=dict(flip_test=True, flip_mode="heatmap", shift_heatmap=False)).eval()
    est.head.load_state_dict({k[5:]: v for k, v in tiny.items() if k.startswith("head.")}, strict=True)
    for k, v in tiny.items():
        out["est_w_" + k] = v.numpy()
    crops = synth.make_crops(2, seed=9)
    inputs = model_oracle.ProbPoseRef.preprocess(crops)
    out["est_crops"] = crops.numpy()
    centers = np.array([[320.5, 240.25], [100.0, 411.5]], np.float32)
    scales = np.array([[150.0, 200.0], [90.5, 120.75]], np.float32)
    bboxes = np.concatenate([centers - scales / 2.5, centers + scales / 2.5], 1).astype(np.float32)
    out["est_centers"], out["est_scales"], out["est_bboxes"] = centers, scales, bboxes
    ds = []
    for i in range(2):
        ds.append(ns.PoseDataSample(
            metainfo=dict(flip_indices=FLIP, input_center=centers[i], input_scale=scales[i], input_size=(192, 256)),
            gt_instances=ns.InstanceData(bboxes=bboxes[i:i + 1], bbox_scores=np.array([0.9 - 0.1 * i], np.float32))))
    with torch.no_grad():
        res = est.predict(inputs, ds)
    for key in FIELDS + ("bboxes", "bbox_scores"):
        out[f"est_pred_{key}"] = np.stack([np.asarray(r.pred_instances[key]) for r in res])

    path = os.path.join(ROOT, "tests", "golden", "head_kat.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(out), "arrays")
    for k in sorted(out):
        if "_w_" not in k:
            print(f"  {k:44s} {out[k].dtype} {out[k].shape}")


if __name__ == "__main__":
    main()
