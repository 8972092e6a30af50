"""Oracle comparisons AT THE BASELINE.json SIZES (not miniatures): multi-wave schedules - persistent attention with
1 536 units, 12-wave deconv GEMMs, decode of 4 352 maps - are where scheduling bugs hide.

* config 2: 64 crops, flip-TTA, end to end through ``TopdownPoseEstimator.test_step`` vs the CPU oracle
  (fp32 torch restatement pinned to the genuine head by tests/test_oracle_head_pin.py + genuine-pinned decode).
* config 3: the decode kernel at batch 256, plain and flip-TTA, on this model's random-init logits and on planted
  peaks, vs ``decode_oracle.expected_value_decode_batch`` on identical input bits.
* config 5: ViT-B backbone at batch 128 vs torch fp32 on the GPU (TF32 off) for every crop, with the fp32 reference
  itself checked against float64 on a chunk.
* the genuine-reference head fixture (tests/golden/head_kat.npz, shipped configuration) through the product head.

Tolerances: keypoints 1e-3 input px, presence probability and the other scalars 1e-4 (BASELINE.json north_star)."""
import os

import numpy as np
import pytest
import torch

from oracle import cases, model_oracle
from oracle import decode_oracle as d
from probpose_code_b200 import synth

pytestmark = pytest.mark.gpu

KPT_TOL_PX = 1e-3
PROB_TOL = 1e-4
GOLD = os.path.join(os.path.dirname(__file__), "golden", "head_kat.npz")


@pytest.fixture(scope="module")
def small():
    sd = synth.make_state_dict(seed=0)
    ref = model_oracle.ProbPoseRef().eval()
    ref.load_state_dict(sd)
    return dict(sd=sd, ref=ref)


def test_config2_batch64_flip_end_to_end_vs_oracle(small):
    import probpose_code_b200.mmpose_api as api
    crops = synth.make_crops(64, seed=123)
    model = api.MODELS.build(api.probpose_small_cfg(precision="fp16x3"))
    model.load_state_dict(small["sd"])
    model.to("cuda:0")
    out = model.test_step(dict(inputs=[c for c in crops], data_samples=api.make_data_samples(64)))
    assert len(out) == 64
    ref = small["ref"].predict(small["ref"].preprocess(crops), flip_test=True)
    get = lambda name: np.concatenate([getattr(o.pred_instances, name) for o in out], 0)
    assert np.abs(get("keypoints") - ref[..., :2]).max() <= KPT_TOL_PX
    assert np.abs(get("keypoints_probs") - ref[..., 3]).max() <= PROB_TOL
    assert np.abs(get("keypoints_visible") - ref[..., 4]).max() <= PROB_TOL
    assert np.abs(get("keypoints_oks") - ref[..., 5]).max() <= PROB_TOL
    assert np.abs(get("keypoints_error") - ref[..., 6]).max() <= PROB_TOL
    assert np.abs(get("keypoints_conf") - ref[..., 2]).max() <= 1e-4
    assert np.array_equal(get("keypoint_scores"), get("keypoints_oks"))  # probmap_head.py:796-798


def _model_logits(small, n, seed):
    """This model's heatmap logits and branch scalars for n persons, both passes, from the engine (the decode tests
    below compare kernel and oracle on IDENTICAL input bits, so where the logits come from does not matter)."""
    from probpose_code_b200.engine import Engine
    eng = Engine(precision="fp16x3", max_batch=64).load_state_dict(small["sd"])
    crops = synth.make_crops(n, seed=seed)
    x = small["ref"].preprocess(crops)
    lg, sc = [], []
    for flip in (False, True):
        for lo in range(0, n, 64):
            xs = x[lo:lo + 64]
            xs = (xs.flip(-1) if flip else xs).cuda().contiguous()
            a, b = eng.head(eng.backbone(xs))
            lg.append(a.cpu())
            sc.append(b.cpu())
    h = len(lg) // 2
    return (torch.cat(lg[:h]).numpy(), torch.cat(lg[h:]).numpy(), torch.cat(sc[:h]).numpy(), torch.cat(sc[h:]).numpy())


@pytest.mark.parametrize("family", ["model", "planted"])
def test_config3_decode_batch256_vs_oracle(small, family):
    from probpose_code_b200 import ops
    fi = d.COCO_FLIP_INDICES
    if family == "model":
        z, zf, s, sf = _model_logits(small, 256, seed=7)
    else:
        zt, zft = synth.planted_logit_pair(256, seed=4)
        z, zf = zt.numpy(), zft.numpy()
        rng = np.random.default_rng(6)
        s, sf = rng.random((256, 4, 17), dtype=np.float32), rng.random((256, 4, 17), dtype=np.float32)
    assert z.shape == (256, 17, 64, 48)
    cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    # plain
    p = d.heatmaps_from_logits(z)
    locs, vals = d.expected_value_decode_batch(p)
    rec, merged = ops.decode(cu(z), scalars=cu(s), input_is_logits=True, return_heatmaps=True)
    r, m = rec.cpu().numpy(), merged.cpu().numpy()
    np.testing.assert_allclose(m, p, rtol=0, atol=2e-6)
    lk, vk = d.expected_value_decode_batch(m)  # the kernel's own heatmap bits through the oracle: tight
    assert np.abs(r[..., :2] - lk).max() <= 2e-4
    np.testing.assert_array_equal(r[..., 2], vk)
    assert np.abs(r[..., :2] - locs).max() <= 2.4e-4  # = 1e-3 input px
    np.testing.assert_array_equal(r[..., 3:6], s[:, :3].transpose(0, 2, 1))
    # flip-TTA
    pm = d.tta_merge(p, d.heatmaps_from_logits(zf), fi)
    locs, vals = d.expected_value_decode_batch(pm)
    rec, merged = ops.decode(cu(z), cu(zf), fi, cu(s), cu(sf), input_is_logits=True, return_heatmaps=True)
    r, m = rec.cpu().numpy(), merged.cpu().numpy()
    np.testing.assert_allclose(m, pm, rtol=0, atol=2e-6)
    lk, vk = d.expected_value_decode_batch(m)
    assert np.abs(r[..., :2] - lk).max() <= 2e-4
    np.testing.assert_array_equal(r[..., 2], vk)
    assert np.abs(r[..., :2] - locs).max() <= 2.4e-4
    sm = (s + sf[:, :, list(fi)]) * np.float32(0.5)
    np.testing.assert_array_equal(r[..., 3:6], sm[:, :3].transpose(0, 2, 1))
    np.testing.assert_allclose(r[..., 6], sm[:, 3] / np.float32(80.0), rtol=1e-6)


def test_config5_vit_base_batch128_vs_torch_fp32_on_the_gpu():
    from probpose_code_b200.engine import Engine
    sd = synth.make_state_dict(seed=2, arch=synth.VIT_BASE)
    bsd = {k[len("backbone."):]: v for k, v in sd.items() if k.startswith("backbone.")}
    ref = model_oracle.VisionTransformerRef(**synth.VIT_BASE).eval()
    ref.load_state_dict(bsd)
    x = model_oracle.ProbPoseRef.preprocess(synth.make_crops(128, seed=31)).cuda()
    old = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = False
    try:
        with torch.no_grad():
            ref32 = ref.cuda()
            want = torch.cat([ref32(x[i:i + 32])[0] for i in range(0, 128, 32)])
            want64 = ref32.double()(x[:8].double())[0]
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old
    scale = want64.abs().max().item()
    assert ((want[:8].double() - want64).abs().max() / scale).item() <= 1e-5  # the fp32 reference is itself sound
    eng = Engine(precision="fp16x3", max_batch=64, embed_dim=768, heads=12, ffn_dim=3072, deconv_channels=0)
    eng.load_state_dict(sd, prefixes=("backbone.",))
    got = eng.backbone(x.contiguous())  # 128 images = 2 x max_batch single-pass
    assert got.shape == want.shape
    assert ((got - want).abs().max() / scale).item() <= 2e-5
    assert ((got[:8].double() - want64).abs().max() / scale).item() <= 2e-5


def test_product_head_against_the_genuine_reference_fixture():
    """tests/golden/head_kat.npz ``full_*``: outputs of the GENUINE ``ProbMapHead.forward`` / ``predict`` (shipped
    configuration) recorded by oracle/gen_golden_head.py - no restatement in between."""
    import probpose_code_b200.mmpose_api as api
    g = np.load(GOLD)
    sd = synth.make_state_dict(seed=0)
    head = api.MODELS.build(api.probpose_small_cfg(precision="fp16x3")).head
    head.load_state_dict({k[5:]: v for k, v in sd.items() if k.startswith("head.")})
    head.to("cuda:0")
    fa = torch.from_numpy(g["full_feats_a"].astype(np.float32)).cuda()
    fb = torch.from_numpy(g["full_feats_b"].astype(np.float32)).cuda()
    for name, f in (("a", fa), ("b", fb)):
        hm, prob, vis, oks, err = head.forward((f,))
        want = g[f"full_fwd_{name}_heatmaps"]
        np.testing.assert_allclose(hm.cpu().numpy(), want, rtol=0, atol=1e-4)
        for got, key in ((prob, "prob"), (vis, "vis"), (oks, "oks"), (err, "err")):
            assert np.abs(got.cpu().numpy() - g[f"full_fwd_{name}_{key}"]).max() <= PROB_TOL
    samples = api.make_data_samples(1)
    for flip in (0, 1):
        cfg = dict(flip_test=bool(flip), flip_mode="heatmap", shift_heatmap=False, output_heatmaps=True)
        preds, fields = head.predict([(fa,), (fb,)] if flip else (fa,), samples, test_cfg=cfg)
        pre = f"full_pred_flip{flip}_"
        p = preds[0]
        assert np.abs(p.keypoints - g[pre + "keypoints"][0]).max() <= KPT_TOL_PX
        assert p.keypoints.dtype == g[pre + "keypoints"].dtype and p.keypoints.shape == g[pre + "keypoints"][0].shape
        for key in ("keypoint_scores", "keypoints_conf", "keypoints_probs", "keypoints_visible", "keypoints_oks",
                    "keypoints_error"):
            got, want = np.asarray(p[key]), g[pre + key][0]
            assert got.shape == want.shape, key
            assert np.abs(got - want).max() <= PROB_TOL, key
        np.testing.assert_allclose(fields[0].heatmaps.cpu().numpy(), g[pre + "heatmaps"][0], rtol=0, atol=1e-4)


def test_later_frames_with_more_persons_rebuild_the_engine(small):
    """A 10-person frame followed by a 24-person frame (flip_test doubles the images): the cached engine must grow."""
    import probpose_code_b200.mmpose_api as api
    model = api.MODELS.build(api.probpose_small_cfg(precision="fp16x3"))
    model.load_state_dict(small["sd"])
    model.to("cuda:0")
    crops = synth.make_crops(24, seed=5)
    a = model.test_step(dict(inputs=[c for c in crops[:10]], data_samples=api.make_data_samples(10)))
    b = model.test_step(dict(inputs=[c for c in crops], data_samples=api.make_data_samples(24)))
    assert len(a) == 10 and len(b) == 24
    for i in range(10):  # persons are independent units: the same records from either call
        np.testing.assert_array_equal(a[i].pred_instances.keypoints, b[i].pred_instances.keypoints)
