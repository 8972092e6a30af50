"""The MMPose-shaped plugin classes on the GPU: same calls, shapes, dtypes and values as the
reference's predict path (patterned on the reference's tests/test_models/test_heads/
test_heatmap_heads/test_heatmap_head.py, test_pose_estimators/test_topdown.py and
test_codecs/test_udp_heatmap.py, but with value checks against the oracle)."""
import numpy as np
import pytest
import torch

import probpose_code_b200.mmpose_api as api
from oracle import cases, model_oracle
from oracle import decode_oracle as d
from probpose_code_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def world():
    sd = synth.make_state_dict(seed=0)
    ref = model_oracle.ProbPoseRef().eval()
    ref.load_state_dict(sd)
    crops = synth.make_crops(3, seed=1)
    model = api.MODELS.build(api.probpose_small_cfg(precision="fp16x3"))
    model.load_state_dict(sd)
    model.cuda()
    return dict(sd=sd, ref=ref, crops=crops, x=ref.preprocess(crops), model=model)


def test_codec_decode_and_batch_decode(golden):
    codec = api.KEYPOINT_CODECS.build(dict(type="ProbMap", input_size=(192, 256), heatmap_size=(48, 64), sigma=-1))
    hms = d.heatmaps_from_logits(cases.planted_peak_logits(12, seed=0))
    k, s = codec.decode(hms[0])
    assert k.shape == (1, 17, 2) and k.dtype == np.float64 and s.shape == (1, 17) and s.dtype == np.float32
    assert np.abs(k[0] - golden["planted/keypoints"][0, 0]).max() <= 1e-3
    np.testing.assert_array_equal(s[0], golden["planted/vals"][0])
    ks, ss = codec.batch_decode(torch.from_numpy(hms).cuda())
    assert len(ks) == 12 and ks[3].shape == (1, 17, 2)
    assert np.abs(np.concatenate(ks) - golden["planted/keypoints"][:, 0]).max() <= 1e-3
    with pytest.raises(AssertionError):
        codec.decode(hms)  # (B, K, H, W) is not a single instance (the reference raises for B > 1 too)
    with pytest.raises(NotImplementedError):
        codec.encode(np.zeros((1, 17, 2)))


def test_backbone_and_head_forward(world):
    m, ref = world["model"], world["ref"]
    x = world["x"].cuda()
    feats = m.backbone(x)
    assert isinstance(feats, tuple) and feats[0].shape == (3, 384, 16, 12)
    with torch.no_grad():
        rf = ref.backbone(world["x"])
        rh = ref.head(rf)
    assert (feats[0].cpu() - rf[0]).abs().max() <= 1e-4
    out = m.head.forward(feats)
    assert out[0].shape == (3, 17, 64, 48) and all(o.shape == (3, 17, 1, 1) for o in out[1:])
    assert (out[0].cpu() - rh[0]).abs().max() <= 1e-4
    for a, b in zip(out[1:], rh[1:]):
        assert (a.cpu() - b).abs().max() <= 1e-4
    # tensor mode of the estimator = backbone + head.forward
    t = m.forward(x, None, mode="tensor")
    assert torch.equal(t[0], out[0])


@pytest.mark.parametrize("flip", [True, False])
def test_predict_paths_agree_with_oracle(world, flip):
    m, ref = world["model"], world["ref"]
    m.test_cfg["flip_test"] = flip
    rec = ref.predict(world["x"], flip_test=flip)
    samples = api.make_data_samples(3)
    # (1) fused: test_step on uint8 crops
    out = m.test_step(dict(inputs=[c for c in world["crops"]], data_samples=samples))
    # (2) generic: extract_feat + head.predict (what the reference's predict does module by module)
    x = world["x"].cuda()
    feats = [m.extract_feat(x), m.extract_feat(x.flip(-1))] if flip else m.extract_feat(x)
    preds = m.head.predict(feats, samples, test_cfg=m.test_cfg)
    for i in range(3):
        pi = out[i].pred_instances
        assert pi.keypoints.shape == (1, 17, 2) and pi.keypoints.dtype == np.float64
        assert np.abs(pi.keypoints[0] - rec[i, :, :2]).max() <= 1e-3
        assert np.abs(pi.keypoints_probs[0] - rec[i, :, 3]).max() <= 1e-4
        assert np.abs(pi.keypoints_visible[0] - rec[i, :, 4]).max() <= 1e-4
        assert np.abs(pi.keypoint_scores[0] - rec[i, :, 5]).max() <= 1e-4  # := oks
        assert np.abs(pi.keypoints_error[0] - rec[i, :, 6]).max() <= 1e-4
        assert np.abs(pi.keypoints_conf[0] - rec[i, :, 2]).max() <= 1e-4
        assert np.abs(preds[i].keypoints[0] - rec[i, :, :2]).max() <= 1e-3
        assert np.abs(preds[i].keypoints_probs[0] - rec[i, :, 3]).max() <= 1e-4
    m.test_cfg["flip_test"] = True


def test_output_heatmaps_and_weight_refresh(world):
    m = world["model"]
    m.test_cfg["output_heatmaps"] = True
    out = m.test_step(dict(inputs=world["crops"], data_samples=api.make_data_samples(3)))
    hm = out[0].pred_fields.heatmaps
    assert hm.shape == (17, 64, 48) and hm.is_cuda
    m.test_cfg.pop("output_heatmaps")
    before = out[0].pred_instances.keypoints_probs.copy()
    # weights are owned by the module: changing them must change the engine's output
    with torch.no_grad():
        m.head.probability_layers[12].bias.add_(1.0)
    after = m.test_step(dict(inputs=world["crops"], data_samples=api.make_data_samples(3)))[0].pred_instances.keypoints_probs
    assert (after > before).all()
    with torch.no_grad():
        m.head.probability_layers[12].bias.sub_(1.0)


@pytest.mark.parametrize("n", [1, 3, 5, 9])
def test_every_input_form_gives_the_same_samples(world, n):
    """test_step accepts what mmengine's pseudo_collate hands over (a list of pageable per-person tensors: gathered in
    quarters into pinned memory), a stacked pageable tensor, a pinned tensor and a device tensor; whatever the form and the
    batch size (quarters of unequal length, a single person), the samples must be the same bit for bit - keypoints in
    image space for per-person boxes, all seven fields, containers that are views of one batch array each - and a
    second call must not change what the first one returned."""
    m = world["model"]
    m.test_cfg["flip_test"] = True
    crops = synth.make_crops(n, seed=60 + n)

    def samples():
        s = api.make_data_samples(n)
        for i, d_ in enumerate(s):
            d_.set_metainfo(dict(input_center=np.array([300.0 + 11 * i, 200.0 - 3 * i], dtype=np.float32),
                                 input_scale=np.array([150.0 + i, 200.0 + 2 * i], dtype=np.float32)))
        return s

    forms = dict(list=[c.clone() for c in crops], stacked=crops.clone(), pinned=crops.clone().pin_memory(), device=crops.cuda())
    outs = {k: m.test_step(dict(inputs=v, data_samples=samples())) for k, v in forms.items()}
    names = ("keypoints", "keypoint_scores", "keypoints_conf", "keypoints_probs", "keypoints_visible", "keypoints_oks", "keypoints_error")
    first = {k: {f: np.array(getattr(o.pred_instances, f)) for f in names} for k, o in zip(range(n), outs["list"])}
    for form, out in outs.items():
        assert len(out) == n
        for i, o in enumerate(out):
            for f in names:
                assert np.array_equal(getattr(o.pred_instances, f), first[i][f]), f"{form}: person {i} field {f}"
            assert o.pred_instances.keypoints.shape == (1, 17, 2) and o.pred_instances.keypoints.dtype == np.float64
            assert np.array_equal(o.pred_instances.bboxes, o.gt_instances.bboxes)
    # image-space mapping (topdown.py:165-167) against the same crop run alone with the identity geometry of
    # make_data_samples (bbox = whole image: image space == input space up to the last bit)
    k_in = m.test_step(dict(inputs=crops[:1], data_samples=api.make_data_samples(1)))[0].pred_instances.keypoints
    meta = outs["list"][0].metainfo
    want = k_in / meta["input_size"] * meta["input_scale"] + meta["input_center"] - 0.5 * meta["input_scale"]
    assert np.abs(outs["list"][0].pred_instances.keypoints - want).max() <= 1e-9
    # results of an earlier call are not views of a reused staging buffer
    m.test_step(dict(inputs=synth.make_crops(n, seed=999), data_samples=samples()))
    for i, o in enumerate(outs["list"]):
        for f in names:
            assert np.array_equal(getattr(o.pred_instances, f), first[i][f])
