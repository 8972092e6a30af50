"""ViT backbone + HeatmapHead + UDPHeatmap (the ViTPose td-hm configs) through the MMPose-shaped
plugin API, against the fp32 torch restatement of the model and the DARK-UDP decode oracle (pinned to the reference)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _state_dict(seed=3):
    from probpose_code_b200 import synth
    sd = synth.make_state_dict(seed=seed)
    sd = {k: v for k, v in sd.items() if k.startswith("backbone.") or k.startswith("head.deconv_layers") or k.startswith("head.final_layer")}
    sd["head.final_layer.bias"] = sd["head.final_layer.bias"] + 1.0  # positive maps, like trained heatmap heads
    return sd


@pytest.fixture(scope="module")
def setup():
    import probpose_code_b200.mmpose_api as api
    from oracle import model_oracle
    from probpose_code_b200 import synth
    sd = _state_dict()
    model = api.MODELS.build(api.vitpose_cfg("small", precision="fp16x3"))
    model.load_state_dict(sd)
    model.to("cuda:0")
    ref = model_oracle.ViTPoseRef().eval()
    ref.load_state_dict(sd)
    crops = synth.make_crops(3, seed=11)
    return api, model, ref, crops


def test_state_dict_layout(setup):
    api, model, ref, _ = setup
    keys = {k for k in model.state_dict() if k.startswith("head.")}
    assert keys == {f"head.deconv_layers.{i}.{n}" for i, names in ((0, ["weight"]), (3, ["weight"]),
                    (1, ["weight", "bias", "running_mean", "running_var", "num_batches_tracked"]),
                    (4, ["weight", "bias", "running_mean", "running_var", "num_batches_tracked"])) for n in names} | {
        "head.final_layer.weight", "head.final_layer.bias"}
    assert isinstance(model.head, api.HeatmapHead) and isinstance(model.head.decoder, api.UDPHeatmap)


@pytest.mark.parametrize("flip", [True, False])
def test_fused_predict_matches_oracle(setup, flip):
    from oracle import udp_oracle
    api, model, ref, crops = setup
    model.test_cfg = dict(flip_test=flip, flip_mode="heatmap", shift_heatmap=False, output_heatmaps=True)
    out = model.test_step(dict(inputs=[c for c in crops], data_samples=api.make_data_samples(len(crops))))
    hm = torch.stack([o.pred_fields.heatmaps for o in out]).cpu().numpy()
    hm_ref = ref.heatmaps(ref.preprocess(crops), flip_test=flip).numpy()
    assert np.abs(hm - hm_ref).max() <= 2e-5 * np.abs(hm_ref).max(), "heatmaps deviate from the fp32 torch model"
    kpts = np.concatenate([o.pred_instances.keypoints for o in out])
    scores = np.concatenate([o.pred_instances.keypoint_scores for o in out])
    assert kpts.dtype == np.float64 and kpts.shape == (3, 17, 2) and scores.shape == (3, 17)
    # decode parity on identical heatmap bits (bbox = whole image, so image space == input space)
    kp_o, sc_o = udp_oracle.decode_instances(hm)
    np.testing.assert_array_equal(scores, np.concatenate(sc_o))
    # identical heatmap bits in, blur and rescale bit-identical to cv2 / NumPy: the one rounding left between kernel and
    # reference is the float32 log (the kernel's is correctly rounded, numpy's SIMD log is not for ~2 % of its inputs),
    # and a random-init head's flat maps have near-singular DARK Hessians that amplify that last bit.  Most keypoints
    # are identical to ~1e-14 px; 1e-3 px wherever the Hessian is conditioned at all (|eig| >= 5e-3; a trained sigma = 2 peak
    # has ~1/8 - tests/test_decode_udp_gpu.py asserts 1e-3 px on EVERY trained-like, special and no-response map).
    eig = np.stack([udp_oracle.hessian_min_eig(h) for h in hm])
    well = (scores > 0) & (eig >= 5e-3)
    d = np.abs(kpts - np.concatenate(kp_o)).max(-1)
    assert well.sum() >= 3 and d[well].max() <= 1e-3, f"{d[well].max()} px on conditioned maps"
    assert np.mean(d < 1e-6) >= 0.4 and np.median(d) <= 1e-4, "about half of these flat maps decode identically to the last bit"
    # and end to end against the oracle pipeline
    rec = ref.predict(ref.preprocess(crops), flip_test=flip)
    assert np.abs(scores - rec[..., 2]).max() <= 2e-5 * np.abs(rec[..., 2]).max()
    same_peak = np.abs(kpts - rec[..., :2]).max(-1) < 4.0  # flat maps may move their arg max by rounding
    assert same_peak.mean() >= 0.8


def test_unfused_head_and_codec_api(setup):
    from oracle import udp_oracle
    api, model, ref, crops = setup
    x = ref.preprocess(crops).cuda()
    feats = model.extract_feat(x)
    hm = model.head.forward(feats)
    assert tuple(hm.shape) == (3, 17, 64, 48)
    preds = model.head.predict(feats, api.make_data_samples(3), test_cfg=dict(flip_test=False))
    k_np, s_np = model.head.decoder.decode(hm[0].cpu().numpy())
    assert k_np.shape == (1, 17, 2) and k_np.dtype == np.float64 and s_np.shape == (1, 17)
    np.testing.assert_array_equal(k_np, preds[0].keypoints)
    kb, sb = model.head.decoder.batch_decode(hm)
    np.testing.assert_array_equal(kb[0], k_np)
    assert model.head.decoder.support_batch_decoding
    # shift_heatmap=True (the default of tta.flip_heatmaps) leaves the fused kernel for the tensor path: same result as
    # merging by hand and decoding the merged maps
    from probpose_code_b200.mmpose_api.utils import flip_heatmaps
    feats_f = model.extract_feat(x.flip(-1))
    cfg = dict(flip_test=True, flip_mode="heatmap", shift_heatmap=True)
    got = model.head.predict([feats, feats_f], api.make_data_samples(3), test_cfg=cfg)
    merged = (hm + flip_heatmaps(model.head.forward(feats_f), flip_indices=api.COCO_FLIP_INDICES, shift_heatmap=True)) * 0.5
    kb2, _ = model.head.decoder.batch_decode(merged)
    np.testing.assert_array_equal(got[1].keypoints, kb2[1])
    with pytest.raises(ValueError):
        api.MODELS.build(dict(type="HeatmapHead", in_channels=384, out_channels=17, deconv_out_channels=(256, 256), deconv_kernel_sizes=(4,)))
    with pytest.raises(ValueError):
        api.KEYPOINT_CODECS.build(dict(type="UDPHeatmap", input_size=(192, 256), heatmap_size=(48, 64), heatmap_type="nope"))
