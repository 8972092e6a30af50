"""Host logic of the MMPose-shaped plugin layer that needs no GPU: registry resolution, constructor
contracts and error behaviour mirrored from the reference, checkpoint key layout, result packing."""
import numpy as np
import pytest
import torch

import probpose_code_b200.mmpose_api as api
from probpose_code_b200 import synth


def test_registry_resolves_the_config_type_strings():
    assert api.MODELS.get("mmpretrain.VisionTransformer") is api.VisionTransformer
    assert api.MODELS.get("VisionTransformer") is api.VisionTransformer
    assert api.MODELS.get("ProbMapHead") is api.ProbMapHead
    assert api.MODELS.get("TopdownPoseEstimator") is api.TopdownPoseEstimator
    assert api.KEYPOINT_CODECS.get("ProbMap") is api.ProbMap
    with pytest.raises(KeyError):
        api.MODELS.build(dict(type="NoSuchThing"))
    with pytest.raises(TypeError):
        api.MODELS.build(dict(foo=1))


def test_state_dict_layout_is_the_reference_checkpoint_layout():
    model = api.MODELS.build(api.probpose_small_cfg())
    sd = synth.make_state_dict(seed=0)
    assert set(model.state_dict()) == set(sd)
    res = model.load_state_dict(sd, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    # witnesses from the reference: deconv weight is (Cin, Cout, 4, 4); head.test_cfg is a copy of the model's
    assert model.head.deconv_layers[0].weight.shape == (384, 256, 4, 4)
    assert model.head.deconv_layers[3].weight.shape == (256, 256, 4, 4)
    assert model.head.test_cfg == model.test_cfg and model.head.test_cfg is not model.test_cfg
    assert model.head.decoder.support_batch_decoding
    names = model.head.engine_tensors()
    assert "head.probability_layers.9.running_var" in names and not any(n.endswith("num_batches_tracked") for n in names)


def test_constructor_errors_match_the_reference():
    with pytest.raises(ValueError, match="same length"):  # probmap_head.py:211-217
        api.ProbMapHead(in_channels=384, out_channels=17, deconv_out_channels=(256, 256), deconv_kernel_sizes=(4,),
                        normalize=1.0, decoder=dict(type="ProbMap", input_size=(192, 256), heatmap_size=(48, 64)))
    with pytest.raises(ValueError, match="heatmap_type"):  # probmap.py:91-96
        api.ProbMap(input_size=(192, 256), heatmap_size=(48, 64), heatmap_type="nope")
    with pytest.raises(NotImplementedError):  # outside the ProbPose configuration: loud, not silent
        api.ProbMapHead(in_channels=384, out_channels=17, deconv_out_channels=(256, 256, 256), normalize=1.0,
                        decoder=dict(type="ProbMap", input_size=(192, 256), heatmap_size=(48, 64)))
    with pytest.raises(ValueError):
        api.VisionTransformer(arch="nope", img_size=(256, 192), out_type="featmap", with_cls_token=False)
    with pytest.raises(NotImplementedError):
        api.VisionTransformer(arch="base", img_size=(256, 192))  # cls-token classifier configuration
    head = api.ProbMapHead(in_channels=384, out_channels=17, deconv_out_channels=(256, 256), deconv_kernel_sizes=(4, 4),
                           normalize=1.0, decoder=None)
    with pytest.raises(RuntimeError, match="decoder has not been set"):  # base_head.py:50-55
        head.decode(torch.zeros(1, 17, 64, 48))
    with pytest.raises(NotImplementedError):
        head.loss(None, None)
    # other flip modes / codecs are not an error: they leave the fused kernels for the tensor path (tta.py:9-67)
    assert head.fused_test_cfg(dict(flip_test=True, flip_mode="heatmap", shift_heatmap=False))
    assert not head.fused_test_cfg(dict(flip_test=True, flip_mode="udp_combined"))
    assert not head.fused_test_cfg(dict(flip_test=True, shift_heatmap=True))
    assert not head.fused_decoder()
    default = api.ProbMapHead(in_channels=384, out_channels=17, deconv_out_channels=(256, 256), deconv_kernel_sizes=(4, 4),
                              normalize=1.0)  # constructor default decoder = UDPHeatmap, as in the reference
    assert isinstance(default.decoder, api.UDPHeatmap) and not default.fused_decoder()


def test_flip_heatmaps_matches_the_reference_semantics():
    """utils.flip_heatmaps restates mmpose/models/utils/tta.py:9-67 (all modes, shift)."""
    from probpose_code_b200.mmpose_api.utils import flip_heatmaps
    g = torch.Generator().manual_seed(0)
    fi = [0, 2, 1]
    hm = torch.rand(2, 3, 4, 5, generator=g)
    out = flip_heatmaps(hm, flip_indices=fi, flip_mode="heatmap", shift_heatmap=False)
    assert torch.equal(out, hm.flip(-1)[:, fi])
    sh = flip_heatmaps(hm, flip_indices=fi, flip_mode="heatmap", shift_heatmap=True)
    assert torch.equal(sh[..., 1:], out[..., :-1]) and torch.equal(sh[..., 0], out[..., 0])
    comb = torch.rand(2, 9, 4, 5, generator=g)
    out = flip_heatmaps(comb, flip_indices=fi, flip_mode="udp_combined", shift_heatmap=False)
    ref = comb.view(2, 3, 3, 4, 5).flip(-1)[:, fi].clone()
    ref[:, :, 1] = -ref[:, :, 1]
    assert torch.equal(out, ref.view(2, 9, 4, 5))
    off = torch.rand(2, 6, 4, 5, generator=g)
    out = flip_heatmaps(off, flip_indices=fi, flip_mode="offset", shift_heatmap=False)
    ref = off.view(2, 3, 2, 4, 5).flip(-1)[:, fi].clone()
    ref[:, :, 0] = -ref[:, :, 0]
    assert torch.equal(out, ref.view(2, 6, 4, 5))
    with pytest.raises(ValueError, match="Invalid flip_mode"):
        flip_heatmaps(hm, flip_mode="nope")


def test_no_cpu_fallback():
    """Without a CUDA device the product path must fail loudly."""
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from probpose_code_b200 import _lib
    from probpose_code_b200.engine import Engine

    with pytest.raises(_lib.PPError):
        Engine(precision="fp16x3", max_batch=1)
    model = api.MODELS.build(api.probpose_small_cfg())
    with pytest.raises(Exception):
        model.test_step(dict(inputs=[torch.zeros(3, 256, 192, dtype=torch.uint8)], data_samples=api.make_data_samples(1)))
    from probpose_code_b200 import ops
    with pytest.raises(_lib.PPError):
        ops.decode(torch.zeros(1, 17, 64, 48), input_is_logits=False)


def test_pack_records_and_image_space_mapping():
    model = api.MODELS.build(api.probpose_small_cfg())
    rec = torch.zeros(2, 17, 7)
    rec[..., 0], rec[..., 1] = 47.0, 63.0  # bottom-right heatmap pixel
    rec[..., 2], rec[..., 3], rec[..., 4], rec[..., 5], rec[..., 6] = 0.9, 0.8, 0.7, 0.6, 0.5
    preds = model.head.pack_records(rec)
    p = preds[0]
    assert p.keypoints.shape == (1, 17, 2) and p.keypoints.dtype == np.float64
    np.testing.assert_allclose(p.keypoints[0, 0], [192.0, 256.0])  # probmap.py:218: / (W-1, H-1) * input_size
    assert p.keypoint_scores.shape == (1, 17) and p.keypoint_scores.dtype == np.float32
    np.testing.assert_allclose(p.keypoint_scores, 0.6)  # replaced by oks (freeze_oks=False), probmap_head.py:796-798
    np.testing.assert_allclose(p.keypoints_conf, 0.9)
    np.testing.assert_allclose(p.keypoints_probs, 0.8)
    np.testing.assert_allclose(p.keypoints_visible, 0.7)
    np.testing.assert_allclose(p.keypoints_error, 0.5)
    samples = api.make_data_samples(2)
    samples[1].set_metainfo(dict(input_center=np.array([500.0, 400.0]), input_scale=np.array([96.0, 128.0])))
    out = model.add_pred_to_datasample(preds, None, samples)
    np.testing.assert_allclose(out[0].pred_instances.keypoints[0, 0], [192.0, 256.0])
    # topdown.py:165-167: kpts / input_size * input_scale + input_center - 0.5 * input_scale
    np.testing.assert_allclose(out[1].pred_instances.keypoints[0, 0], [192 / 192 * 96 + 500 - 48, 256 / 256 * 128 + 400 - 64])
    assert out[1].pred_instances.bboxes.shape == (1, 4)


def test_two_phase_record_packing_matches_the_reference_arithmetic():
    """The fused path builds the per-person containers before the records arrive (alloc_records) and fills the batch
    arrays in place afterwards (fill_records): the values must be those of the reference's per-person arithmetic -
    probmap.py:218 (`keypoints / [W - 1, H - 1] * input_size`, float64 through numpy's promotion) and
    topdown.py:165-167 (`keypoints / input_size * input_scale + input_center - 0.5 * input_scale`) - bit for bit."""
    from probpose_code_b200.mmpose_api.head import _record_xy
    model = api.MODELS.build(api.probpose_small_cfg())
    rng = np.random.default_rng(3)
    b = 5
    rec = rng.random((b, 17, 7), dtype=np.float32) * np.float32(47.0)
    assert np.array_equal(_record_xy(rec), rec[:, :, :2])
    codec = model.head.decoder
    want_in = rec[:, :, :2] / [47, 63] * codec.input_size
    assert np.array_equal(codec.keypoints_from_locs(rec[:, :, :2]), want_in)
    assert np.array_equal(codec.keypoints_from_locs(rec[0, :, :2]), want_in[0])
    samples = api.make_data_samples(b)
    for i, s in enumerate(samples):
        s.set_metainfo(dict(input_center=np.array([100.0 + 7 * i, 50.0 + i], dtype=np.float32),
                            input_scale=np.array([90.0 + i, 120.0 + 3 * i], dtype=np.float32)))
    arrays, preds = model.head.alloc_records(b)
    geo = [(s.metainfo["input_size"], s.metainfo["input_scale"], s.metainfo["input_center"]) for s in samples]

    def to_image(k):
        out = np.empty_like(k)
        for i, (size, scale, center) in enumerate(geo):
            out[i] = k[i] / size * scale + center - 0.5 * scale
        return out

    model.head.fill_records(arrays, rec, to_image=to_image)
    for i, p in enumerate(preds):
        size, scale, center = geo[i]
        assert np.array_equal(p.keypoints[0], want_in[i] / size * scale + center - 0.5 * scale)
        assert p.keypoints.dtype == np.float64 and p.keypoints.shape == (1, 17, 2)
        for j, name in enumerate(("keypoints_conf", "keypoints_probs", "keypoints_visible", "keypoints_oks", "keypoints_error")):
            assert np.array_equal(getattr(p, name), rec[i:i + 1, :, 2 + j]) and getattr(p, name).dtype == np.float32
        assert np.array_equal(p.keypoint_scores, rec[i:i + 1, :, 5])  # := oks (probmap_head.py:796-798)


def test_frame_geometry_matches_the_reference_functions_bit_for_bit():
    """mmpose_api.inference computes centres, extents and UDP matrices for all boxes of a frame in one pass; they must be
    the float32 values the genuine bbox_xyxy2cs / _fix_aspect_ratio / get_udp_warp_matrix produced per person
    (tests/golden/crop_kat.npz, oracle/gen_golden_crops.py), singly and batched."""
    import os
    from probpose_code_b200.mmpose_api.inference import TopdownAffine, bbox_xywh2xyxy, get_udp_warp_matrix
    kat = np.load(os.path.join(os.path.dirname(__file__), "golden", "crop_kat.npz"))
    aff = TopdownAffine(input_size=(192, 256), use_udp=True, input_padding=1.25)
    for fi in range(3):
        boxes = kat[f"f{fi}/boxes"]
        c, s, m = aff.batch_geometry(boxes)
        assert m.dtype == np.float32
        np.testing.assert_array_equal(c, kat[f"f{fi}/centers"])
        np.testing.assert_array_equal(s, kat[f"f{fi}/scales"])
        np.testing.assert_array_equal(m, kat[f"f{fi}/mats"])
        c1, s1, m1 = aff.geometry(boxes[3][None])
        np.testing.assert_array_equal(m1, kat[f"f{fi}/mats"][3])
        np.testing.assert_array_equal(get_udp_warp_matrix(c1, s1, 0.0, (192, 256)), m1)
    xywh = np.array([[10.0, 20.0, 30.0, 40.0, 0.9]], dtype=np.float32)
    np.testing.assert_array_equal(bbox_xywh2xyxy(xywh), np.array([[10.0, 20.0, 40.0, 60.0, 0.9]], dtype=np.float32))
    assert xywh[0, 2] == 30.0  # the input is not modified
    rot = get_udp_warp_matrix(np.array([50.0, 60.0], np.float32), np.array([90.0, 120.0], np.float32), 30.0, (192, 256))
    assert rot.shape == (2, 3) and abs(rot[0, 0] - np.cos(np.pi / 6) * 191 / 90) < 1e-5
