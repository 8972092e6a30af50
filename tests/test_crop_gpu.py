"""GPU crop front-end (pp_crop_warp) and the inference_topdown caller contract:
bit-exact crops against the cv2-pinned oracle and the golden vectors; frame + boxes -> image-space
keypoints against the oracle pipeline within the north-star tolerance."""
import hashlib
import os

import numpy as np
import pytest
import torch

from oracle import crop_oracle as co
from oracle import model_oracle
from probpose_code_b200 import synth

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "crop_kat.npz")
FRAMES = [(480, 640), (333, 517), (1080, 1920)]


@pytest.mark.parametrize("fi", [0, 1, 2])
def test_crops_are_bit_exact(fi):
    from probpose_code_b200 import ops
    kat = np.load(GOLDEN)
    fh, fw = FRAMES[fi]
    frame = co.synthetic_frame(10 + fi, fh, fw)
    mats = kat[f"f{fi}/mats"]
    crops = ops.crop_warp(torch.from_numpy(frame).cuda(), torch.from_numpy(mats).cuda()).cpu().numpy()
    assert crops.shape == (12, 3, 256, 192)
    hwc = crops.transpose(0, 2, 3, 1)
    np.testing.assert_array_equal(hwc[:2], kat[f"f{fi}/crops_head"])  # cv2.warpAffine's own output
    assert [hashlib.sha256(np.ascontiguousarray(c).tobytes()).hexdigest()[:16] for c in hwc] == list(kat[f"f{fi}/crop_sha"])


def test_random_matrices_and_pitched_frames():
    """General affine matrices (rotation / shear) and a frame that is a view with a row pitch."""
    from probpose_code_b200 import ops
    rng = np.random.default_rng(3)
    big = co.synthetic_frame(5, 300, 500)
    frame = big[:, 20:420]  # non-contiguous rows: pitch 1500 bytes, width 400
    mats = np.stack([np.array([[rng.uniform(0.3, 2.5), rng.uniform(-0.6, 0.6), rng.uniform(-150, 150)],
                               [rng.uniform(-0.6, 0.6), rng.uniform(0.3, 2.5), rng.uniform(-150, 150)]], np.float32) for _ in range(9)])
    ft = torch.from_numpy(big).cuda()[:, 20:420]
    assert not ft.is_contiguous()
    with pytest.raises(ValueError):
        ops.crop_warp(ft, torch.from_numpy(mats).cuda())  # the wrapper wants one contiguous (H, W, 3) frame
    got = ops.crop_warp(ft.contiguous(), torch.from_numpy(mats).cuda(), out_hw=(64, 48)).cpu().numpy()
    for i, m in enumerate(mats):
        ref = co.warp_affine_u8(np.ascontiguousarray(frame), m, (48, 64)).transpose(2, 0, 1)
        np.testing.assert_array_equal(got[i], ref)
    assert ops.crop_warp(ft.contiguous(), torch.zeros((0, 2, 3), device="cuda")).shape == (0, 3, 256, 192)


def test_inference_topdown_from_frame_and_boxes():
    import probpose_code_b200.mmpose_api as api
    sd = synth.make_state_dict(seed=0)
    ref = model_oracle.ProbPoseRef().eval()
    ref.load_state_dict(sd)
    frame = co.synthetic_frame(31, 480, 640)
    boxes = co.synthetic_boxes(32, 5, 480, 640)
    # oracle: reference pipeline on the CPU
    crops, centers, scales, _ = co.topdown_crops(frame, boxes)
    rec = ref.predict(ref.preprocess(torch.from_numpy(crops)), flip_test=True)
    model = api.MODELS.build(api.probpose_small_cfg(precision="fp16x3", flip_test=True))
    model.load_state_dict(sd)
    model.cuda()
    out = api.inference_topdown(model, frame, boxes)
    assert len(out) == 5
    for i, ds in enumerate(out):
        kp_ref = rec[i, :, :2] / np.array([192, 256]) * scales[i] + centers[i] - 0.5 * scales[i]  # topdown.py:165-167
        tol = 1e-3 * float(np.max(scales[i] / np.array([192, 256])))  # 1e-3 INPUT px, mapped to image space
        assert np.abs(ds.pred_instances.keypoints[0] - kp_ref).max() <= tol
        assert np.abs(ds.pred_instances.keypoints_probs[0] - rec[i, :, 3]).max() <= 1e-4
        np.testing.assert_array_equal(ds.pred_instances.bboxes, boxes[i][None])
    # no boxes -> the whole image, xywh boxes -> converted
    whole = api.inference_topdown(model, frame)
    assert len(whole) == 1 and whole[0].pred_instances.keypoints.shape == (1, 17, 2)
    xywh = boxes.copy()
    xywh[:, 2:] -= xywh[:, :2]
    out2 = api.inference_topdown(model, frame, xywh, bbox_format="xywh")
    np.testing.assert_allclose(out2[2].pred_instances.keypoints, out[2].pred_instances.keypoints, atol=5e-3)
