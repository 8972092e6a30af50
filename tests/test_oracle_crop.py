"""The crop front-end oracle (GetBBoxCenterScale + TopdownAffine(use_udp) + cv2.warpAffine)
against the golden vectors captured from the genuine reference functions and cv2 itself."""
import hashlib
import os

import numpy as np
import pytest

from oracle import crop_oracle as co

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "crop_kat.npz")
FRAMES = [(480, 640), (333, 517), (1080, 1920)]


@pytest.fixture(scope="module")
def kat():
    return np.load(GOLDEN)


@pytest.mark.parametrize("fi", [0, 1, 2])
def test_geometry_matches_reference_functions(kat, fi):
    """centre / scale / UDP matrix == bbox_xyxy2cs + _fix_aspect_ratio + get_udp_warp_matrix, bit for bit."""
    for i, bbox in enumerate(kat[f"f{fi}/boxes"]):
        c, s, m = co.topdown_geometry(bbox)
        np.testing.assert_array_equal(c, kat[f"f{fi}/centers"][i])
        np.testing.assert_array_equal(s, kat[f"f{fi}/scales"][i])
        np.testing.assert_array_equal(m, kat[f"f{fi}/mats"][i])
        assert m.dtype == np.float32 and abs(s[0] / s[1] - 0.75) < 1e-6  # fixed aspect ratio w / h


@pytest.mark.parametrize("fi", [0, 1, 2])
def test_warp_matches_cv2_bit_for_bit(kat, fi):
    fh, fw = FRAMES[fi]
    frame = co.synthetic_frame(10 + fi, fh, fw)
    assert hashlib.sha256(frame.tobytes()).hexdigest()[:16] == str(kat[f"f{fi}/frame_sha"])
    crops, centers, scales, mats = co.topdown_crops(frame, kat[f"f{fi}/boxes"])
    assert crops.shape == (12, 3, 256, 192) and crops.dtype == np.uint8
    hwc = crops.transpose(0, 2, 3, 1)
    np.testing.assert_array_equal(hwc[:2], kat[f"f{fi}/crops_head"])
    sha = [hashlib.sha256(np.ascontiguousarray(c).tobytes()).hexdigest()[:16] for c in hwc]
    assert sha == list(kat[f"f{fi}/crop_sha"])


def test_whole_image_box_is_the_demo_default(kat):
    """inference_topdown without boxes uses [0, 0, w, h] (apis/inference.py:161-168)."""
    c, s, m = co.topdown_geometry(np.array([0, 0, 640, 480], np.float32))
    np.testing.assert_allclose(c, [320, 240])
    np.testing.assert_allclose(s, [800, 800 / 0.75], rtol=1e-6)  # 640 * 1.25 wide, height from the 3:4 aspect
