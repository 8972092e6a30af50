"""pp_decode_udp (fused flip-TTA merge + DARK-UDP decode) against the oracle, which is pinned bit-for-bit to the genuine
reference code (tests/test_oracle_udp.py).  Tolerance: 1e-3 input pixels on keypoints (the blur runs in fp32 with a
different summation order than OpenCV's SIMD filter), scores bit-equal."""
import numpy as np
import pytest
import torch

from oracle import udp_oracle as u

pytestmark = pytest.mark.gpu
FI = u.COCO_FLIP_INDICES


def _to_input(rec):  # udp_heatmap.py:194-195 in double
    return rec[..., :2].astype(np.float64) / [47, 63] * (192, 256)


@pytest.mark.parametrize("name,hms", [("gauss", u.gaussian_heatmaps(6, seed=0)), ("noisy", u.gaussian_heatmaps(3, seed=1, noise=0.05)),
                                      ("special", u.special_heatmaps())])
def test_udp_decode_matches_oracle(name, hms):
    from probpose_code_b200 import ops
    rec = ops.decode_udp(torch.from_numpy(hms).cuda()).cpu().numpy()
    kp, sc = u.decode_instances(hms)
    kp, sc = np.concatenate(kp), np.concatenate(sc)
    np.testing.assert_array_equal(rec[..., 2], sc)
    # maps around the clip at 1e-3 / with a flat 2-pixel top have a (near-)singular Hessian: excluded from the px bound
    ok = np.ones(kp.shape[:2], bool)
    if name == "special":
        ok[0, [10, 11, 15]] = False
    d = np.abs(_to_input(rec) - kp).max(-1)
    assert d[ok].max() <= 1e-3, f"max keypoint deviation {d[ok].max()} px"


def test_udp_decode_flip_merge_and_heatmap_output():
    from probpose_code_b200 import ops
    a = u.gaussian_heatmaps(5, seed=2)
    inv = np.argsort(FI)
    bflip = np.ascontiguousarray(u.gaussian_heatmaps(5, seed=2, noise=0.02)[:, inv][..., ::-1])
    rec, merged = ops.decode_udp(torch.from_numpy(a).cuda(), torch.from_numpy(bflip).cuda(), FI, return_heatmaps=True)
    ref_merged = u.merge_flip(a, bflip)
    np.testing.assert_array_equal(merged.cpu().numpy(), ref_merged)
    kp, sc = u.decode_instances(ref_merged)
    np.testing.assert_array_equal(rec.cpu().numpy()[..., 2], np.concatenate(sc))
    assert np.abs(_to_input(rec.cpu().numpy()) - np.concatenate(kp)).max() <= 1e-3


def test_udp_decode_no_response_and_errors():
    from probpose_code_b200 import ops
    hm = torch.zeros(1, 17, 64, 48, device="cuda")
    hm[0, 1] = -0.5
    hm[0, 2, 10, 7] = 0.25
    rec = ops.decode_udp(hm).cpu().numpy()
    assert rec[0, 0].tolist() == [-1.0, -1.0, 0.0] and rec[0, 1].tolist() == [-1.0, -1.0, -0.5]
    assert rec[0, 2, 2] == 0.25 and abs(rec[0, 2, 0] - 7) < 0.5 and abs(rec[0, 2, 1] - 10) < 0.5
    with pytest.raises(ValueError):
        ops.decode_udp(hm, hm, None)
    with pytest.raises(Exception):
        ops.decode_udp(hm, blur_kernel_size=10)
    assert ops.decode_udp(hm[:0]).shape == (0, 17, 3)
