"""pp_decode_udp (fused flip-TTA merge + DARK-UDP decode) against the oracle, which is pinned bit-for-bit to the genuine
reference code (tests/test_oracle_udp.py).  The blur reproduces cv2.GaussianBlur's float arithmetic bit for bit, so the
only rounding left between kernel and reference is logf (CUDA's vs numpy's float32 log, <= 1 ulp each).  Tolerance:
1e-3 input pixels on EVERY map - corners, ties, clipped maps, and maps without a response (maximum <= 0), where the
reference's index arithmetic wraps into the previous keypoint's map; scores bit-equal."""
import numpy as np
import pytest
import torch

from oracle import udp_oracle as u

pytestmark = pytest.mark.gpu
FI = u.COCO_FLIP_INDICES


def _to_input(rec):  # udp_heatmap.py:194-195 in double
    return rec[..., :2].astype(np.float64) / [47, 63] * (192, 256)


@pytest.mark.parametrize("name,hms", [("gauss", u.gaussian_heatmaps(6, seed=0)), ("noisy", u.gaussian_heatmaps(3, seed=1, noise=0.05)),
                                      ("special", u.special_heatmaps()), ("noresp", u.no_response_heatmaps())])
def test_udp_decode_matches_oracle(name, hms):
    from probpose_code_b200 import ops
    rec = ops.decode_udp(torch.from_numpy(hms).cuda()).cpu().numpy()
    kp, sc = u.decode_instances(hms)
    kp, sc = np.concatenate(kp), np.concatenate(sc)
    np.testing.assert_array_equal(rec[..., 2], sc)
    d = np.abs(_to_input(rec) - kp).max(-1)
    assert d.max() <= 1e-3, f"max keypoint deviation {d.max()} px at {np.unravel_index(d.argmax(), d.shape)}"


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
    kp, sc = u.decode_instances(hm.cpu().numpy())
    assert rec[0, 0, 2] == 0.0 and rec[0, 1, 2] == -0.5 and rec[0, 2, 2] == 0.25
    assert np.abs(_to_input(rec) - np.concatenate(kp)).max() <= 1e-3  # (-1, -1) minus the wrapped refinement step, as the reference
    assert abs(rec[0, 2, 0] - 7) < 0.5 and abs(rec[0, 2, 1] - 10) < 0.5
    with pytest.raises(ValueError):
        ops.decode_udp(hm, hm, None)
    with pytest.raises(Exception):
        ops.decode_udp(hm, blur_kernel_size=10)
    assert ops.decode_udp(hm[:0]).shape == (0, 17, 3)
