"""The UDPHeatmap (DARK-UDP) decode oracle against golden vectors captured from the genuine reference code
(tests/golden/udp_kat.npz, made by oracle/gen_golden_udp.py).  CPU only."""
import os

import numpy as np
import pytest

from oracle import cases
from oracle import udp_oracle as u

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "udp_kat.npz")


@pytest.fixture(scope="module")
def udp_golden():
    return np.load(GOLDEN)


def _families():
    a, b = u.gaussian_heatmaps(4, seed=2), u.gaussian_heatmaps(4, seed=2, noise=0.02)
    inv = np.argsort(u.COCO_FLIP_INDICES)
    return {
        "gauss": u.gaussian_heatmaps(8, seed=0),
        "gauss_noisy": u.gaussian_heatmaps(4, seed=1, noise=0.05),
        "special": u.special_heatmaps(),
        "noresp": u.no_response_heatmaps(),
        "merged": u.merge_flip(a, np.ascontiguousarray(b[:, inv][..., ::-1])),
    }


@pytest.mark.parametrize("name", ["gauss", "gauss_noisy", "special", "noresp", "merged"])
def test_udp_oracle_matches_reference_bit_for_bit(udp_golden, name):
    hms = _families()[name]
    assert cases.checksum(hms) == str(udp_golden[f"{name}/input_sha"]), "seeded inputs drifted from the golden run"
    kpts, scores = u.decode_instances(hms)
    kpts, scores = np.stack(kpts), np.stack(scores)
    assert kpts.dtype == np.float64 and scores.dtype == np.float32 and kpts.shape == (len(hms), 1, 17, 2)
    np.testing.assert_array_equal(kpts, udp_golden[f"{name}/keypoints"])
    np.testing.assert_array_equal(scores, udp_golden[f"{name}/scores"])


def test_blur_is_a_zero_padded_separable_gaussian(udp_golden):
    """What the CUDA kernel relies on: the reference's padded cv2.GaussianBlur == zero-padded separable 11-tap filter
    with OpenCV's float taps, up to float rounding."""
    hm = u.gaussian_heatmaps(8, seed=0)[0]
    g = u.gaussian_kernel_1d(11).astype(np.float64)
    pad = np.zeros((17, 64 + 10, 48 + 10))
    pad[:, 5:-5, 5:-5] = hm
    rows = sum(g[j] * pad[:, :, j:j + 48] for j in range(11))
    out = sum(g[j] * rows[:, j:j + 64, :] for j in range(11))
    out *= (hm.reshape(17, -1).max(1) / (out.reshape(17, -1).max(1) + 1e-12))[:, None, None]
    np.testing.assert_allclose(out, udp_golden["gauss/blurred0"], rtol=2e-6, atol=1e-8)


def test_maximum_rule():
    hm = np.zeros((3, 4, 5), np.float32)
    hm[1, 2, 3] = hm[1, 3, 1] = 0.7  # tie: the first flat index wins
    hm[2] = -1.0
    locs, vals = u.heatmap_maximum(hm)
    assert locs.tolist() == [[-1, -1], [3, 2], [-1, -1]] and vals.tolist() == [0.0, np.float32(0.7), -1.0]


@pytest.mark.parametrize("ksize", [11, 17])
def test_exact_blur_emulation_equals_cv2(ksize):
    """oracle.udp_oracle.gaussian_blur_exact (= the arithmetic of csrc/decode_udp.cu) against cv2.GaussianBlur on the
    reference's zero-padded image, bit for bit, on every family incl. negative maps and large magnitudes."""
    import cv2
    b = (ksize - 1) // 2
    rng = np.random.default_rng(3)
    maps = [u.gaussian_heatmaps(1, seed=7)[0, 2], u.special_heatmaps()[0, 12], u.special_heatmaps()[0, 14],
            u.no_response_heatmaps()[0, 4], rng.normal(0, 1, (64, 48)).astype(np.float32),
            (rng.random((64, 48)) * 1e-4).astype(np.float32)]
    for m in maps:
        pad = np.zeros((64 + 2 * b, 48 + 2 * b), np.float32)
        pad[b:-b, b:-b] = m
        want = cv2.GaussianBlur(pad, (ksize, ksize), 0)[b:-b, b:-b]
        np.testing.assert_array_equal(u.gaussian_blur_exact(m, ksize), want)
