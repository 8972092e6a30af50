"""The decode oracle against the golden vectors captured from the genuine reference code
(tests/golden/decode_kat.npz, made by oracle/gen_golden.py).  CPU only."""
import numpy as np
import pytest

from oracle import cases
from oracle import decode_oracle as d

FAMILIES = {
    "planted": lambda: d.heatmaps_from_logits(cases.planted_peak_logits(12, seed=0)),
    "noise_1e-3": lambda: d.heatmaps_from_logits(cases.noise_logits(3, 1, 1e-3)),
    "noise_1e-1": lambda: d.heatmaps_from_logits(cases.noise_logits(3, 2, 1e-1)),
    "noise_1": lambda: d.heatmaps_from_logits(cases.noise_logits(3, 3, 1.0)),
    "uniform": lambda: cases.uniform_heatmaps(3, 4),
    "special": cases.special_heatmaps,
}


@pytest.mark.parametrize("name", list(FAMILIES))
def test_oracle_matches_reference_bit_for_bit(golden, name):
    hms = FAMILIES[name]()
    assert cases.checksum(hms) == str(golden[f"{name}/input_sha"]), "seeded inputs drifted from the golden run"
    n = min(len(hms), 4)  # scipy path costs ~30 ms / person
    for i in range(n):
        locs, vals = d.expected_value_decode(hms[i])
        assert locs.dtype == np.float32 and vals.dtype == np.float32
        np.testing.assert_array_equal(locs, golden[f"{name}/locs"][i])
        np.testing.assert_array_equal(vals, golden[f"{name}/vals"][i])
        kpts, scores = d.probmap_decode(hms[i])
        assert kpts.shape == (1, 17, 2) and kpts.dtype == np.float64 and scores.shape == (1, 17)
        np.testing.assert_array_equal(kpts, golden[f"{name}/keypoints"][i])


@pytest.mark.parametrize("name", list(FAMILIES))
def test_separable_batch_path_equals_reference(golden, name):
    hms = FAMILIES[name]()
    locs, vals = d.expected_value_decode_batch(hms)
    np.testing.assert_array_equal(locs, golden[f"{name}/locs"])
    np.testing.assert_array_equal(vals, golden[f"{name}/vals"])


def test_oks_kernels_match_reference(golden):
    kern = d.oks_kernels_2d(17, 64, 48)
    s, radius = d.oks_variances(17, 64, 48)
    assert [k.shape[-1] for k in kern] == list(golden["kernels/diam"]) == [2 * r + 1 for r in radius]
    for k in range(17):
        np.testing.assert_array_equal(kern[k], golden[f"kernels/{k}"])
        g = d.oks_kernel_1d(s[k], radius[k])
        np.testing.assert_allclose(np.outer(g, g), kern[k][0], rtol=0, atol=1e-15)  # exactly separable


def test_reference_rejects_batched_input(golden):
    # post_processing.py:352 raises for B > 1, hence the per-instance loop the kernel replaces
    assert bool(golden["batched_raises"])


def test_edge_cases_from_special_family(golden):
    locs, vals = golden["special/locs"][0], golden["special/vals"][0]
    assert tuple(locs[0]) == (0.0, 0.0) and vals[0] == 0.0          # all-zero map -> flat index 0
    assert tuple(locs[1]) == (0.0, 0.0) and vals[1] == 0.25         # constant map -> flat index 0
    assert tuple(locs[11]) == (10.5, 20.0)                          # exact horizontal tie -> midpoint
    assert tuple(locs[16]) == (47.0, 63.0) and vals[16] == 0.0      # reflect border pulls the peak to the corner


def test_sparsemax_properties():
    rng = np.random.default_rng(0)
    for std in (1e-3, 0.3, 4.0):
        z = rng.normal(0, std, (5, 3072)).astype(np.float32)
        p = d.sparsemax_rows(z)
        assert (p >= 0).all()
        np.testing.assert_allclose(p.sum(-1), 1.0, atol=2e-5)
        for zi, pi in zip(z, p):  # the support is a top-k set
            assert zi[pi > 0].min() >= zi[pi == 0].max() if (pi == 0).any() else True
    # independent check: Euclidean projection onto the simplex by bisection on tau
    z = rng.normal(0, 1, (4, 64)).astype(np.float64)
    lo, hi = z.max(-1, keepdims=True) - 1, z.max(-1, keepdims=True)
    for _ in range(60):
        mid = (lo + hi) / 2
        s = np.maximum(z - mid, 0).sum(-1, keepdims=True)
        lo, hi = np.where(s > 1, mid, lo), np.where(s > 1, hi, mid)
    np.testing.assert_allclose(d.sparsemax_rows(z.astype(np.float32)), np.maximum(z - lo, 0), atol=1e-6)


def test_tta_merge_matches_definition():
    rng = np.random.default_rng(1)
    p, pf = rng.random((2, 17, 4, 6), dtype=np.float32), rng.random((2, 17, 4, 6), dtype=np.float32)
    m = d.tta_merge(p, pf, d.COCO_FLIP_INDICES)
    for k in range(17):
        np.testing.assert_array_equal(m[:, k], (p[:, k] + pf[:, d.COCO_FLIP_INDICES[k], :, ::-1]) * np.float32(0.5))
