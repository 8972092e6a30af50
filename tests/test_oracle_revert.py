"""The heatmap read-back oracle (revert_heatmap + padded max-merge) against golden vectors made with the
genuine reference get_warp_matrix + the verbatim utils.py lines (oracle/gen_golden_revert.py).  CPU only."""
import hashlib
import os

import numpy as np
import pytest

from oracle import revert_oracle as r

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "revert_kat.npz")
CASES = [(3, 240, 320), (6, 333, 517), (1, 480, 640)]


@pytest.mark.parametrize("ci", [0, 1, 2])
def test_revert_oracle_matches_reference_bit_for_bit(ci):
    g = np.load(GOLDEN)
    n, ih, iw = CASES[ci]
    hms, centers, scales = r.synthetic_people(40 + ci, n, ih, iw)
    merged, pad = r.merged_padded_heatmaps(hms, centers, scales, (ih, iw))
    np.testing.assert_array_equal(pad, g[f"c{ci}/pad"])
    assert hashlib.sha256(np.ascontiguousarray(merged).tobytes()).hexdigest()[:16] == str(g[f"c{ci}/merged_sha"])
    np.testing.assert_array_equal(merged[0], g[f"c{ci}/merged_k0"])
    np.testing.assert_array_equal(r.get_warp_matrix(centers[0], scales[0], 0, (48, 64), inv=True), g[f"c{ci}/mat0"])
    np.testing.assert_array_equal(r.revert_heatmap(hms[0], centers[0], scales[0], (ih, iw))[3], g[f"c{ci}/single0"])
