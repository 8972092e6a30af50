#!/usr/bin/env python
"""Generate tests/golden/decode_kat.npz by running the GENUINE reference decode.

Runs only in the build container (needs /root/reference).  It loads
``mmpose/codecs/utils/post_processing.py`` by file path (the file has no
package-relative imports, so it works without mmengine/mmcv) and records what
``get_heatmap_expected_value`` / ``_prepare_oks_kernels`` return for the seeded input
families in ``oracle/cases.py``.  The one remaining line of ``ProbMap.decode``
(probmap.py:218, ``keypoints / [W-1, H-1] * input_size``) is applied here verbatim
because ``probmap.py`` itself imports the mmpose registry and cannot be loaded.

    python oracle/gen_golden.py            # rewrites tests/golden/decode_kat.npz
"""
import importlib.util
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import cases, decode_oracle  # noqa: E402

REF_FILE = "/root/reference/mmpose/codecs/utils/post_processing.py"


def load_reference():
    spec = importlib.util.spec_from_file_location("ref_post_processing", REF_FILE)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    ref = load_reference()
    families = {
        "planted": decode_oracle.heatmaps_from_logits(cases.planted_peak_logits(12, seed=0)),
        "noise_1e-3": decode_oracle.heatmaps_from_logits(cases.noise_logits(3, 1, 1e-3)),
        "noise_1e-1": decode_oracle.heatmaps_from_logits(cases.noise_logits(3, 2, 1e-1)),
        "noise_1": decode_oracle.heatmaps_from_logits(cases.noise_logits(3, 3, 1.0)),
        "uniform": cases.uniform_heatmaps(3, 4),
        "special": cases.special_heatmaps(),
    }
    out = {}
    for name, hms in families.items():
        locs, vals, kpts = [], [], []
        for hm in hms:
            l, v = ref.get_heatmap_expected_value(hm.copy())
            locs.append(l)
            vals.append(v)
            kpts.append(l[None] / [cases.W - 1, cases.H - 1] * (192, 256))  # probmap.py:218
        out[f"{name}/locs"] = np.stack(locs)
        out[f"{name}/vals"] = np.stack(vals)
        out[f"{name}/keypoints"] = np.stack(kpts)
        out[f"{name}/input_sha"] = np.array(cases.checksum(hms))
        print(name, hms.shape, "nonzero frac %.4f" % (hms != 0).mean())
    # the hand-built maps are tiny once compressed: store them too
    out["special/heatmaps"] = families["special"]
    kern = ref._prepare_oks_kernels(cases.K, cases.H, cases.W)
    out["kernels/diam"] = np.array([k.shape[-1] for k in kern])
    for k, kk in enumerate(kern):
        out[f"kernels/{k}"] = kk
    # B>1 must raise inside the reference (post_processing.py:352) -> record that fact
    try:
        ref.get_heatmap_expected_value(families["uniform"][:2].copy())
        out["batched_raises"] = np.array(False)
    except Exception:  # noqa: BLE001
        out["batched_raises"] = np.array(True)
    path = os.path.join(ROOT, "tests", "golden", "decode_kat.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
