#!/usr/bin/env python
"""Generate tests/golden/udp_kat.npz by running the GENUINE reference UDPHeatmap decode.

Runs only in the build container (needs /root/reference).  ``mmpose/codecs/utils/post_processing.py`` has no
package-relative imports and is loaded by file path; ``mmpose/codecs/utils/refinement.py`` imports
``.post_processing`` relatively, so both are loaded into a throw-away package ``_ref_codecs_utils``.  The remaining
lines of ``UDPHeatmap.decode`` (udp_heatmap.py:161-168,194-195) are applied here verbatim because ``udp_heatmap.py``
imports the mmpose registry and cannot be loaded without mmengine.

    python oracle/gen_golden_udp.py            # rewrites tests/golden/udp_kat.npz
"""
import importlib.util
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import cases, udp_oracle  # noqa: E402

REF_DIR = "/root/reference/mmpose/codecs/utils"


def load_reference():
    pkg = types.ModuleType("_ref_codecs_utils")
    pkg.__path__ = [REF_DIR]
    sys.modules["_ref_codecs_utils"] = pkg
    mods = {}
    for name in ("post_processing", "refinement"):
        spec = importlib.util.spec_from_file_location(f"_ref_codecs_utils.{name}", os.path.join(REF_DIR, f"{name}.py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[spec.name] = mod
        spec.loader.exec_module(mod)
        mods[name] = mod
    return mods["post_processing"], mods["refinement"]


def reference_decode(pp, rf, heatmaps, input_size=(192, 256), blur_kernel_size=11):
    hm = heatmaps.copy()                                   # udp_heatmap.py:159
    keypoints, scores = pp.get_heatmap_maximum(hm)         # :162
    keypoints = keypoints[None]                            # :164
    scores = scores[None]
    keypoints = rf.refine_keypoints_dark_udp(keypoints, hm, blur_kernel_size=blur_kernel_size)  # :167
    h, w = heatmaps.shape[1:]
    keypoints = keypoints / [w - 1, h - 1] * input_size    # :194-195
    return keypoints, scores


def main():
    pp, rf = load_reference()
    fams = {
        "gauss": udp_oracle.gaussian_heatmaps(8, seed=0),
        "gauss_noisy": udp_oracle.gaussian_heatmaps(4, seed=1, noise=0.05),
        "special": udp_oracle.special_heatmaps(),
        "noresp": udp_oracle.no_response_heatmaps(),
    }
    a, b = udp_oracle.gaussian_heatmaps(4, seed=2), udp_oracle.gaussian_heatmaps(4, seed=2, noise=0.02)
    inv = np.argsort(udp_oracle.COCO_FLIP_INDICES)
    fams["merged"] = udp_oracle.merge_flip(a, np.ascontiguousarray(b[:, inv][..., ::-1]))
    out = {}
    for name, hms in fams.items():
        kp, sc = [], []
        for hm in hms:
            k, s = reference_decode(pp, rf, hm)
            kp.append(k)
            sc.append(s)
        out[f"{name}/keypoints"] = np.stack(kp)
        out[f"{name}/scores"] = np.stack(sc)
        out[f"{name}/input_sha"] = np.array(cases.checksum(hms))
        print(name, hms.shape, out[f"{name}/keypoints"].dtype, out[f"{name}/scores"].dtype)
    blurred = pp.gaussian_blur(fams["gauss"][0].copy(), 11)
    out["gauss/blurred0"] = blurred
    path = os.path.join(ROOT, "tests", "golden", "udp_kat.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
