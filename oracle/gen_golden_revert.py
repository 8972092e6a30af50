#!/usr/bin/env python
"""Generate tests/golden/revert_kat.npz with the GENUINE reference geometry + the verbatim read-back lines.

Runs only in the build container (needs /root/reference and cv2).  ``mmpose/structures/bbox/transforms.py`` has no
package-relative imports, so ``get_warp_matrix`` is loaded from the reference file; ``mmpose/structures/utils.py`` imports
mmengine, so ``revert_heatmap`` (:146-175) and the padding / max-merge lines of ``merge_data_samples`` (:69-118) are
applied here verbatim.

    python oracle/gen_golden_revert.py
"""
import hashlib
import importlib.util
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import revert_oracle  # noqa: E402

REF_FILE = "/root/reference/mmpose/structures/bbox/transforms.py"


def main():
    spec = importlib.util.spec_from_file_location("ref_bbox_transforms", REF_FILE)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)

    def revert_heatmap(heatmap, input_center, input_scale, img_shape):  # utils.py:146-175
        ndim = heatmap.ndim
        if ndim == 3:
            heatmap = heatmap.transpose(1, 2, 0)
        hm_h, hm_w = heatmap.shape[:2]
        img_h, img_w = img_shape
        warp_mat = ref.get_warp_matrix(input_center.reshape((2,)), input_scale.reshape((2,)), rot=0, output_size=(hm_w, hm_h), inv=True)
        heatmap = cv2.warpAffine(heatmap, warp_mat, (img_w, img_h), flags=cv2.INTER_LINEAR)
        if ndim == 3:
            heatmap = heatmap.transpose(2, 0, 1)
        return heatmap

    out = {}
    for ci, (n, ih, iw) in enumerate([(3, 240, 320), (6, 333, 517), (1, 480, 640)]):
        hms, centers, scales = revert_oracle.synthetic_people(40 + ci, n, ih, iw)
        ori_shape = (ih, iw)
        max_image_pad = [0, 0, 0, 0]                                                            # utils.py:71-88
        for c, s in zip(centers, scales):
            img_pad = [int(max(s[0] / 2 - c[0] + 10, 0)), int(max(s[1] / 2 - c[1] + 10, 0)),
                       int(max(c[0] + s[0] / 2 - ori_shape[1] + 10, 0)), int(max(c[1] + s[1] / 2 - ori_shape[0] + 10, 0))]
            max_image_pad = np.maximum(max_image_pad, img_pad)
        padded = []
        for hm, c, s in zip(hms, centers, scales):                                              # :90-113
            aw_center = c + np.array([max_image_pad[0], max_image_pad[1]])
            padded_img_shape = (ori_shape[0] + max_image_pad[1] + max_image_pad[3], ori_shape[1] + max_image_pad[0] + max_image_pad[2])
            padded.append(revert_heatmap(hm, aw_center, s, padded_img_shape))
        merged = np.max(padded, axis=0)                                                         # :117
        out[f"c{ci}/pad"] = np.asarray(max_image_pad)
        out[f"c{ci}/merged_sha"] = np.array(hashlib.sha256(np.ascontiguousarray(merged).tobytes()).hexdigest()[:16])
        out[f"c{ci}/merged_k0"] = merged[0]
        out[f"c{ci}/mat0"] = ref.get_warp_matrix(centers[0], scales[0], rot=0, output_size=(48, 64), inv=True)
        out[f"c{ci}/single0"] = revert_heatmap(hms[0], centers[0], scales[0], ori_shape)[3]
        print(ci, merged.shape, "pad", max_image_pad, "nonzero", (merged != 0).mean())
    out["cv2_version"] = np.array(cv2.__version__)
    path = os.path.join(ROOT, "tests", "golden", "revert_kat.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
