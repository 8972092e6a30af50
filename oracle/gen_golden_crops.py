#!/usr/bin/env python
"""Generate tests/golden/crop_kat.npz with the GENUINE reference crop front-end.

Runs only in the build container (needs /root/reference and cv2).  ``mmpose/structures/bbox/transforms.py``
has no package-relative imports, so ``bbox_xyxy2cs`` / ``get_udp_warp_matrix`` are loaded from the
reference file; ``TopdownAffine`` itself imports mmcv, so its three geometry lines
(topdown_transforms.py:93-118) are applied here verbatim, followed by the reference's own call
``cv2.warpAffine(img, warp_mat, (w, h), flags=cv2.INTER_LINEAR)`` (:126).

    python oracle/gen_golden_crops.py
"""
import hashlib
import importlib.util
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import crop_oracle  # noqa: E402

REF_FILE = "/root/reference/mmpose/structures/bbox/transforms.py"


def main():
    spec = importlib.util.spec_from_file_location("ref_bbox_transforms", REF_FILE)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    w, h = crop_oracle.INPUT_SIZE
    out = {}
    for fi, (fh, fw) in enumerate([(480, 640), (333, 517), (1080, 1920)]):
        frame = crop_oracle.synthetic_frame(10 + fi, fh, fw)
        boxes = crop_oracle.synthetic_boxes(20 + fi, 12, fh, fw)
        boxes[0] = [0, 0, fw, fh]  # inference_topdown's default: the whole image (inference.py:161-168)
        crops, mats, centers, scales = [], [], [], []
        for bbox in boxes:
            c, s = ref.bbox_xyxy2cs(bbox[None], padding=crop_oracle.INPUT_PADDING)          # topdown_transforms.py:95
            s = s.reshape(1, 2)
            ww, hh = np.hsplit(s, [1])                                                     # _fix_aspect_ratio :65-67
            s = np.where(ww > hh * (w / h), np.hstack([ww, ww / (w / h)]), np.hstack([hh * (w / h), hh]))
            center, scale = c.reshape(1, 2)[0], s[0]
            m = ref.get_udp_warp_matrix(center, scale, 0.0, output_size=(w, h))             # :115
            crops.append(cv2.warpAffine(frame, m, (int(w), int(h)), flags=cv2.INTER_LINEAR))  # :126
            mats.append(m); centers.append(center); scales.append(scale)
        crops = np.stack(crops)
        out[f"f{fi}/boxes"] = boxes
        out[f"f{fi}/mats"] = np.stack(mats)
        out[f"f{fi}/centers"] = np.stack(centers)
        out[f"f{fi}/scales"] = np.stack(scales)
        out[f"f{fi}/crop_sha"] = np.array([hashlib.sha256(c.tobytes()).hexdigest()[:16] for c in crops])
        out[f"f{fi}/frame_sha"] = np.array(hashlib.sha256(frame.tobytes()).hexdigest()[:16])
        out[f"f{fi}/crops_head"] = crops[:2]  # two full crops per frame (HWC BGR), the rest by checksum
        print(fi, frame.shape, "crops", crops.shape, "mean", crops.mean())
    out["cv2_version"] = np.array(cv2.__version__)
    out["numpy_version"] = np.array(np.__version__)
    path = os.path.join(ROOT, "tests", "golden", "crop_kat.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
