"""pp_revert_heatmaps / merge_data_samples against the oracle, which is pinned bit-for-bit to the genuine
reference geometry + cv2.warpAffine (tests/test_oracle_revert.py).  Bit-exact: OpenCV's CV_32F warp is fixed-point
coordinates + float weights in a defined order."""
import numpy as np
import pytest
import torch

from oracle import revert_oracle as r

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed,n,ih,iw", [(40, 3, 240, 320), (41, 6, 333, 517), (42, 1, 480, 640), (43, 9, 720, 1280)])
def test_merge_data_samples_matches_reference(seed, n, ih, iw):
    import probpose_code_b200.mmpose_api as api
    hms, centers, scales = r.synthetic_people(seed, n, ih, iw)
    samples = []
    for hm, c, s in zip(hms, centers, scales):
        ds = api.PoseDataSample(metainfo=dict(input_center=c, input_scale=s, input_size=np.array([192, 256]), ori_shape=(ih, iw)))
        ds.pred_fields = api.PixelData(heatmaps=torch.from_numpy(hm).cuda())
        ds.pred_instances = api.InstanceData(keypoints=np.zeros((1, 17, 2)), keypoint_scores=np.ones((1, 17), np.float32))
        ds.gt_instances = api.InstanceData(bboxes=np.zeros((1, 4), np.float32))
        samples.append(ds)
    merged = api.merge_data_samples(samples)
    ref, pad = r.merged_padded_heatmaps(hms, centers, scales, (ih, iw))
    got = merged.pred_fields.heatmaps
    assert got.shape == ref.shape and got.dtype == np.float32
    np.testing.assert_array_equal(got, ref)
    assert merged.pred_instances.keypoints.shape == (n, 17, 2) and merged.gt_instances.bboxes.shape == (n, 4)
    assert merged.metainfo["input_center"].shape == (n, 2)


def test_revert_heatmap_single_and_negative_values():
    import probpose_code_b200.mmpose_api as api
    hms, centers, scales = r.synthetic_people(44, 2, 300, 400)
    hms[0] -= 0.1  # HeatmapHead outputs may be negative: the border value 0 then wins outside the footprint
    for i in range(2):
        got = api.revert_heatmap(hms[i], centers[i], scales[i], (300, 400))
        np.testing.assert_array_equal(got, r.revert_heatmap(hms[i], centers[i], scales[i], (300, 400)))
    np.testing.assert_array_equal(api.get_warp_matrix(centers[0], scales[0], 0, (48, 64), inv=True),
                                  r.get_warp_matrix(centers[0], scales[0], 0, (48, 64), inv=True))
    with pytest.raises(ValueError):
        api.merge_data_samples([1, 2])
