"""Host-side mirror of the MMPose plugin API for the ProbPose top-down inference path.

Importing this package registers ``VisionTransformer`` (also as
``mmpretrain.VisionTransformer``), ``ProbMapHead``, ``TopdownPoseEstimator``,
``PoseDataPreprocessor`` in ``MODELS`` and ``ProbMap`` in ``KEYPOINT_CODECS`` - mmpose's own
registries when mmpose is installed, a compatible stand-in otherwise."""
from .backbone import VisionTransformer
from .codec import BaseKeypointCodec, ProbMap, UDPHeatmap
from .estimator import PoseDataPreprocessor, TopdownPoseEstimator
from .head import BaseHead, HeatmapHead, ProbMapHead
from .inference import TopdownAffine, inference_topdown
from .registry import HAVE_MMPOSE, KEYPOINT_CODECS, MODELS, Registry
from .structures import InstanceData, PixelData, PoseDataSample
from .utils import get_warp_matrix, merge_data_samples, revert_heatmap

COCO_FLIP_INDICES = [0, 2, 1, 4, 3, 6, 5, 8, 7, 10, 9, 12, 11, 14, 13, 16, 15]  # configs/_base_/datasets/coco.py:14-30


def probpose_small_cfg(precision: str = None, flip_test: bool = True) -> dict:
    """``model = dict(...)`` of td-pm_ProbPose-small_8xb64-210e_coco-256x192.py:48-92, verbatim
    apart from the optional ``precision`` key this package adds."""
    codec = dict(type="ProbMap", input_size=(192, 256), heatmap_size=(48, 64), sigma=-1)
    cfg = dict(
        type="TopdownPoseEstimator",
        data_preprocessor=dict(type="PoseDataPreprocessor", mean=[123.675, 116.28, 103.53], std=[58.395, 57.12, 57.375],
                               bgr_to_rgb=True),
        backbone=dict(type="mmpretrain.VisionTransformer",
                      arch={"embed_dims": 384, "num_layers": 12, "num_heads": 12, "feedforward_channels": 384 * 4},
                      img_size=(256, 192), patch_size=16, qkv_bias=True, drop_path_rate=0.1, with_cls_token=False,
                      out_type="featmap", patch_cfg=dict(padding=2), init_cfg=None),
        head=dict(type="ProbMapHead", in_channels=384, out_channels=17, deconv_out_channels=(256, 256),
                  deconv_kernel_sizes=(4, 4),
                  keypoint_loss=dict(type="OKSHeatmapLoss", use_target_weight=True, smoothing_weight=0.05),
                  probability_loss=dict(type="BCELoss", use_target_weight=True, use_sigmoid=True),
                  visibility_loss=dict(type="BCELoss", use_target_weight=True, use_sigmoid=True),
                  oks_loss=dict(type="MSELoss", use_target_weight=True),
                  error_loss=dict(type="L1LogLoss", use_target_weight=True), detach_probability=True,
                  detach_visibility=True, normalize=1.0, freeze_error=True, freeze_oks=False, decoder=codec),
        test_cfg=dict(flip_test=flip_test, flip_mode="heatmap", shift_heatmap=False),
    )
    if precision is not None:
        cfg["precision"] = precision
    return cfg


def vitpose_cfg(arch: str = "base", precision: str = None, flip_test: bool = True) -> dict:
    """``model = dict(...)`` of configs/body_2d_keypoint/topdown_heatmap/coco/td-hm_ViTPose-{small,base}_8xb64-210e_coco-256x192.py
    (:40-75): ViT backbone + HeatmapHead + UDPHeatmap codec (SURVEY.md 8f rank 3 / BASELINE config 5)."""
    dims = {"small": dict(embed_dims=384, num_layers=12, num_heads=12, feedforward_channels=384 * 4),
            "base": dict(embed_dims=768, num_layers=12, num_heads=12, feedforward_channels=768 * 4)}[arch]
    codec = dict(type="UDPHeatmap", input_size=(192, 256), heatmap_size=(48, 64), sigma=2)
    cfg = dict(
        type="TopdownPoseEstimator",
        data_preprocessor=dict(type="PoseDataPreprocessor", mean=[123.675, 116.28, 103.53], std=[58.395, 57.12, 57.375],
                               bgr_to_rgb=True),
        backbone=dict(type="mmpretrain.VisionTransformer", arch=dims, img_size=(256, 192), patch_size=16, qkv_bias=True,
                      drop_path_rate=0.3 if arch == "base" else 0.1, with_cls_token=False, out_type="featmap",
                      patch_cfg=dict(padding=2), init_cfg=None),
        head=dict(type="HeatmapHead", in_channels=dims["embed_dims"], out_channels=17, deconv_out_channels=(256, 256),
                  deconv_kernel_sizes=(4, 4), loss=dict(type="KeypointMSELoss", use_target_weight=True), decoder=codec),
        test_cfg=dict(flip_test=flip_test, flip_mode="heatmap", shift_heatmap=False),
    )
    if precision is not None:
        cfg["precision"] = precision
    return cfg


def make_data_samples(batch: int, input_size=(192, 256), flip_indices=COCO_FLIP_INDICES):
    """Data samples for crops that ARE the whole image (bbox = image, like demo/image_demo.py)."""
    import numpy as np

    w, h = input_size
    out = []
    for _ in range(batch):
        ds = PoseDataSample(metainfo=dict(input_size=np.array([w, h], np.float32),
                                          input_center=np.array([w / 2, h / 2], np.float32),
                                          input_scale=np.array([w, h], np.float32), flip_indices=list(flip_indices)))
        ds.gt_instances = InstanceData(bboxes=np.array([[0, 0, w, h]], np.float32), bbox_scores=np.ones(1, np.float32))
        out.append(ds)
    return out
