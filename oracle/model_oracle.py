"""Plain-PyTorch fp32 restatement of the ProbPose forward (TEST INFRASTRUCTURE, see
oracle/__init__.py).  PARITY UNPINNED by the reference's tests - it has none for
this path - so every function cites what it follows.

* ViT backbone: ``mmpretrain==1.2.0`` ``VisionTransformer`` built by config
  ``configs/body_2d_keypoint/topdown_probmap/coco/td-pm_ProbPose-small_8xb64-210e_coco-256x192.py:56-67``
  (third-party, not in /root/reference; in-tree witnesses: patch-embed twin
  ``mmpose/models/utils/transformer.py:153-245``, parameter prefixes
  ``mmpose/engine/optim_wrappers/layer_decay_optim_wrapper.py:8-14``).
* Head: ``mmpose/models/heads/hybrid_heads/probmap_head.py`` - heatmap stack
  ``:197-259,435-472``, scalar branches ``:261-410``, ``forward`` ``:600-625``,
  ``forward_heatmap`` ``:627-648``, ``predict`` ``:715-804``.
* Sparsemax: PyPI ``sparsemax`` (un-vendored), Martins & Astudillo 2016 Alg. 1.
* Flip-TTA: ``mmpose/models/pose_estimators/topdown.py:109-112``,
  ``mmpose/models/utils/tta.py:35-39``.

Module/parameter names mirror the reference so one MMPose-format ``state_dict``
(``backbone.*``, ``head.*``) loads into both this oracle and the CUDA engine.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F
from torch import nn

from . import decode_oracle

PIXEL_MEAN = (123.675, 116.28, 103.53)  # config :53-55, RGB order
PIXEL_STD = (58.395, 57.12, 57.375)
POOL_KERNELS = ((4, 3), (2, 2), (2, 2))  # probmap_head.py:264


def sparsemax(z: torch.Tensor) -> torch.Tensor:
    """Sparsemax over the last dim (sort + cumsum), fp32."""
    z = z - z.max(dim=-1, keepdim=True).values
    zs = torch.sort(z, dim=-1, descending=True).values
    rng = torch.arange(1, z.shape[-1] + 1, dtype=z.dtype, device=z.device)
    in_support = (1 + rng * zs) > torch.cumsum(zs, dim=-1)
    k = (in_support * rng).max(dim=-1, keepdim=True).values
    tau = ((in_support * zs).sum(dim=-1, keepdim=True) - 1) / k
    return torch.clamp(z - tau, min=0)


class _Attention(nn.Module):
    """mmpretrain ``MultiheadAttention``: packed qkv Linear, SDPA, proj Linear."""

    def __init__(self, dim, heads):
        super().__init__()
        self.heads = heads
        self.qkv = nn.Linear(dim, 3 * dim, bias=True)
        self.proj = nn.Linear(dim, dim)

    def forward(self, x):
        b, n, c = x.shape
        qkv = self.qkv(x).reshape(b, n, 3, self.heads, c // self.heads).permute(2, 0, 3, 1, 4)
        o = F.scaled_dot_product_attention(qkv[0], qkv[1], qkv[2])  # scale = d_h ** -0.5
        return self.proj(o.transpose(1, 2).reshape(b, n, c))


class _FFN(nn.Module):
    """mmcv ``FFN``: layers = Sequential(Sequential(Linear, GELU, Dropout), Linear, Dropout)."""

    def __init__(self, dim, hidden):
        super().__init__()
        self.layers = nn.Sequential(nn.Sequential(nn.Linear(dim, hidden), nn.GELU()), nn.Linear(hidden, dim))

    def forward(self, x, identity):
        return identity + self.layers(x)


class _Block(nn.Module):
    def __init__(self, dim, heads, hidden):
        super().__init__()
        self.ln1 = nn.LayerNorm(dim, eps=1e-6)
        self.attn = _Attention(dim, heads)
        self.ln2 = nn.LayerNorm(dim, eps=1e-6)
        self.ffn = _FFN(dim, hidden)

    def forward(self, x):
        x = x + self.attn(self.ln1(x))
        return self.ffn(self.ln2(x), identity=x)


class VisionTransformerRef(nn.Module):
    """``with_cls_token=False, out_type='featmap', final_norm=True, patch padding 2``."""

    def __init__(self, embed_dims=384, num_layers=12, num_heads=12, feedforward_channels=1536,
                 img_size=(256, 192), patch_size=16, padding=2):
        super().__init__()
        self.patch_embed = nn.Module()
        self.patch_embed.projection = nn.Conv2d(3, embed_dims, patch_size, patch_size, padding)
        self.grid = tuple((s + 2 * padding - patch_size) // patch_size + 1 for s in img_size)
        self.pos_embed = nn.Parameter(torch.zeros(1, self.grid[0] * self.grid[1], embed_dims))
        self.layers = nn.ModuleList(_Block(embed_dims, num_heads, feedforward_channels) for _ in range(num_layers))
        self.ln1 = nn.LayerNorm(embed_dims, eps=1e-6)

    def forward(self, x):
        b = x.shape[0]
        x = self.patch_embed.projection(x).flatten(2).transpose(1, 2) + self.pos_embed
        for blk in self.layers:
            x = blk(x)
        x = self.ln1(x)
        return (x.reshape(b, *self.grid, -1).permute(0, 3, 1, 2),)


def _scalar_branch(cin, cout, last):
    mods = []
    for ks in POOL_KERNELS:  # Conv -> BN -> MaxPool -> ReLU, probmap_head.py:266-278
        mods += [nn.Conv2d(cin, cin, 3, 1, 1), nn.BatchNorm2d(cin), nn.MaxPool2d(ks, ks), nn.ReLU()]
    mods += [nn.Conv2d(cin, cout, 1), last]
    return nn.Sequential(*mods)


class ProbMapHeadRef(nn.Module):
    temperature = 0.5  # probmap_head.py:135

    def __init__(self, in_channels=384, out_channels=17, deconv_out_channels=(256, 256), normalize=1.0):
        super().__init__()
        layers, c = [], in_channels
        for co in deconv_out_channels:  # k4 s2 p1, no bias, probmap_head.py:441-470
            layers += [nn.ConvTranspose2d(c, co, 4, 2, 1, 0, bias=False), nn.BatchNorm2d(co), nn.ReLU()]
            c = co
        self.deconv_layers = nn.Sequential(*layers)
        self.final_layer = nn.Conv2d(c, out_channels, 1)
        self.normalize = normalize
        self.probability_layers = _scalar_branch(in_channels, out_channels, nn.Sigmoid())
        self.visibility_layers = _scalar_branch(in_channels, out_channels, nn.Sigmoid())
        self.oks_layers = _scalar_branch(in_channels, out_channels, nn.Sigmoid())
        self.error_layers = _scalar_branch(in_channels, out_channels, nn.ReLU())

    def heatmap_logits(self, x):
        return self.final_layer(self.deconv_layers(x))

    def forward_heatmap(self, x):
        z = self.heatmap_logits(x)
        b, c, h, w = z.shape
        p = sparsemax(z.reshape(b, c, h * w) / self.temperature)
        if self.normalize is not None:
            p = p * self.normalize
        return torch.clamp(p, 0, 1).reshape(b, c, h, w)

    def forward(self, feats):
        x = feats[-1]
        return (self.forward_heatmap(x), self.probability_layers(x), self.visibility_layers(x),
                self.oks_layers(x), self.error_layers(x))

    @torch.no_grad()
    def merged(self, feats, feats_flip=None, flip_indices=decode_oracle.COCO_FLIP_INDICES):
        """``predict`` up to the decode (probmap_head.py:746-774; ``flip_heatmaps`` tta.py:35-39 with
        flip_mode="heatmap", shift_heatmap=False): merged (heatmaps, prob, vis, oks, err), scalars (B, K)."""
        out = self.forward(feats)
        if feats_flip is not None:
            outf = self.forward(feats_flip)
            htm = (out[0] + outf[0].flip(-1)[:, flip_indices]) * 0.5
            scal = [(a + b[:, flip_indices]) * 0.5 for a, b in zip(out[1:], outf[1:])]
        else:
            htm, scal = out[0], list(out[1:])
        return (htm, *[s.flatten(1) for s in scal])

    @torch.no_grad()
    def predict(self, feats, feats_flip=None, flip_indices=decode_oracle.COCO_FLIP_INDICES, input_size=(192, 256),
                timings=None):
        """``ProbMapHead.predict`` (probmap_head.py:715-804) with the ProbMap codec through ``BaseHead.decode``'s
        per-person loop (base_head.py:64-78).  Returns the record array (B, K, 7) float64
        [x, y, conf, prob, vis, oks, err/diag] and the merged heatmaps."""
        import time

        t0 = time.perf_counter()
        htm, prob, vis, oks, err = self.merged(feats, feats_flip, flip_indices)
        t1 = time.perf_counter()
        h, w = htm.shape[-2:]
        kpts, conf = decode_oracle.decode_instances(htm.numpy(), input_size=input_size, heatmap_size=(w, h))
        if timings is not None:
            timings["model_s"] = timings.get("model_s", 0.0) + t1 - t0
            timings["decode_s"] = time.perf_counter() - t1
        rec = np.zeros((htm.shape[0], htm.shape[1], 7))
        rec[..., 0:2] = np.concatenate(kpts, 0)
        rec[..., 2] = np.concatenate(conf, 0)
        rec[..., 3] = prob.numpy()
        rec[..., 4] = vis.numpy()
        rec[..., 5] = oks.numpy()
        rec[..., 6] = err.numpy() / np.sqrt(h**2 + w**2)  # probmap_head.py:786-787 (float32 / float64 scalar)
        return rec, htm



class ProbPoseRef(nn.Module):
    """TopdownPoseEstimator(backbone=ViT, head=ProbMapHead) restricted to ``predict``."""

    def __init__(self, head_kwargs=None, **vit_kwargs):
        super().__init__()
        self.backbone = VisionTransformerRef(**vit_kwargs)
        self.head = ProbMapHeadRef(in_channels=self.backbone.ln1.normalized_shape[0], **(head_kwargs or {}))

    @staticmethod
    def preprocess(crops_u8_bgr: torch.Tensor) -> torch.Tensor:
        """PoseDataPreprocessor (data_preprocessor.py:79-104 + mmengine ImgDataPreprocessor):
        BGR->RGB, float, (x - mean) / std."""
        x = crops_u8_bgr[:, [2, 1, 0]].float()
        mean = torch.tensor(PIXEL_MEAN).view(1, 3, 1, 1)
        std = torch.tensor(PIXEL_STD).view(1, 3, 1, 1)
        return (x - mean) / std

    @torch.no_grad()
    def forward_merged(self, inputs, flip_test=True, flip_indices=decode_oracle.COCO_FLIP_INDICES):
        """Returns merged (heatmaps, prob, vis, oks, err) tensors, err NOT yet /diag
        (topdown.py:109-112 + probmap_head.py:746-774)."""
        feats = self.backbone(inputs)
        feats_flip = self.backbone(inputs.flip(-1)) if flip_test else None
        return self.head.merged(feats, feats_flip, flip_indices)

    @torch.no_grad()
    def predict(self, inputs, flip_test=True, flip_indices=decode_oracle.COCO_FLIP_INDICES, timings=None):
        """Returns the per-person record array (B, K, 7) float64:
        [x, y, conf, prob, vis, oks, err/diag] - SURVEY A.5 - using the reference's
        per-person CPU decode."""
        import time

        t0 = time.perf_counter()
        feats = self.backbone(inputs)
        feats_flip = self.backbone(inputs.flip(-1)) if flip_test else None
        if timings is not None:
            timings["model_s"] = time.perf_counter() - t0
        return self.head.predict(feats, feats_flip, flip_indices, timings=timings)[0]


class ViTPoseRef(nn.Module):
    """TopdownPoseEstimator(backbone=ViT, head=HeatmapHead, codec=UDPHeatmap) restricted to ``predict``
    (configs/body_2d_keypoint/topdown_heatmap/coco/td-hm_ViTPose-*_coco-256x192.py; heatmap_head.py:197-268).
    The head's deconv stack and final layer are the same modules as ProbMapHead's heatmap branch."""

    def __init__(self, **vit_kwargs):
        super().__init__()
        self.backbone = VisionTransformerRef(**vit_kwargs)
        self.head = ProbMapHeadRef(in_channels=self.backbone.ln1.normalized_shape[0])

    preprocess = staticmethod(ProbPoseRef.preprocess)

    def load_state_dict(self, sd, strict=True):  # HeatmapHead checkpoints carry no scalar branches
        own = self.state_dict()
        own.update({k: v for k, v in sd.items() if k in own})
        return super().load_state_dict(own, strict=strict)

    @torch.no_grad()
    def heatmaps(self, inputs, flip_test=True, flip_indices=decode_oracle.COCO_FLIP_INDICES):
        hm = self.head.heatmap_logits(self.backbone(inputs)[-1])
        if flip_test:  # heatmap_head.py:245-256, tta.py:35-39
            hmf = self.head.heatmap_logits(self.backbone(inputs.flip(-1))[-1])
            hm = (hm + hmf.flip(-1)[:, flip_indices]) * 0.5
        return hm

    @torch.no_grad()
    def predict(self, inputs, flip_test=True):
        """Records (B, K, 3) float64: x, y in input pixels, score - through the restated reference decode."""
        from . import udp_oracle

        hm = self.heatmaps(inputs, flip_test).numpy()
        kpts, scores = udp_oracle.decode_instances(hm)
        rec = np.zeros((hm.shape[0], hm.shape[1], 3))
        rec[..., :2] = np.concatenate(kpts, 0)
        rec[..., 2] = np.concatenate(scores, 0)
        return rec
