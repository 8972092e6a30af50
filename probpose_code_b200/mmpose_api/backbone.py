"""``VisionTransformer`` with the constructor and ``state_dict`` layout of
``mmpretrain==1.2.0`` ``mmpretrain.models.backbones.VisionTransformer`` as the ProbPose /
ViTPose configs use it (td-pm_ProbPose-small config :56-67, td-hm_ViTPose-base :46-61):
patch-embed conv (padding via ``patch_cfg``), learned ``pos_embed``, pre-LN blocks with packed
qkv, erf-GELU FFN, final LayerNorm, ``out_type="featmap"``, no cls token.  Forward runs on the
sm_100a engine; the ``nn.Parameter``s here only hold the weights."""
from __future__ import annotations

from typing import Tuple

import torch
from torch import nn

from ._engine_cache import EngineCache
from .registry import MODELS, register

ARCH_ZOO = {  # mmpretrain arch_zoo entries relevant to the pose configs
    **dict.fromkeys(["s", "small"], dict(embed_dims=768, num_layers=8, num_heads=8, feedforward_channels=768 * 3)),
    **dict.fromkeys(["b", "base"], dict(embed_dims=768, num_layers=12, num_heads=12, feedforward_channels=3072)),
    **dict.fromkeys(["l", "large"], dict(embed_dims=1024, num_layers=24, num_heads=16, feedforward_channels=4096)),
}


class _Params(nn.Module):
    pass


@register(MODELS, ["VisionTransformer", "mmpretrain.VisionTransformer"])
class VisionTransformer(nn.Module):
    def __init__(self, arch="base", img_size=224, patch_size=16, in_channels=3, out_indices=-1, drop_rate=0.0,
                 drop_path_rate=0.0, qkv_bias=True, norm_cfg=dict(type="LN", eps=1e-6), final_norm=True,
                 out_type="cls_token", with_cls_token=True, frozen_stages=-1, interpolate_mode="bicubic",
                 layer_scale_init_value=0.0, patch_cfg=dict(), layer_cfgs=dict(), pre_norm=False, init_cfg=None,
                 precision: str = None):
        super().__init__()
        if isinstance(arch, str):
            if arch.lower() not in ARCH_ZOO:
                raise ValueError(f"Arch {arch} is not in default archs {set(ARCH_ZOO)}")
            arch = ARCH_ZOO[arch.lower()]
        else:
            essential = {"embed_dims", "num_layers", "num_heads", "feedforward_channels"}
            if not (isinstance(arch, dict) and essential <= set(arch)):
                raise ValueError(f"Custom arch needs a dict with keys {essential}")
        self.arch_settings = dict(arch)
        self.embed_dims, self.num_layers = arch["embed_dims"], arch["num_layers"]
        self.img_size = (img_size, img_size) if isinstance(img_size, int) else tuple(img_size)
        self.patch_size = patch_size
        self.padding = int(dict(patch_cfg).get("padding", 0))
        unsupported = []
        if in_channels != 3: unsupported.append("in_channels != 3")
        if out_type != "featmap": unsupported.append(f'out_type="{out_type}" (only "featmap")')
        if with_cls_token: unsupported.append("with_cls_token=True")
        if not final_norm: unsupported.append("final_norm=False")
        if not qkv_bias: unsupported.append("qkv_bias=False")
        if pre_norm: unsupported.append("pre_norm=True")
        if layer_scale_init_value: unsupported.append("layer scale")
        if out_indices not in (-1, [-1], (-1,), self.num_layers - 1): unsupported.append("intermediate out_indices")
        if unsupported:
            raise NotImplementedError("probpose_code_b200 VisionTransformer covers the pose-estimation configuration "
                                      "only; unsupported: " + ", ".join(unsupported))
        d, ff = self.embed_dims, arch["feedforward_channels"]
        self.grid = tuple((s + 2 * self.padding - patch_size) // patch_size + 1 for s in self.img_size)
        self.eps = float(dict(norm_cfg).get("eps", 1e-5))
        # ---- parameters, mmpretrain names ----
        self.patch_embed = _Params()
        self.patch_embed.projection = nn.Conv2d(3, d, patch_size, patch_size, self.padding)
        self.pos_embed = nn.Parameter(torch.zeros(1, self.grid[0] * self.grid[1], d))
        self.layers = nn.ModuleList()
        for _ in range(self.num_layers):
            blk = _Params()
            blk.ln1 = nn.LayerNorm(d, eps=self.eps)
            blk.attn = _Params()
            blk.attn.qkv = nn.Linear(d, 3 * d)
            blk.attn.proj = nn.Linear(d, d)
            blk.ln2 = nn.LayerNorm(d, eps=self.eps)
            blk.ffn = _Params()
            blk.ffn.layers = nn.Sequential(nn.Sequential(nn.Linear(d, ff), nn.GELU()), nn.Linear(ff, d))
            self.layers.append(blk)
        self.ln1 = nn.LayerNorm(d, eps=self.eps)
        self.init_weights()
        self._cache = EngineCache(dict(img_size=self.img_size, patch=patch_size, patch_pad=self.padding, embed_dim=d,
                                       depth=self.num_layers, heads=arch["num_heads"], ffn_dim=ff, deconv_channels=0,
                                       num_keypoints=0, ln_eps=self.eps), precision)
        for p in self.parameters():  # inference-only module
            p.requires_grad_(False)

    def init_weights(self):
        nn.init.trunc_normal_(self.pos_embed, std=0.02)
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.trunc_normal_(m.weight, std=0.02)
                nn.init.zeros_(m.bias)

    def engine_tensors(self, prefix: str = "backbone."):
        """MMPose checkpoint name -> tensor (the Parameter objects survive ``load_state_dict``
        and ``.to()``, so the dict is built once)."""
        if getattr(self, "_named", None) is None or self._named[0] != prefix:
            self._named = (prefix, {prefix + k: v for k, v in self.named_parameters()})
        return self._named[1]

    @torch.no_grad()
    def forward(self, x: torch.Tensor) -> Tuple[torch.Tensor]:
        """fp32 normalised RGB (B, 3, H, W) -> ``(featmap (B, C, gh, gw),)``."""
        eng = self._cache.get(self.engine_tensors(), x.shape[0], x.device)
        return (eng.backbone(x.float().contiguous()),)
