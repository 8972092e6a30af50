"""Seeded random-init weights and synthetic crops for ProbPose (no checkpoints offline).

The ``state_dict`` uses the MMPose key layout (``backbone.*`` = mmpretrain
``VisionTransformer``, ``head.*`` = ``ProbMapHead``; SURVEY.md §5 checkpoint row) so
the same dict loads into the CUDA engine and any reference-shaped module.

The reference's default head init is ``Normal(std=0.001)``
(probmap_head.py:592-598), which makes every logit ~0 and the argmax a coin toss; the
bench/parity weights therefore use larger stated stds so heatmaps are peaky
(SURVEY.md §7 "Random-init weights make argmax ill-conditioned").
"""
from __future__ import annotations

import torch

VIT_SMALL = dict(embed_dims=384, num_layers=12, num_heads=12, feedforward_channels=1536)
VIT_BASE = dict(embed_dims=768, num_layers=12, num_heads=12, feedforward_channels=3072)


def _tn(gen, shape, std):
    t = torch.empty(shape)
    torch.nn.init.trunc_normal_(t, std=std, a=-2 * std, b=2 * std, generator=gen)
    return t


def make_state_dict(seed: int = 0, arch: dict = VIT_SMALL, head_std: float = 0.05, final_std: float = 0.03,
                    branch_std: float = 0.02, tokens: int = 192, keypoints: int = 17,
                    deconv_channels=(256, 256)) -> dict:
    """ViT: trunc_normal(0.02) weights / pos_embed, LN affine slightly perturbed;
    head: Normal(head_std) deconvs, Normal(final_std) final 1x1, Normal(branch_std)
    scalar-branch convs; BN running stats / affine randomised (eval-mode BN must not be
    an identity or folding bugs would hide)."""
    g = torch.Generator().manual_seed(seed)
    d, nl, ff = arch["embed_dims"], arch["num_layers"], arch["feedforward_channels"]
    sd = {}

    def norm(prefix, n):
        sd[prefix + ".weight"] = 1 + 0.1 * torch.randn(n, generator=g)
        sd[prefix + ".bias"] = 0.05 * torch.randn(n, generator=g)

    def lin(prefix, o, i, std=0.02):
        sd[prefix + ".weight"] = _tn(g, (o, i), std)
        sd[prefix + ".bias"] = 0.02 * torch.randn(o, generator=g)

    sd["backbone.patch_embed.projection.weight"] = _tn(g, (d, 3, 16, 16), 0.02)
    sd["backbone.patch_embed.projection.bias"] = 0.02 * torch.randn(d, generator=g)
    sd["backbone.pos_embed"] = _tn(g, (1, tokens, d), 0.02)
    for l in range(nl):
        p = f"backbone.layers.{l}"
        norm(p + ".ln1", d)
        lin(p + ".attn.qkv", 3 * d, d)
        lin(p + ".attn.proj", d, d)
        norm(p + ".ln2", d)
        lin(p + ".ffn.layers.0.0", ff, d)
        lin(p + ".ffn.layers.1", d, ff)
    norm("backbone.ln1", d)

    def bn(prefix, n):
        sd[prefix + ".weight"] = 1 + 0.1 * torch.randn(n, generator=g)
        sd[prefix + ".bias"] = 0.1 * torch.randn(n, generator=g)
        sd[prefix + ".running_mean"] = 0.1 * torch.randn(n, generator=g)
        sd[prefix + ".running_var"] = 0.5 + torch.rand(n, generator=g)
        sd[prefix + ".num_batches_tracked"] = torch.tensor(0)

    c = d
    for i, co in enumerate(deconv_channels):  # ConvTranspose2d weight is (Cin, Cout, 4, 4)
        sd[f"head.deconv_layers.{3 * i}.weight"] = head_std * torch.randn(c, co, 4, 4, generator=g)
        bn(f"head.deconv_layers.{3 * i + 1}", co)
        c = co
    sd["head.final_layer.weight"] = final_std * torch.randn(keypoints, c, 1, 1, generator=g)
    sd["head.final_layer.bias"] = 0.1 * torch.randn(keypoints, generator=g)
    for br in ("probability", "visibility", "oks", "error"):
        for j in range(3):
            sd[f"head.{br}_layers.{4 * j}.weight"] = branch_std * torch.randn(d, d, 3, 3, generator=g)
            sd[f"head.{br}_layers.{4 * j}.bias"] = 0.02 * torch.randn(d, generator=g)
            bn(f"head.{br}_layers.{4 * j + 1}", d)
        sd[f"head.{br}_layers.12.weight"] = 0.05 * torch.randn(keypoints, d, 1, 1, generator=g)
        sd[f"head.{br}_layers.12.bias"] = 0.1 * torch.randn(keypoints, generator=g)
    return sd


def make_crops(batch: int, seed: int = 0, height: int = 256, width: int = 192) -> torch.Tensor:
    """uint8 BGR crops (B, 3, H, W): smooth random blobs + noise, so that the ViT sees
    spatial structure (pure white noise gives near-identical tokens)."""
    g = torch.Generator().manual_seed(seed)
    low = torch.rand(batch, 3, height // 16, width // 16, generator=g)
    img = torch.nn.functional.interpolate(low, size=(height, width), mode="bilinear", align_corners=False)
    img = img * 200 + 55 * torch.rand(batch, 3, height, width, generator=g)
    return img.clamp(0, 255).to(torch.uint8)


def planted_logit_pair(batch: int, seed: int = 0, device="cpu", keypoints: int = 17, height: int = 64, width: int = 48,
                       flip_indices=(0, 2, 1, 4, 3, 6, 5, 8, 7, 10, 9, 12, 11, 14, 13, 16, 15)):
    """Trained-model-like heatmap logits ("planted peaks"): per map
    a * exp(-r^2 / 2 s^2) + N(0, 0.05), centre uniform over the map including the border, a in [2, 8],
    s in [1, 3]; plus the matching flipped-pass logits (mirror image of the partner keypoint's map, amplitude
    +-10 %, fresh noise), as flip-TTA sees them.  Returns (logits, logits_flipped_pass) on `device`."""
    g = torch.Generator().manual_seed(seed)
    cx = (torch.rand(batch, keypoints, 1, 1, generator=g) * width - 0.5).to(device)
    cy = (torch.rand(batch, keypoints, 1, 1, generator=g) * height - 0.5).to(device)
    a = (2 + 6 * torch.rand(batch, keypoints, 1, 1, generator=g)).to(device)
    s = (1 + 2 * torch.rand(batch, keypoints, 1, 1, generator=g)).to(device)
    scale = (0.9 + 0.2 * torch.rand(batch, keypoints, 1, 1, generator=g)).to(device)
    yy = torch.arange(height, device=device, dtype=torch.float32).view(1, 1, height, 1)
    xx = torch.arange(width, device=device, dtype=torch.float32).view(1, 1, 1, width)
    bump = a * torch.exp(-((xx - cx) ** 2 + (yy - cy) ** 2) / (2 * s * s))
    gd = torch.Generator(device=device).manual_seed(seed + 1) if str(device) != "cpu" else g
    z = bump + 0.05 * torch.randn(bump.shape, generator=gd, device=device)
    inv = torch.argsort(torch.tensor(flip_indices)).to(device)
    zf = (bump * scale)[:, inv].flip(-1) + 0.05 * torch.randn(bump.shape, generator=gd, device=device)
    return z.contiguous(), zf.contiguous()
