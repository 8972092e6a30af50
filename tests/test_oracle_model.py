"""The model oracle (plain-torch restatement; PARITY UNPINNED by the reference's own tests) is
cross-checked against independent implementations available in this image."""
import numpy as np
import torch

from oracle import decode_oracle, model_oracle


def test_sparsemax_is_the_simplex_projection():
    """Independent check: sparsemax(z) = argmin_p ||p - z||^2 on the simplex, via bisection on tau."""
    rng = np.random.default_rng(0)
    for scale in (0.01, 1.0, 10.0):
        z = (rng.normal(0, scale, (5, 3072))).astype(np.float32)
        p = model_oracle.sparsemax(torch.from_numpy(z)).numpy()
        assert (p >= 0).all() and np.allclose(p.sum(-1), 1, atol=1e-5)
        zz = z.astype(np.float64) - z.max(-1, keepdims=True)
        lo, hi = np.full(5, -1.0 - 1e-9), np.zeros(5)
        for _ in range(80):
            mid = (lo + hi) / 2
            s = np.maximum(zz - mid[:, None], 0).sum(-1)
            lo, hi = np.where(s > 1, mid, lo), np.where(s > 1, hi, mid)
        ref = np.maximum(zz - ((lo + hi) / 2)[:, None], 0)
        assert np.abs(p - ref).max() < 2e-6
        # numpy twin used by the decode tests agrees with the torch one
        assert np.abs(decode_oracle.sparsemax_rows(z) - p).max() < 1e-6


def test_vit_block_matches_torchvision_encoder_block():
    """mmpretrain's pre-LN block == torchvision EncoderBlock under the key map of SURVEY 8c."""
    from torchvision.models.vision_transformer import EncoderBlock

    torch.manual_seed(0)
    blk = model_oracle._Block(384, 12, 1536).eval()
    for p in blk.parameters():
        torch.nn.init.normal_(p, std=0.05)
    tv = EncoderBlock(12, 384, 1536, 0.0, 0.0).eval()
    tv.load_state_dict({
        "ln_1.weight": blk.ln1.weight, "ln_1.bias": blk.ln1.bias,
        "self_attention.in_proj_weight": blk.attn.qkv.weight, "self_attention.in_proj_bias": blk.attn.qkv.bias,
        "self_attention.out_proj.weight": blk.attn.proj.weight, "self_attention.out_proj.bias": blk.attn.proj.bias,
        "ln_2.weight": blk.ln2.weight, "ln_2.bias": blk.ln2.bias,
        "mlp.0.weight": blk.ffn.layers[0][0].weight, "mlp.0.bias": blk.ffn.layers[0][0].bias,
        "mlp.3.weight": blk.ffn.layers[1].weight, "mlp.3.bias": blk.ffn.layers[1].bias,
    })
    x = torch.randn(2, 192, 384)
    with torch.no_grad():
        assert (blk(x) - tv(x)).abs().max() < 2e-5


def test_patch_embed_geometry_and_flip_merge():
    vit = model_oracle.VisionTransformerRef()
    assert vit.grid == (16, 12) and vit.pos_embed.shape == (1, 192, 384)
    # rows 254-255 / cols 190-191 of the input are never read (16x12 patches of 16 with pad 2)
    x = torch.randn(1, 3, 256, 192)
    y = x.clone()
    y[:, :, 254:] = 7.0
    y[:, :, :, 190:] = -3.0
    with torch.no_grad():
        assert torch.equal(vit.patch_embed.projection(x), vit.patch_embed.projection(y))
    p = np.random.default_rng(0).random((2, 17, 64, 48)).astype(np.float32)
    pf = np.random.default_rng(1).random((2, 17, 64, 48)).astype(np.float32)
    m = decode_oracle.tta_merge(p, pf, decode_oracle.COCO_FLIP_INDICES)
    assert m[1, 1, 5, 7] == np.float32(0.5) * (p[1, 1, 5, 7] + pf[1, 2, 5, 40])


def test_head_shapes_and_state_dict_names():
    from probpose_code_b200 import synth

    ref = model_oracle.ProbPoseRef().eval()
    sd = synth.make_state_dict(seed=3)
    assert set(ref.state_dict()) == set(sd)
    ref.load_state_dict(sd)
    with torch.no_grad():
        out = ref.head((torch.randn(2, 384, 16, 12),))
    assert out[0].shape == (2, 17, 64, 48) and all(o.shape == (2, 17, 1, 1) for o in out[1:])
    s = out[0].flatten(2).sum(-1)
    assert torch.allclose(s, torch.ones_like(s), atol=1e-4)  # sparsemax rows sum to 1
