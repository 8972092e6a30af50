"""pp_attention (tcgen05 and mma.sync kernels) against plain PyTorch fp32 scaled_dot_product_attention,
the op mmpretrain's MultiheadAttention.forward calls (SURVEY.md 8c "Backbone")."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _case(batch, heads, dh, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    qkv = (torch.randn(batch * 192, 3 * heads * dh, generator=g) * scale).cuda()
    return qkv


def _ref(qkv, batch, heads, dh):
    q, k, v = qkv.view(batch, 192, 3, heads, dh).permute(2, 0, 3, 1, 4).double()
    s = (q @ k.transpose(-1, -2)) * dh ** -0.5
    o = torch.softmax(s, -1) @ v
    return o.transpose(1, 2).reshape(batch * 192, heads * dh).float()


# fp16x3 is the parity mode: fp32-grade; fp16 / bf16 carry their operand rounding
TOL = {"fp16x3": 3e-6, "fp16": 3e-3, "bf16": 2e-2}


@pytest.mark.parametrize("impl", [1, 2])
@pytest.mark.parametrize("prec_name", ["fp16x3", "fp16", "bf16"])
@pytest.mark.parametrize("batch,heads,dh", [(1, 12, 32), (5, 12, 32), (3, 12, 64), (2, 3, 32)])
def test_attention_matches_torch(impl, prec_name, batch, heads, dh):
    from probpose_code_b200 import _lib, ops
    prec = _lib.PRECISIONS[prec_name]
    qkv = _case(batch, heads, dh, seed=batch * 100 + heads + dh)
    qkv_op = ops.to_operand(qkv, prec)
    # the reference sees the same rounded operand values the kernel sees
    qkv_r = ops.from_operand(qkv_op, batch * 192, 3 * heads * dh, prec)
    out = ops.from_operand(ops.attention(qkv_op, batch, 192, heads, dh, prec, impl=impl), batch * 192, heads * dh, prec)
    ref = _ref(qkv_r, batch, heads, dh)
    err = (out - ref).abs().max().item()
    assert err <= TOL[prec_name] * max(1.0, ref.abs().max().item()), f"{prec_name} impl {impl}: max abs err {err}"


def test_attention_kernels_agree_on_peaky_scores():
    """Large score range (one dominant key per row): the softmax max-subtraction path."""
    from probpose_code_b200 import _lib, ops
    prec = _lib.PRECISIONS["fp16x3"]
    batch, heads, dh = 2, 12, 32
    qkv = _case(batch, heads, dh, seed=7, scale=4.0)
    qkv_op = ops.to_operand(qkv, prec)
    a = ops.from_operand(ops.attention(qkv_op, batch, 192, heads, dh, prec, impl=1), batch * 192, heads * dh, prec)
    b = ops.from_operand(ops.attention(qkv_op, batch, 192, heads, dh, prec, impl=2), batch * 192, heads * dh, prec)
    ref = _ref(ops.from_operand(qkv_op, batch * 192, 3 * heads * dh, prec), batch, heads, dh)
    assert (a - ref).abs().max().item() <= 2e-5 * ref.abs().max().item()
    assert (b - ref).abs().max().item() <= 2e-5 * ref.abs().max().item()


def test_attention_rejects_unsupported_shapes():
    from probpose_code_b200 import _lib, ops
    prec = _lib.PRECISIONS["fp16x3"]
    buf = torch.zeros(1 << 20, dtype=torch.uint8, device="cuda")
    with pytest.raises(Exception):
        ops.attention(buf, 1, 100, 12, 32, prec)
    with pytest.raises(Exception):
        ops.attention(buf, 1, 192, 12, 48, prec)
