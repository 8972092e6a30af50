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


@pytest.mark.parametrize("prec_name", ["fp16x3", "fp16"])
@pytest.mark.parametrize("batch,heads,dh", [(40, 12, 32), (70, 12, 32), (20, 12, 64)])
def test_persistent_attention_many_units(prec_name, batch, heads, dh):
    """More (image, head) units than resident CTAs (2 x 148 at d_h 32, 148 at d_h 64): every CTA walks several units
    handed out by the atomic counter, with the next unit's Q / K / V requested while the current one is in its softmax.
    Checked against torch on every row, against a second launch bit for bit (the launch's last CTA re-arms the
    counter), and against the one-unit-per-CTA case on a sub-batch (a unit's result must not depend on which CTA ran
    it or what ran before it in the same shared memory / TMEM)."""
    from probpose_code_b200 import _lib, ops
    prec = _lib.PRECISIONS[prec_name]
    qkv = _case(batch, heads, dh, seed=batch + dh)
    qkv_op = ops.to_operand(qkv, prec)
    qkv_r = ops.from_operand(qkv_op, batch * 192, 3 * heads * dh, prec)
    out_op = ops.attention(qkv_op, batch, 192, heads, dh, prec, impl=2)
    out = ops.from_operand(out_op, batch * 192, heads * dh, prec)
    ref = _ref(qkv_r, batch, heads, dh)
    err = (out - ref).abs().max().item()
    assert err <= TOL[prec_name] * max(1.0, ref.abs().max().item()), f"{prec_name}: max abs err {err}"
    for _ in range(3):
        again = ops.attention(qkv_op, batch, 192, heads, dh, prec, impl=2)
        assert torch.equal(again, out_op), "a repeated launch differs (unit counter not re-armed?)"
    sub = 3  # 36 units: one per CTA
    row_bytes = qkv_op.numel() // (batch * 192)
    sub_op = ops.attention(qkv_op[: sub * 192 * row_bytes].clone(), sub, 192, heads, dh, prec, impl=2)
    out_row_bytes = out_op.numel() // (batch * 192)
    assert torch.equal(sub_op, out_op[: sub * 192 * out_row_bytes]), "a unit's result depends on the CTA schedule"


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
