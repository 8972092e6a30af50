"""The tcgen05 GEMM (through pp_gemm) against an fp64 torch matmul, all precisions / epilogues."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from probpose_code_b200 import _lib  # noqa: E402

TOL = {_lib.PREC_FP16X3: 3e-6, _lib.PREC_BF16: 1.2e-2, _lib.PREC_FP16: 1.5e-3, _lib.PREC_FP32_SIMT: 2e-6}


def _ops():
    from probpose_code_b200 import ops
    return ops


def _operand_to_f64(buf, rows, k, prec):
    """Decode an operand buffer back to float64 (what the tensor cores will multiply)."""
    if prec == _lib.PREC_FP32_SIMT:
        return buf.view(torch.float32).view(rows, k).double()
    if prec == _lib.PREC_BF16:
        return buf.view(torch.bfloat16).view(rows, k).double()
    if prec == _lib.PREC_FP16:
        return buf.view(torch.float16).view(rows, k).double()
    h = buf.view(torch.float16).view(rows, 2 * k).double()
    return (h[:, :k] + h[:, k:]) / 64.0  # FP16X3: hi + lo of the 64x-scaled value


def _rel(a, ref):
    return ((a.double() - ref).abs().max() / ref.abs().max()).item()


@pytest.mark.parametrize("prec", [_lib.PREC_FP16X3, _lib.PREC_BF16, _lib.PREC_FP16, _lib.PREC_FP32_SIMT])
@pytest.mark.parametrize("m,n,k", [(256, 384, 384), (192, 1152, 384), (1000, 1536, 384), (640, 384, 1536),
                                   (128, 256, 1024), (300, 64, 128), (77, 17, 256)])
def test_plain_gemm(prec, m, n, k):
    g = torch.Generator(device="cuda").manual_seed(m + n + k)
    a = torch.randn(m, k, device="cuda", generator=g)
    w = torch.randn(n, k, device="cuda", generator=g) * 0.05
    ops = _ops()
    ao, wo = ops.to_operand(a, prec), ops.to_operand(w, prec)
    # operand round trip is itself part of the contract
    assert _rel(_operand_to_f64(ao, m, k, prec), a.double()) <= {0: 3e-7, 1: 4e-3, 2: 5e-4, 3: 0}[prec]
    out = ops.gemm(ao, wo, m, n, k, prec)
    ref = a.double() @ w.double().t()
    assert _rel(out, ref) <= TOL[prec], f"rel err {_rel(out, ref)}"
    if prec != _lib.PREC_FP32_SIMT:  # against exactly what the tensor cores were given: accumulation error only
        ref_q = _operand_to_f64(ao, m, k, prec) @ _operand_to_f64(wo, n, k, prec).t()
        assert _rel(out, ref_q) <= 3e-6


@pytest.mark.parametrize("prec", [_lib.PREC_FP16X3, _lib.PREC_BF16, _lib.PREC_FP32_SIMT])
@pytest.mark.parametrize("tile_n", [0, 32, 64, 128, 192, 256])
def test_tile_widths(prec, tile_n):
    if prec == _lib.PREC_FP32_SIMT and tile_n:
        pytest.skip("tile_n only applies to the tensor-core kernel")
    m, n, k = 384, 768, 512
    g = torch.Generator(device="cuda").manual_seed(tile_n)
    a, w = torch.randn(m, k, device="cuda", generator=g), torch.randn(n, k, device="cuda", generator=g) * 0.05
    ops = _ops()
    out = ops.gemm(ops.to_operand(a, prec), ops.to_operand(w, prec), m, n, k, prec, tile_n=tile_n)
    assert _rel(out, a.double() @ w.double().t()) <= TOL[prec]


@pytest.mark.parametrize("prec", [_lib.PREC_FP16X3, _lib.PREC_BF16, _lib.PREC_FP32_SIMT])
def test_epilogues(prec):
    ops = _ops()
    m, n, k = 384, 384, 384
    g = torch.Generator(device="cuda").manual_seed(1)
    a, w = torch.randn(m, k, device="cuda", generator=g), torch.randn(n, k, device="cuda", generator=g) * 0.05
    scale, shift = torch.rand(n, device="cuda", generator=g) + 0.5, torch.randn(n, device="cuda", generator=g)
    res = torch.randn(m, n, device="cuda", generator=g)
    ao, wo = ops.to_operand(a, prec), ops.to_operand(w, prec)
    acc = a.double() @ w.double().t()
    tol = TOL[prec] * 4
    # bias + GELU
    out = ops.gemm(ao, wo, m, n, k, prec, shift=shift, act=_lib.ACT_GELU)
    assert _rel(out, torch.nn.functional.gelu(acc + shift.double())) <= tol
    # BN-style scale/shift + ReLU
    out = ops.gemm(ao, wo, m, n, k, prec, scale=scale, shift=shift, act=_lib.ACT_RELU)
    assert _rel(out, torch.relu(acc * scale.double() + shift.double())) <= tol
    # bias + residual (in place on the residual stream, like x += proj(...))
    out = ops.gemm(ao, wo, m, n, k, prec, shift=shift, residual=res, out=res.clone())
    assert _rel(out, acc + shift.double() + res.double()) <= tol
    # operand output feeds the next GEMM
    nxt = ops.gemm(ao, wo, m, n, k, prec, shift=shift, act=_lib.ACT_GELU, out_kind=_lib.OUT_OPERAND)
    back = _operand_to_f64(nxt, m, n, prec)
    assert _rel(back, torch.nn.functional.gelu(acc + shift.double())) <= max(tol, {0: 3e-6, 1: 8e-3, 3: 2e-6}[prec])
    # channel-major planes, N = 17 (final 1x1 conv)
    w17 = w[:17].contiguous()
    out = ops.gemm(ao, ops.to_operand(w17, prec), m, 17, k, prec, shift=shift[:17].contiguous(), out_kind=_lib.OUT_PLANES, plane=96)
    ref = (acc[:, :17] + shift[:17].double()).view(4, 96, 17).permute(0, 2, 1)
    assert out.shape == (4, 17, 96) and _rel(out, ref) <= tol
    # deconv phase scatter: rows (b, i, j) -> (b, 2i+py, 2j+px)
    hin, win = 8, 12  # m = 4 * 96
    full = torch.zeros(4 * m, n, device="cuda")
    for py in range(2):
        for px in range(2):
            ops.gemm(ao, wo, m, n, k, prec, out=full, up=(hin, win, py, px))
    v = full.view(4, 2 * hin, 2 * win, n)
    for py in range(2):
        for px in range(2):
            assert _rel(v[:, py::2, px::2].reshape(m, n), acc) <= tol


def test_argument_validation_needs_no_launch():
    ops = _ops()
    a = torch.zeros(128, 100, device="cuda")
    with pytest.raises(ValueError):
        ops.gemm(ops.to_operand(a, _lib.PREC_BF16), ops.to_operand(a, _lib.PREC_BF16), 128, 128, 100, _lib.PREC_BF16)


def _padded_rows(x_nhwc, shared=False):
    """(B, h, w, c) -> zero-padded rows, the tap-GEMM A operand: (B * (h + 2) * (w + 2), c) with a full border, or the
    shared-border layout (B * (h + 1) * (w + 1), c): zero row on top of every image, zero column on the right."""
    b, h, w, c = x_nhwc.shape
    if shared:
        p = torch.zeros(b, h + 1, w + 1, c, device=x_nhwc.device)
        p[:, 1:, :w] = x_nhwc
    else:
        p = torch.zeros(b, h + 2, w + 2, c, device=x_nhwc.device)
        p[:, 1:-1, 1:-1] = x_nhwc
    return p.reshape(-1, c).contiguous()


@pytest.mark.parametrize("shared", [False, True])
@pytest.mark.parametrize("prec", [_lib.PREC_FP16X3, _lib.PREC_BF16, _lib.PREC_FP32_SIMT])
def test_tap_gemm_is_a_3x3_convolution(prec, shared):
    """Implicit GEMM: 9 row-shifted reads of the padded map == Conv2d(3x3, padding=1), no im2col copy."""
    ops = _ops()
    b, h, w, cin, cout = 3, 16, 12, 128, 192
    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.randn(b, h, w, cin, device="cuda", generator=g)
    wt = torch.randn(cout, cin, 3, 3, device="cuda", generator=g) * 0.05
    bias = torch.randn(cout, device="cuda", generator=g)
    a = _padded_rows(x, shared)
    m = a.shape[0]
    pitch = w + 1 if shared else w + 2
    wp = wt.permute(0, 2, 3, 1).reshape(cout, 9 * cin).contiguous()  # column (ky*3+kx)*cin + ci
    taps = [(t // 3 - 1) * pitch + (t % 3 - 1) for t in range(9)]
    out = ops.gemm(ops.to_operand(a, prec), ops.to_operand(wp, prec), m, cout, 9 * cin, prec, shift=bias,
                   taps=taps, in_pad=(h, w), out_rows=b * h * w, shared_border=shared)
    ref = torch.nn.functional.conv2d(x.permute(0, 3, 1, 2).double(), wt.double(), bias.double(), padding=1)
    ref = ref.permute(0, 2, 3, 1).reshape(b * h * w, cout)
    assert _rel(out, ref) <= TOL[prec] * 4


@pytest.mark.parametrize("prec", [_lib.PREC_FP16X3, _lib.PREC_FP32_SIMT])
@pytest.mark.parametrize("shared", [False, True])
@pytest.mark.parametrize("out_pad", [False, True])
def test_tap_gemm_deconv_phases(prec, out_pad, shared):
    """Four sub-pixel phase GEMMs (2x2 taps each) == ConvTranspose2d(k4, s2, p1); with out_pad the
    result lands inside a bordered map whose border stays zero (ready for the next tap GEMM)."""
    ops = _ops()
    b, h, w, cin, cout = 2, 8, 6, 64, 128
    g = torch.Generator(device="cuda").manual_seed(9)
    x = torch.randn(b, h, w, cin, device="cuda", generator=g)
    wt = torch.randn(cin, cout, 4, 4, device="cuda", generator=g) * 0.05
    a = ops.to_operand(_padded_rows(x, shared), prec)
    pitch = w + 1 if shared else w + 2
    m = b * (h + 1) * (w + 1) if shared else b * (h + 2) * (w + 2)
    op = int(out_pad)
    ex = op * (1 if shared else 2)  # extra rows / columns of the output map
    H2, W2 = 2 * h + ex, 2 * w + ex
    full = torch.zeros(b * H2 * W2, cout, device="cuda")

    def tap(phase, t):  # (shift d, kernel index k) of tap t for output parity `phase`
        return ((0, 1), (-1, 3))[t] if phase == 0 else ((1, 0), (0, 2))[t]

    for py in range(2):
        for px in range(2):
            cols, shifts = [], []
            for t in range(4):
                dy, ky = tap(py, t >> 1)
                dx, kx = tap(px, t & 1)
                cols.append(wt[:, :, ky, kx].t())  # (cout, cin)
                shifts.append(dy * pitch + dx)
            wp = torch.cat(cols, dim=1).contiguous()
            ops.gemm(a, ops.to_operand(wp, prec), m, cout, 4 * cin, prec, act=_lib.ACT_RELU, out=full,
                     up=(h, w, py, px), taps=shifts, in_pad=(h, w), out_pad=out_pad, shared_border=shared)
    ref = torch.relu(torch.nn.functional.conv_transpose2d(x.permute(0, 3, 1, 2).double(), wt.double(), stride=2, padding=1))
    ref = ref.permute(0, 2, 3, 1)
    v = full.view(b, H2, W2, cout)
    inner = v[:, op:, :2 * w] if shared else v[:, op:H2 - op, op:W2 - op]
    assert _rel(inner, ref) <= TOL[prec] * 4
    if out_pad and shared:
        assert v[:, 0].abs().max() == 0 and v[:, :, -1].abs().max() == 0
    elif out_pad:
        assert v[:, 0].abs().max() == 0 and v[:, -1].abs().max() == 0 and v[:, :, 0].abs().max() == 0 and v[:, :, -1].abs().max() == 0


@pytest.mark.parametrize("prec", [_lib.PREC_FP16X3, _lib.PREC_BF16])
@pytest.mark.parametrize("m,n,k,tile_n", [(512, 384, 384, 128), (1000, 1152, 384, 192), (2048, 1536, 384, 256),
                                          (300, 384, 1536, 128), (4096, 256, 1024, 256), (257, 768, 128, 192)])
def test_cta_pair_gemm(prec, m, n, k, tile_n):
    """tcgen05 cta_group::2: two CTAs share one 256-row tile (each stages half of W); forced with
    cta_pair=2 so that small shapes exercise it too, incl. ragged M (TMA zero fill + row masks)."""
    g = torch.Generator(device="cuda").manual_seed(m + n + k)
    a = torch.randn(m, k, device="cuda", generator=g)
    w = torch.randn(n, k, device="cuda", generator=g) * 0.05
    shift = torch.randn(n, device="cuda", generator=g)
    res = torch.randn(m, n, device="cuda", generator=g)
    ops = _ops()
    ao, wo = ops.to_operand(a, prec), ops.to_operand(w, prec)
    ref = a.double() @ w.double().t() + shift.double()
    out = ops.gemm(ao, wo, m, n, k, prec, shift=shift, tile_n=tile_n, cta_pair=2)
    assert _rel(out, ref) <= TOL[prec] * 2
    single = ops.gemm(ao, wo, m, n, k, prec, shift=shift, tile_n=tile_n, cta_pair=1)
    assert _rel(out, single.double()) <= 1e-6  # same arithmetic, different tiling of M
    out = ops.gemm(ao, wo, m, n, k, prec, shift=shift, residual=res, out=res.clone(), tile_n=tile_n, cta_pair=2)
    assert _rel(out, ref + res.double()) <= TOL[prec] * 2
    nxt = ops.gemm(ao, wo, m, n, k, prec, shift=shift, act=_lib.ACT_RELU, out_kind=_lib.OUT_OPERAND, tile_n=tile_n, cta_pair=2)
    assert _rel(_operand_to_f64(nxt, m, n, prec), torch.relu(ref)) <= max(TOL[prec] * 2, {0: 3e-6, 1: 8e-3}[prec])
