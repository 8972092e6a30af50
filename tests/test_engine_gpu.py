"""Parity of the CUDA engine (through the C ABI) with the CPU oracle on identical seeded weights
and crops: fp32 torch restatement of ViT-S + ProbMapHead, then the restated reference decode.

Tolerances (BASELINE.json north_star): keypoints within 1e-3 input px, presence probability (and
the other scalar branches) within 1e-4 - required of the PARITY precision modes (fp32_simt: CUDA
cores; fp16x3: tcgen05 tensor cores with split operands).  The throughput modes (fp16, bf16) are
checked against the looser bounds they actually achieve, stated below."""
import numpy as np
import pytest
import torch

from oracle import model_oracle
from probpose_code_b200 import synth

pytestmark = pytest.mark.gpu

KPT_TOL_PX = 1e-3
PROB_TOL = 1e-4
PARITY_MODES = ["fp32_simt", "fp16x3"]
# throughput modes: stage tolerances relative to the tensor's max magnitude
STAGE_REL = {"fp32_simt": 2e-5, "fp16x3": 2e-5, "fp16": 3e-3, "bf16": 2e-2}


@pytest.fixture(scope="module")
def setup():
    sd = synth.make_state_dict(seed=0)
    ref = model_oracle.ProbPoseRef().eval()
    ref.load_state_dict(sd)
    crops = synth.make_crops(3, seed=1)
    x = ref.preprocess(crops)
    with torch.no_grad():
        feat = ref.backbone(x)[0]
        logits = ref.head.heatmap_logits(feat)
        scal = torch.stack([m(feat).flatten(1) for m in (ref.head.probability_layers, ref.head.visibility_layers,
                                                         ref.head.oks_layers, ref.head.error_layers)], 1)
    return dict(sd=sd, ref=ref, crops=crops, x=x, feat=feat, logits=logits, scal=scal,
                rec_flip=ref.predict(x, flip_test=True), rec=ref.predict(x, flip_test=False))


def _engine(prec, sd, **kw):
    from probpose_code_b200.engine import Engine
    return Engine(precision=prec, max_batch=4, **kw).load_state_dict(sd)


def _rel(a, b):
    a, b = a.detach().cpu().double(), b.double()
    return ((a - b).abs().max() / b.abs().max()).item()


@pytest.mark.parametrize("prec", ["fp32_simt", "fp16x3", "fp16", "bf16"])
def test_stages(setup, prec):
    eng = _engine(prec, setup["sd"])
    assert _rel(eng.backbone(setup["x"].cuda()), setup["feat"]) <= STAGE_REL[prec]
    lg, sc = eng.head(setup["feat"].cuda().contiguous())
    assert _rel(lg, setup["logits"]) <= STAGE_REL[prec]
    assert _rel(sc, setup["scal"]) <= STAGE_REL[prec]


@pytest.mark.parametrize("prec", PARITY_MODES)
@pytest.mark.parametrize("flip", [True, False])
@pytest.mark.parametrize("src", ["u8", "f32"])
def test_end_to_end_parity(setup, prec, flip, src):
    eng = _engine(prec, setup["sd"])
    inp = setup["crops"].cuda() if src == "u8" else setup["x"].cuda()
    rec, hm = eng.infer(inp, flip_test=flip, return_heatmaps=True)
    rec = rec.cpu().numpy().astype(np.float64)
    ref = setup["rec_flip"] if flip else setup["rec"]
    kp = rec[..., :2] / [47, 63] * [192, 256]  # probmap.py:218
    assert np.abs(kp - ref[..., :2]).max() <= KPT_TOL_PX
    assert np.abs(rec[..., 3:] - ref[..., 3:]).max() <= PROB_TOL  # prob, vis, oks, err / diag
    assert np.abs(rec[..., 2] - ref[..., 2]).max() <= 1e-4  # heatmap value at the peak
    s = hm.flatten(2).sum(-1)
    assert torch.allclose(s, torch.ones_like(s), atol=1e-4)  # merged sparsemax maps still sum to 1
    assert eng.last_launch_count > 0


def test_throughput_modes_stay_close(setup):
    """fp16 / bf16: the scalar branches and almost all keypoints stay close; argmax flips on
    near-tied peaks are possible at this precision, so the keypoint bound
    is on the median, not the max."""
    for prec, ptol in (("fp16", 2e-3), ("bf16", 1.5e-2)):
        rec = _engine(prec, setup["sd"]).infer(setup["crops"].cuda()).cpu().numpy().astype(np.float64)
        assert np.abs(rec[..., 3:] - setup["rec_flip"][..., 3:]).max() <= ptol
        kp = rec[..., :2] / [47, 63] * [192, 256]
        assert np.median(np.abs(kp - setup["rec_flip"][..., :2])) <= 0.05


def test_batch_independence_and_empty_batch(setup):
    """Persons are independent units: a person's record does not depend on its batch-mates."""
    eng = _engine("fp16x3", setup["sd"])
    c = setup["crops"].cuda()
    full = eng.infer(c)
    single = eng.infer(c[1:2].contiguous())
    assert torch.equal(full[1:2], single)
    assert eng.infer(c[:0].contiguous()).shape == (0, 17, 7)
    with pytest.raises(ValueError):
        eng.infer(torch.zeros(9, 3, 256, 192, dtype=torch.uint8, device="cuda"))  # > max_batch
    with pytest.raises(ValueError):
        eng.infer(torch.zeros(1, 3, 128, 192, dtype=torch.uint8, device="cuda"))


def test_weight_loading_errors(setup):
    from probpose_code_b200 import _lib
    from probpose_code_b200.engine import Engine
    eng = Engine(precision="fp16x3", max_batch=1)
    with pytest.raises(_lib.PPError):  # nothing loaded
        eng.infer(torch.zeros(1, 3, 256, 192, dtype=torch.uint8, device="cuda"))
    sd = dict(setup["sd"])
    sd.pop("head.final_layer.bias")
    with pytest.raises(_lib.PPError, match="head.final_layer.bias"):
        eng.load_state_dict(sd)
    with pytest.raises(ValueError):
        eng.load_state_dict({"backbone.nope": torch.zeros(3)})
    with pytest.raises(ValueError):
        eng.load_state_dict({"backbone.ln1.weight": torch.zeros(3)})  # wrong size
    # backbone-only and head-only engines
    bb = Engine(precision="fp16x3", max_batch=2, deconv_channels=0).load_state_dict(setup["sd"], prefixes=("backbone.",))
    assert _rel(bb.backbone(setup["x"][:2].cuda().contiguous()), setup["feat"][:2]) <= STAGE_REL["fp16x3"]
    with pytest.raises(_lib.PPError):
        bb.head(setup["feat"].cuda().contiguous())


def test_vit_base_backbone(setup):
    """BASELINE config 5: ViT-B backbone (D 768, 12 heads of 64, FFN 3072)."""
    from probpose_code_b200.engine import Engine
    sd = synth.make_state_dict(seed=2, arch=synth.VIT_BASE)
    ref = model_oracle.VisionTransformerRef(**synth.VIT_BASE).eval()
    ref.load_state_dict({k[len("backbone."):]: v for k, v in sd.items() if k.startswith("backbone.")})
    x = setup["x"][:2]
    with torch.no_grad():
        feat = ref(x)[0]
    eng = Engine(precision="fp16x3", max_batch=2, embed_dim=768, heads=12, ffn_dim=3072, deconv_channels=0)
    eng.load_state_dict(sd, prefixes=("backbone.",))
    assert _rel(eng.backbone(x.cuda().contiguous()), feat) <= STAGE_REL["fp16x3"]


@pytest.mark.parametrize("flip", [True, False])
def test_batch_invariance(setup, flip):
    """Persons are independent units: a person's record must not depend on who else is in the batch, bit
    for bit - the property the per-rank sharding relies on.  Also exercises single-crop calls (192 rows = 1.5 GEMM
    tiles, attention tile 1 zero-filled beyond the last image)."""
    from probpose_code_b200 import synth as s
    eng = _engine("fp16x3", setup["sd"], )
    crops = s.make_crops(4, seed=21).cuda()
    full = eng.infer(crops, flip_test=flip).cpu()
    for lo, hi in ((0, 1), (1, 3), (3, 4)):
        part = eng.infer(crops[lo:hi].contiguous(), flip_test=flip).cpu()
        assert torch.equal(part, full[lo:hi]), f"records of crops[{lo}:{hi}] depend on the batch composition"


@pytest.mark.parametrize("flip", [True, False])
def test_graph_replay_is_bit_identical(setup, flip):
    """pp_engine_set_graph: the captured middle of pp_engine_infer (first call plain, second captured, later calls
    one cudaGraphLaunch) must give the records of plain launches bit for bit, on fresh inputs at fresh addresses,
    interleaved with another shape, and on a non-default stream."""
    from probpose_code_b200 import synth as s
    eng = _engine("fp16x3", setup["sd"])
    batches = [s.make_crops(3, seed=40 + i).cuda() for i in range(5)]
    other = s.make_crops(2, seed=50).cuda()
    eng.set_graph(0)
    plain = [eng.infer(c, flip_test=flip).cpu() for c in batches]
    plain_other = eng.infer(other, flip_test=flip).cpu()
    launches = eng.last_launch_count
    assert eng.graph_replay_count == 0
    eng.set_graph(-1)
    for i, c in enumerate(batches):
        got = eng.infer(c.clone(), flip_test=flip).cpu()
        assert torch.equal(got, plain[i]), f"call {i} (graph path) differs from plain launches"
        assert eng.last_launch_count == launches
        if i == 1:
            assert torch.equal(eng.infer(other, flip_test=flip).cpu(), plain_other)
    assert eng.graph_replay_count == len(batches) - 1  # call 0 plain, call 1 captured + replayed, 2.. replayed
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        got = eng.infer(batches[0], flip_test=flip)
    st.synchronize()
    assert torch.equal(got.cpu(), plain[0])
    eng.profile_begin()  # profiling brackets every launch with events: always plain launches
    n = eng.graph_replay_count
    got = eng.infer(batches[2], flip_test=flip).cpu()
    prof = eng.profile_end()
    assert eng.graph_replay_count == n and prof["gemm"]["launches"] > 0
    assert torch.equal(got, plain[2])


def test_parity_at_a_batch_that_fills_the_gpu(setup):
    """BASELINE config 2 shape in miniature: 32 crops with flip-TTA = 64 images = 768 attention units (every persistent
    attention CTA walks several), 48 GEMM row tiles per CTA pair, decode of 544 maps.  Parity against the oracle on every
    person, and records bit-identical to the same crops sent in batches of 4 (persons are independent units)."""
    from probpose_code_b200.engine import Engine
    n = 32
    eng = Engine(precision="fp16x3", max_batch=n).load_state_dict(setup["sd"])
    crops = synth.make_crops(n, seed=77)
    rec = eng.infer(crops.cuda(), flip_test=True).cpu()
    ref = setup["ref"].predict(setup["ref"].preprocess(crops), flip_test=True)
    r = rec.numpy().astype(np.float64)
    kp = r[..., :2] / [47, 63] * [192, 256]
    assert np.abs(kp - ref[..., :2]).max() <= KPT_TOL_PX
    assert np.abs(r[..., 3:] - ref[..., 3:]).max() <= PROB_TOL
    for lo in range(0, n, 4):
        part = eng.infer(crops[lo:lo + 4].cuda().contiguous(), flip_test=True).cpu()
        assert torch.equal(part, rec[lo:lo + 4]), f"records of crops[{lo}:{lo + 4}] depend on the batch composition"


def test_operand_range_guard(setup):
    """fp16x3 operands clamp at |a| > 1023.5: a checkpoint whose activations get there must be flagged, never decoded
    silently (pp_operand_overflow; the estimator raises after the first call of an engine)."""
    import probpose_code_b200.mmpose_api as api
    from probpose_code_b200 import _lib
    eng = _engine("fp16x3", setup["sd"])
    eng.operand_overflow()  # clear whatever earlier tests left behind
    eng.infer(setup["crops"].cuda())
    assert not eng.operand_overflow()
    sd = dict(setup["sd"])
    sd["backbone.layers.3.ln2.weight"] = sd["backbone.layers.3.ln2.weight"] * 3000.0  # LN output ~ 3000: beyond the range
    bad = _engine("fp16x3", sd)
    bad.infer(setup["crops"].cuda())
    assert bad.operand_overflow() and not bad.operand_overflow()  # sticky until read, cleared by the read
    model = api.MODELS.build(api.probpose_small_cfg(precision="fp16x3"))
    model.load_state_dict(sd)
    model.to("cuda:0")
    with pytest.raises(_lib.PPError, match="operand range"):
        model.test_step(dict(inputs=[c for c in setup["crops"]], data_samples=api.make_data_samples(3)))
    ok = _engine("fp32_simt", sd)  # the CUDA-core path has no operand range limit
    ok.infer(setup["crops"].cuda())
    assert not ok.operand_overflow()


def test_l2_residency_of_the_residual_stream_changes_no_bit(setup, monkeypatch):
    """The encoder's launches carry an access-policy window that keeps the fp32 residual stream persisting in L2
    (common.cuh L2Window; PP_NO_L2_PERSIST=1, read when an engine is created, turns it off): a cache policy, so the
    records must be identical - for an engine that uses the carve-out and for one too large for it."""
    from probpose_code_b200.engine import Engine
    for batch in (4, 80):
        crops = synth.make_crops(batch, seed=90 + batch).cuda()
        on = Engine(precision="fp16x3", max_batch=batch).load_state_dict(setup["sd"])
        monkeypatch.setenv("PP_NO_L2_PERSIST", "1")
        off = Engine(precision="fp16x3", max_batch=batch).load_state_dict(setup["sd"])
        monkeypatch.delenv("PP_NO_L2_PERSIST")
        for flip in (True, False):
            assert torch.equal(on.infer(crops, flip_test=flip), off.infer(crops, flip_test=flip))


@pytest.mark.parametrize("head_kind", ["probmap", "heatmap"])
def test_grouped_deconvolution_phases_change_no_bit(setup, monkeypatch, head_kind):
    """The four sub-pixel phases of each deconvolution run as ONE grouped tcgen05 GEMM launch (gemm_tc.cu "Grouped
    launch"; PP_NO_GROUPED_GEMM=1, read when an engine is created, launches them one by one).  Same tiles, same
    arithmetic: logits and records must be identical, and the grouped engine launches six kernels fewer."""
    from probpose_code_b200.engine import Engine
    sd = setup["sd"] if head_kind == "probmap" else {k: v for k, v in synth.make_state_dict(seed=0).items()
                                                      if k.startswith("backbone.") or k.startswith("head.deconv") or k.startswith("head.final")}
    kw = dict(precision="fp16x3", max_batch=6, head_kind=head_kind)
    grouped = Engine(**kw).load_state_dict(sd)
    monkeypatch.setenv("PP_NO_GROUPED_GEMM", "1")
    single = Engine(**kw).load_state_dict(sd)
    monkeypatch.delenv("PP_NO_GROUPED_GEMM")
    for e in (grouped, single):
        e.set_graph(0)
    crops = synth.make_crops(5, seed=123).cuda()
    for flip in (True, False):
        a, b = grouped.infer(crops, flip_test=flip), single.infer(crops, flip_test=flip)
        assert torch.equal(a, b)
        assert single.last_launch_count - grouped.last_launch_count == 6
    if head_kind == "probmap":
        la, sa = grouped.head(setup["feat"].cuda().contiguous())
        lb, sb = single.head(setup["feat"].cuda().contiguous())
        assert torch.equal(la, lb) and torch.equal(sa, sb)
