"""GPU: option variants and size-independent properties of the fused loss (through the public mixin / C ABI).

Variants are checked against the CPU oracle like tests/test_gpu_parity.py; the properties run at the BASELINE
benchmark size (B=8, 192x640, C=64), where the oracle would take minutes."""
import pytest
import torch

from golden_util import reference_noise, run_restatement, spec_from_meta
from gpu_util import make_loss_net, pkg, rel_l2, run_cuda
from test_gpu_parity import _check, _synthetic_record

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("override", [dict(automask=False), dict(disp_norm=False),
                                      dict(automask=False, disp_norm=False, smoothness_weight=0.1, disparity_smoothness=0.1)])
def test_option_variants_match_oracle(override):
    rec = _synthetic_record("baseline", 1, 64, 96, 0, 2001, frames="waves")
    rec["meta"]["opt"].update(override)
    _check(rec, f"baseline-{override}")


def test_config4_shape_matches_oracle():
    """BASELINE config 4 resolution (320x1024), one image."""
    rec = _synthetic_record("baseline", 1, 320, 1024, 0, 2002, frames="smooth")
    _check(rec, "baseline-320x1024")


def test_align_corners_true_matches_oracle(monkeypatch):
    """torch 1.1 F.grid_sample convention (what the reference was written for)."""
    import golden_util
    rec = _synthetic_record("baseline", 1, 64, 96, 0, 2003, frames="waves")
    orig = golden_util.spec_from_meta

    def spec_ac(meta):
        s = orig(meta)
        s.align_corners = True
        return s
    monkeypatch.setattr(golden_util, "spec_from_meta", spec_ac)
    import gpu_util
    orig_net = gpu_util.make_loss_net

    def net_ac(opt, kind):
        n = orig_net(opt, kind)
        n.grid_sample_align_corners = True
        return n
    monkeypatch.setattr(gpu_util, "make_loss_net", net_ac)
    import test_gpu_parity
    monkeypatch.setattr(test_gpu_parity, "spec_from_meta", spec_ac)
    _check(rec, "baseline-align_corners")


def test_frozen_extractor_has_no_feature_grads():
    rec = _synthetic_record("fm", 1, 64, 96, 8, 2004, frames="waves")
    meta = rec["meta"]
    noise = reference_noise(spec_from_meta(meta), meta)
    net = make_loss_net(meta["opt"], "fm")
    dev = "cuda"
    inputs = {k: v.to(dev) for k, v in rec["inputs"].items()}
    leaves = {k: v.to(dev).clone() for k, v in rec["leaves"].items()}
    for k, v in leaves.items():
        if isinstance(k, tuple) and k[0] in ("disp", "cam_T_cam"):
            v.requires_grad_(True)
    outputs = {k: v for k, v in leaves.items() if isinstance(k, tuple) and k[0] in ("disp", "cam_T_cam")}
    src = {f: leaves[("src_feat", f)] for f in (-1, 1)}
    loss = net.compute_losses_fm(inputs, outputs, {s: {f: n.to(dev) for f, n in d.items()} for s, d in noise.items()},
                                 leaves["tgt_feat"], src)
    loss.total().backward()
    full, _, grads = run_cuda(rec, noise)
    for s in range(4):
        assert rel_l2(leaves[("disp", 0, s)].grad.cpu(), grads[("disp", 0, s)]) < 1e-6
    assert leaves["tgt_feat"].grad is None and src[1].grad is None
    assert abs(float(loss.total().detach()) - float(sum(full.values()))) < 1e-7


def test_generate_images_pred_standalone():
    """The per-scale reference method still works on its own (mono/model/mono_fm/net.py:157-170)."""
    rec = _synthetic_record("baseline", 1, 64, 96, 0, 2005, frames="smooth")
    net = make_loss_net(rec["meta"]["opt"], "baseline")
    inputs = {k: v.cuda() for k, v in rec["inputs"].items()}
    outputs = {k: v.cuda() for k, v in rec["leaves"].items()}
    outputs = net.generate_images_pred(inputs, outputs, 2)
    _, ref_out, _ = run_restatement(rec)
    for f in (-1, 1):
        assert ("color", f, 2) in outputs and ("color", f, 0) not in outputs
        assert rel_l2(outputs[("color", f, 2)].cpu(), ref_out[("color", f, 2)].detach()) < 1e-5
    # the warped images carry no autograd history: asking for them with inputs that require grad is an error, not a
    # silent no-op (a subclass building its own term on outputs[("color", f, s)] would otherwise train nothing) ...
    outputs[("disp", 0, 2)] = outputs[("disp", 0, 2)].detach().requires_grad_(True)
    with pytest.raises(RuntimeError, match="without autograd history"):
        net.generate_images_pred(inputs, outputs, 2)
    with torch.no_grad():                      # ... and fine when the caller says it only inspects them
        net.generate_images_pred(inputs, outputs, 2)


def test_projection_is_full_fp32_under_tf32_matmul():
    """Training scripts switch TF32 matmuls on for the networks; the 3x4 projection matrices must not inherit that."""
    rec = _synthetic_record("baseline", 1, 64, 96, 0, 2007, frames="waves")
    saved = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        _check(rec, "baseline-tf32-matmul-flag")
    finally:
        torch.backends.cuda.matmul.allow_tf32 = saved


def test_reference_noise_mode_reproduces_reference_rng_stream():
    """noise_mode='reference' draws torch.randn from the global CPU generator in the reference's order."""
    rec = _synthetic_record("baseline", 1, 64, 96, 0, 2006, frames="smooth")
    meta = rec["meta"]
    noise = reference_noise(spec_from_meta(meta), meta)
    explicit, _, _ = run_cuda(rec, noise)
    net = make_loss_net(meta["opt"], "baseline")
    net.noise_mode = "reference"
    inputs = {k: v.cuda() for k, v in rec["inputs"].items()}
    outputs = {k: v.cuda() for k, v in rec["leaves"].items()}
    torch.manual_seed(meta["seed"])
    loss = net.compute_losses_baseline(inputs, outputs)
    for k, v in explicit.items():
        assert float(loss[k]) == float(v), k


# ------------------------------------------------------------------------------------------------ properties, full size
@pytest.fixture(scope="module")
def full_size():
    tdl = pkg()
    B, H, W, C = 8, 192, 640, 64
    inputs, outputs, extras = tdl.synth.make_inputs(B, H, W, seed=77, feat_channels=C, with_noise=True)
    opt = dict(frame_ids=[0, -1, 1], imgs_per_gpu=B, height=H, width=W, scales=[0, 1, 2, 3], min_depth=0.1,
               max_depth=100.0, automask=True, disp_norm=True, perception_weight=1e-3, smoothness_weight=1e-3)
    return tdl, opt, inputs, outputs, extras


def _run(tdl, opt, inputs, outputs, extras, idx=None, scale=1.0):
    net = make_loss_net(opt, "fm")
    sel = (lambda t: t[idx]) if idx is not None else (lambda t: t)
    inp = {k: sel(v).cuda().contiguous() for k, v in inputs.items()}
    leaves = {k: sel(v).cuda().contiguous().requires_grad_(True) for k, v in outputs.items()}
    tgt = sel(extras["tgt_feat"]).cuda().contiguous().requires_grad_(True)
    src = {f: sel(v).cuda().contiguous().requires_grad_(True) for f, v in extras["src_feats"].items()}
    noise = {s: {f: sel(n).cuda().contiguous() for f, n in d.items()} for s, d in extras["noise"].items()}
    loss = net.compute_losses_fm(inp, dict(leaves), noise, tgt, src)
    (loss.total() * scale).backward()
    return loss, leaves, tgt, src


def test_full_size_determinism_linearity_and_batch_structure(full_size):
    tdl, opt, inputs, outputs, extras = full_size
    l1, g1, t1, s1 = _run(tdl, opt, inputs, outputs, extras)
    l2, g2, t2, s2 = _run(tdl, opt, inputs, outputs, extras, scale=3.0)
    for k in l1:                                    # same inputs -> same loss scalars (fp64 accumulation)
        assert float(l1[k]) == float(l2[k]), k
    for k in g1:                                    # gradients are linear in the upstream gradient
        assert rel_l2(g2[k].grad, 3.0 * g1[k].grad) < 2e-5, k      # fp32 atomics: summation order varies
    assert rel_l2(t2.grad, 3.0 * t1.grad) < 2e-5
    # per-image independence: the batch loss is the mean of the single-image losses, and an image's gradient
    # does not depend on its batch neighbours (up to the 1/B normalisation)
    B = opt["imgs_per_gpu"]
    o1 = dict(opt, imgs_per_gpu=1)
    acc = {k: 0.0 for k in l1}
    for b in (0, B - 1):
        lb, gb, tb, sb = _run(tdl, o1, inputs, outputs, extras, idx=slice(b, b + 1))
        for k in ("disp", 0, 0), ("disp", 0, 3), ("cam_T_cam", 0, 1):
            assert rel_l2(gb[k].grad[0] / B, g1[k].grad[b]) < 5e-5, (b, k)
        assert rel_l2(sb[-1].grad[0] / B, s1[-1].grad[b]) < 5e-5
    perm = torch.arange(B - 1, -1, -1)
    lp, gp, tp, sp = _run(tdl, opt, inputs, outputs, extras, idx=perm)
    for k in l1:                                    # permuting the batch leaves the (mean) losses unchanged ...
        assert abs(float(lp[k]) - float(l1[k])) <= 2e-7 * abs(float(l1[k])), k
    assert rel_l2(gp[("disp", 0, 1)].grad, g1[("disp", 0, 1)].grad[perm]) < 2e-5     # ... and permutes the gradients


def test_disparity_maps_at_factor_1_and_32_match_oracle():
    """The decoder's maps sit at H/2 .. H/16; the ABI allows every power of two up to 32 (include/tdl.h).  Full-resolution
    and H/32 maps exercise the single-pixel and whole-tile cells of the scoring kernel's target pyramid."""
    rec = _synthetic_record("baseline", 2, 64, 96, 0, 2011, frames="waves")
    g = torch.Generator().manual_seed(2011)
    smooth = lambda h, w: torch.sigmoid(torch.nn.functional.avg_pool2d(torch.randn(2, 1, h + 2, w + 2, generator=g), 3, 1))
    rec["leaves"][("disp", 0, 0)] = smooth(64, 96)          # factor 1
    rec["leaves"][("disp", 0, 3)] = smooth(2, 3)            # factor 32
    _check(rec, "baseline-fac1-fac32")
