"""GPU: the two backward strategies of photo_bwd_kernel -- dense (3x3 box-sum gather over the whole tile) and sparse
(the few selected windows of an auto-masked tile scatter their adjoint, one thread per live pixel) -- must give the same
gradients.  The library option photo_sparse_max = 0 (tdl_set_option) forces the dense path everywhere; the default (and maximum) switches per tile at 128 selected
windows, 16 moves the switch point so that more tiles of the test images take the dense path.  Both are also checked against the CPU oracle
(mono/model/mono_fm/net.py:63-106 via autograd) by running the oracle comparison of tests/test_gpu_parity.py under each setting.

A third strategy sits above both: (image, scale) pairs with at most photo_list_max (default 4096) selected windows are
differentiated by photo_bwd_list_kernel from the work list the scoring kernel emitted, and photo_bwd_kernel skips them;
photo_list_max = -1 switches it off.  The tile-path tests below run with it off; the list tests compare it with the dense
tile path and with the oracle, on images that are listed, not listed, and mixed within one batch."""
import pytest
import torch

from golden_util import reference_noise, spec_from_meta
from gpu_util import pkg, rel_l2, run_cuda
from test_gpu_parity import _check, _synthetic_record

pytestmark = pytest.mark.gpu


def _grads(rec, noise):
    loss, outs, grads = run_cuda(rec, noise)
    return loss, {k: g for k, g in grads.items() if g is not None}


@pytest.mark.parametrize("frames,automask", [("waves", True), ("smooth", True), ("white", True), ("waves", False)])
def test_sparse_and_dense_backward_agree(frames, automask):
    rec = _synthetic_record("baseline", 2, 96, 160, 0, 5000 + len(frames), frames=frames)
    rec["meta"]["opt"]["automask"] = automask
    noise = reference_noise(spec_from_meta(rec["meta"]), rec["meta"])
    res = {}
    for tag, val in (("dense", 0), ("default", 128), ("sparse16", 16)):
        with pkg()._lib.options(photo_sparse_max=val, photo_list_max=-1):
            res[tag] = _grads(rec, noise)
    loss_d, grads_d = res["dense"]
    for tag in ("default", "sparse16"):
        loss_t, grads_t = res[tag]
        for k in loss_d:
            assert float(loss_t[k]) == float(loss_d[k]), (tag, k)          # the forward does not depend on the setting
        for k, g in grads_d.items():
            if float(g.abs().max()) == 0.0:
                assert float(grads_t[k].abs().max()) == 0.0, (tag, k)
                continue
            assert rel_l2(grads_t[k], g) < 2e-5, (tag, k, rel_l2(grads_t[k], g))


@pytest.mark.parametrize("setting", [0, 128])
def test_each_backward_strategy_matches_oracle(setting):
    rec = _synthetic_record("baseline", 1, 64, 96, 0, 5100, frames="waves")
    with pkg()._lib.options(photo_sparse_max=setting, photo_list_max=-1):
        _check(rec, f"baseline-sparse_max={setting}")


@pytest.mark.parametrize("env", [{"no_tma": 1}, {"fused_fwd": 1}, {"no_tma": 1, "photo_sparse_max": 0}])
def test_fallback_kernels_match_oracle(env):
    """The plain-load variants (no TMA: what a tensor that TMA cannot describe gets -- unaligned base, odd strides) and the fused forward kernel (what
    non-materialised warps get) are product code too: same oracle comparison as the default path."""
    rec = _synthetic_record("baseline", 1, 64, 96, 0, 5200, frames="waves")
    with pkg()._lib.options(**env):
        _check(rec, f"baseline-{env}")



def _selected_per_pair(rec, noise):
    """number of windows whose arg-min is a warped frame, per (scale, image) -- what the scoring kernel counts"""
    _, outs, _ = run_cuda(rec, noise)
    S = len(rec["meta"]["opt"]["frame_ids"]) - 1
    chan0 = S if rec["meta"]["opt"]["automask"] else 0
    return {s: (outs[("min_index", s)] >= chan0).flatten(1).sum(1).tolist() for s in range(4)}


@pytest.mark.parametrize("frames,B,H,W,fids,expect", [
    ("smooth", 2, 96, 160, (0, -1, 1), "listed"),            # auto-masked: a few windows per image
    ("white", 2, 96, 160, (0, -1, 1), "tiles"),               # white noise: 70 % of the windows selected
    ("smooth", 1, 96, 112, (0, -1), "listed"),                # one source frame, partial tiles
    ("smooth", 1, 64, 96, (0, -1, 1, "s"), "listed"),         # three source frames
    ("waves", 2, 96, 160, (0, -1, 1), "any"),                 # band-limited frames
    ("scene", 8, 64, 96, (0, -1, 1), "mixed"),                # 7 moving images (tiles) + 1 static image (list) in one batch
])
def test_list_backward_agrees_with_tile_backward(frames, B, H, W, fids, expect):
    rec = _synthetic_record("baseline", B, H, W, 0, 5300 + H, frames=frames, frame_ids=fids)
    noise = reference_noise(spec_from_meta(rec["meta"]), rec["meta"])
    counts = [c for per in _selected_per_pair(rec, noise).values() for c in per]
    cap = pkg()._lib.get_option("photo_list_max")          # default = the list capacity (4096)
    if expect == "listed":
        assert 0 < max(counts) <= cap, counts
    elif expect == "tiles":
        assert min(counts) > cap, counts
    elif expect == "mixed":
        assert min(counts) <= cap < max(counts), counts
    with pkg()._lib.options(photo_sparse_max=0, photo_list_max=-1):
        loss_d, grads_d = _grads(rec, noise)
    variants = {"list": dict(), "list+dense-tiles": dict(photo_sparse_max=0)}
    if expect != "tiles":
        small = max(1, sorted(counts)[len(counts) // 2])          # about half of the listed pairs fall back to the tiles
        variants["list_max=median"] = dict(photo_list_max=small)
    for tag, o in variants.items():
        with pkg()._lib.options(**o):
            loss_t, grads_t = _grads(rec, noise)
        for k in loss_d:
            assert float(loss_t[k]) == float(loss_d[k]), (tag, k)
        for k, g in grads_d.items():
            if float(g.abs().max()) == 0.0:
                assert float(grads_t[k].abs().max()) == 0.0, (tag, k)
                continue
            assert rel_l2(grads_t[k], g) < 2e-5, (tag, k, rel_l2(grads_t[k], g))


@pytest.mark.parametrize("frames", ["smooth", "white"])
def test_list_backward_matches_oracle(frames):
    rec = _synthetic_record("fm", 2, 96, 160, 8, 5400, frames=frames)
    _check(rec, f"fm-list-{frames}")


def test_list_backward_border_windows():
    """Selected windows on the image border (reflection padding: a tap outside the image is its mirror pixel) through the
    list path."""
    rec = _synthetic_record("baseline", 1, 64, 96, 0, 5500, frames="smooth")
    noise = reference_noise(spec_from_meta(rec["meta"]), rec["meta"])
    assert max(_selected_per_pair(rec, noise)[0]) <= pkg()._lib.get_option("photo_list_max")
    _, outs, _ = run_cuda(rec, noise)
    sel = outs[("min_index", 0)] >= 2
    if not bool(sel[:, 0, :].any() or sel[:, -1, :].any() or sel[:, :, 0].any() or sel[:, :, -1].any()):
        pytest.skip("no border window selected by this seed")
    _check(rec, "baseline-list-border")
