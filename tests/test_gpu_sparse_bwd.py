"""GPU: the two backward strategies of photo_bwd_kernel -- dense (3x3 box-sum gather over the whole tile) and sparse
(the few selected windows of an auto-masked tile scatter their adjoint, one thread per live pixel) -- must give the same
gradients.  The library option photo_sparse_max = 0 (tdl_set_option) forces the dense path everywhere; the default (and maximum) switches per tile at 128 selected
windows, 16 moves the switch point so that more tiles of the test images take the dense path.  Both are also checked against the CPU oracle
(mono/model/mono_fm/net.py:63-106 via autograd) by running the oracle comparison of tests/test_gpu_parity.py under each setting."""
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
        with pkg()._lib.options(photo_sparse_max=val):
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
    with pkg()._lib.options(photo_sparse_max=setting):
        _check(rec, f"baseline-sparse_max={setting}")


@pytest.mark.parametrize("env", [{"no_tma": 1}, {"fused_fwd": 1}, {"no_tma": 1, "photo_sparse_max": 0}])
def test_fallback_kernels_match_oracle(env):
    """The plain-load variants (no TMA: what a tensor that TMA cannot describe gets -- unaligned base, odd strides) and the fused forward kernel (what
    non-materialised warps get) are product code too: same oracle comparison as the default path."""
    rec = _synthetic_record("baseline", 1, 64, 96, 0, 5200, frames="waves")
    with pkg()._lib.options(**env):
        _check(rec, f"baseline-{env}")

