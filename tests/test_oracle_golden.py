"""CPU: oracle.restatement must reproduce the committed golden vectors, which were
produced by the real reference code (tests/golden/make_golden.py).  The
restatement uses the same torch ops in the same order, so the comparison is
bit-exact on this container's torch build (tolerance 0 for tensors; NaN == NaN)."""
import pytest
import torch

from golden_util import CASES, load_case, run_restatement


@pytest.mark.parametrize("name", CASES)
def test_restatement_matches_reference_golden(name):
    rec = load_case(name)
    loss, outputs, leaves = run_restatement(rec)
    assert set(loss) == set(rec["loss"])
    for k, v in rec["loss"].items():
        torch.testing.assert_close(loss[k].detach(), v, rtol=0, atol=0, equal_nan=True, msg=str(k))
    for k, v in rec["out"].items():
        got = outputs[k].detach()
        if got.dtype == torch.int64:
            got = got.to(torch.uint8)
        torch.testing.assert_close(got, v, rtol=0, atol=0, msg=str(k))
    sum(v.mean() for v in loss.values()).backward()     # a NaN term (empty difference map) still back-props its finite parts
    for k, g in rec["grad"].items():
        if g is None:
            assert leaves[k].grad is None or leaves[k].grad.abs().max() == 0
            continue
        # NaN loss terms (empty difference maps) carry no gradient in the reference either
        torch.testing.assert_close(leaves[k].grad, g, rtol=1e-6, atol=1e-9, msg=str(k))


def test_known_answers():
    """Closed forms: SSIM(x,x)=0, robust_l1(x,x)=1e-3,
    constant disparity => zero smoothness."""
    from oracle import restatement as R
    x = torch.rand(1, 3, 16, 24)
    assert R.ssim(x, x).abs().max() < 1e-6
    torch.testing.assert_close(R.robust_l1(x, x), torch.full_like(x, 1e-3))
    d = torch.full((1, 1, 8, 12), 0.3)
    assert R.smooth_loss(R.normalise_disp(d), torch.rand(1, 3, 16, 24)) == 0
