"""CPU: oracle.restatement against the REAL reference classes on fresh seeded inputs (not the committed fixtures).

Runs wherever the reference is reachable: /root/reference in the build container, or the unmodified loss modules that
oracle/make_ref.py placed under the git-ignored oracle/_ref (what travels to the GPU box).  Skipped otherwise -- the
committed golden vectors (tests/test_oracle_golden.py) pin the restatement everywhere else.  Bit-exact: both sides run
the same torch ops in the same order on the same torch build."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))

from oracle import ref_loader  # noqa: E402

pytestmark = pytest.mark.skipif(not ref_loader.reference_available(), reason="reference sources not reachable")


@pytest.mark.parametrize("kind,B,H,W,frames,C,seed", [
    ("baseline", 1, 64, 96, "smooth", 0, 9001),
    ("fm", 1, 64, 96, "waves", 4, 9002),
    ("joint", 1, 96, 128, "smooth", 4, 9003),
    ("tripled", 1, 96, 128, "smooth", 4, 9004),
])
def test_restatement_is_bit_identical_to_the_reference(kind, B, H, W, frames, C, seed):
    import make_golden
    from golden_util import run_restatement
    rec = make_golden.run_reference(kind, B, H, W, frames, C, seed)          # executes the reference's compute_losses
    loss, outputs, leaves = run_restatement(rec)
    assert set(loss) == set(rec["loss"])
    for k, v in rec["loss"].items():
        torch.testing.assert_close(loss[k].detach(), v, rtol=0, atol=0, equal_nan=True, msg=str(k))
    for k, v in rec["out"].items():
        got = outputs[k].detach()
        if got.dtype == torch.int64:
            got = got.to(torch.uint8)
        torch.testing.assert_close(got, v, rtol=0, atol=0, msg=str(k))
    sum(v.mean() for v in loss.values()).backward()
    for k, g in rec["grad"].items():
        if g is not None:
            torch.testing.assert_close(leaves[k].grad, g, rtol=1e-6, atol=1e-9, msg=str(k))


def test_reference_root_is_unmodified():
    """oracle/_ref (when that is what is being used) matches the SHA-256 manifest written from /root/reference."""
    from oracle import make_ref
    if ref_loader.REF_KIND == "oracle/_ref":
        assert make_ref.verify()
    else:
        assert os.path.isfile(os.path.join(ref_loader.REF_ROOT, "mono", "model", "mono_fm", "layers.py"))
