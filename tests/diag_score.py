"""Diagnostic (not a test): scoring kernel v1 (round-1 strip, bit-faithful order) vs v2 (micro-tile) vs the CPU oracle."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
from golden_util import load_case, reference_noise, run_restatement, spec_from_meta
from gpu_util import pkg, run_cuda
from test_gpu_parity import _synthetic_record

tdl = pkg()
cases = [("waves 1x192x640 (seed 1238)", _synthetic_record("baseline", 1, 192, 640, 0, 1238, frames="waves")),
         ("waves 2x192x640", _synthetic_record("baseline", 2, 192, 640, 0, 1234, frames="waves")),
         ("scene 8x192x640", _synthetic_record("baseline", 8, 192, 640, 0, 1234, frames="scene"))]
for name, rec in cases:
    for automask in (True, False):
        rec["meta"]["opt"]["automask"] = automask
        meta = rec["meta"]
        noise = reference_noise(spec_from_meta(meta), meta)
        ref_loss, ref_out, _ = run_restatement(rec)
        res = {}
        for v1 in (1, 0):
            with tdl._lib.options(photo_v1=v1):
                res[v1] = run_cuda(rec, noise)
        print("==", name, "automask", automask)
        for s in range(4):
            k = ("min_reconstruct_loss", s)
            r, a, b = float(ref_loss[k]), float(res[1][0][k]), float(res[0][0][k])
            ia, ib, ir = res[1][1][("min_index_photo", s)], res[0][1][("min_index_photo", s)], ref_out[("min_index_photo", s)]
            st = ref_out[("reproj_stack", s)]
            # where v2's arg-min differs from the oracle: gap of the oracle's values
            d = ib != ir
            gap = (st.gather(1, ib.unsqueeze(1)).squeeze(1) - st.min(1).values)[d]
            print(f"  s{s} ref {r:.9e} v1 rel {abs(a-r)/r:.2e} v2 rel {abs(b-r)/r:.2e} (v2-ref {b-r:+.2e}) flips v1 {int((ia!=ir).sum())} v2 {int(d.sum())} maxgap {float(gap.max()) if d.any() else 0:.1e}")
