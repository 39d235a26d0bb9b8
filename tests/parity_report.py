"""Prints the full parity report (no assertions): python tests/parity_report.py [case ...]"""
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
from golden_util import CASES, load_case, reference_noise, spec_from_meta
from gpu_util import rel_l2, run_cuda


def main():
    names = sys.argv[1:] or CASES
    for name in names:
        rec = load_case(name)
        noise = reference_noise(spec_from_meta(rec["meta"]), rec["meta"])
        loss, outs, grads = run_cuda(rec, noise)
        print("==", name)
        for k, v in rec["loss"].items():
            print("  loss", k, float(loss[k]), float(v), "rel", abs(float(loss[k]) - float(v)) / max(abs(float(v)), 1e-30))
        for k, v in rec["out"].items():
            g = outs[k]
            if v.dtype in (torch.uint8, torch.int64):
                d = g.to(torch.int64) != v.to(torch.int64)
                print("  out", k, "flips", int(d.sum()), "of", d.numel())
            else:
                print("  out", k, "rel", rel_l2(g, v), "maxabs", float((g - v).abs().max()))
        for k, v in rec["grad"].items():
            if v is None:
                continue
            g = grads[k]
            if g is None:
                print("  grad", k, "MISSING")
                continue
            d = (g - v).abs()
            idx = int(d.argmax())
            print("  grad", k, "rel", rel_l2(g, v), "maxabs", float(d.max()), "at", idx, "ref", float(v.flatten()[idx]),
                  "got", float(g.flatten()[idx]), "norm", float(v.norm()))


if __name__ == "__main__":
    main()
