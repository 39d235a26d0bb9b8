"""Diagnostic (not a test): where does the gradient error of a synthetic case concentrate?
usage: python tests/diag_grad.py kind B H W C seed"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
from golden_util import reference_noise, run_restatement, spec_from_meta
from gpu_util import rel_l2, run_cuda
from test_gpu_parity import _synthetic_record

kind, B, H, W, C, seed = sys.argv[1], *map(int, sys.argv[2:7])
frames = sys.argv[7] if len(sys.argv) > 7 else "smooth"
rec = _synthetic_record(kind, B, H, W, C, seed, frames=frames)
meta = rec["meta"]
noise = reference_noise(spec_from_meta(meta), meta)
loss, outs, grads = run_cuda(rec, noise)
ref_loss, ref_out, ref_leaves = run_restatement(rec)
forced = {}
for s in range(4):
    ours = outs[("min_index_photo", s)]
    d = ours != ref_out[("min_index_photo", s)]
    print("scale", s, "flips", int(d.sum()))
    if d.any():
        forced[("photo", s)] = ours
if kind == "fm":
    ours = outs[("min_index", 0)]
    d = ours != ref_out[("min_index", 0)]
    print("feat flips", int(d.sum()), [(int(i) // ours.shape[-1] % ours.shape[-2], int(i) % ours.shape[-1]) for i in d.flatten().nonzero().flatten()[:10]])
    if d.any():
        forced["feat"] = ours
ref_loss, ref_out, ref_leaves = run_restatement(rec, forced=forced)
sum(ref_loss.values()).backward()
for k, leaf in ref_leaves.items():
    r = leaf.grad
    if r is None:
        continue
    g = grads[k]
    e = (g - r).flatten()
    e2 = e.pow(2)
    top = e2.topk(min(10, e2.numel()))
    w = g.shape[-1]
    hh = g.shape[-2]
    srt = e2.double().sort(descending=True).values
    rest = (srt.sum() - srt.cumsum(0)).clamp_min(0).sqrt() / r.double().norm()
    need = int((rest > 1e-4).sum()) + 1 if float(rel_l2(g, r)) > 1e-4 else 0
    print(k, "cells to drop for 1e-4:", need, "of", e2.numel(), end="  ")
    print("rel", f"{rel_l2(g, r):.2e}", "top10 energy", f"{float(top.values.sum() / e2.sum()):.3f}",
          "at", [(int(i) // (w * hh), int(i) // w % hh, int(i) % w) for i in top.indices[:4]],
          "err", [f"{float(x):.2e}" for x in e[top.indices[:4]]], "ref", [f"{float(x):.2e}" for x in r.flatten()[top.indices[:4]]])
