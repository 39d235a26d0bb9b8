"""Formats the JSON lines written by the GPU parity tests (TDL_PARITY_REPORT=file pytest -m gpu) as profiles/parity_r2.md."""
import json
import sys


def main(path):
    recs = [json.loads(l) for l in open(path)]
    print("# Parity report, round 2 -- CUDA path (through the C ABI) vs the CPU oracle / the reference's golden vectors\n")
    print("Produced on a B200 by `TDL_PARITY_REPORT=... python -m pytest tests -m gpu` (`profiles/evidence_r2.sh`), formatted by "
          "`tests/parity_report_md.py`.  Every row is one checked case; all numbers are relative errors (|a-b|/|b| for "
          "scalars, relative L2 for tensors) against `oracle/restatement.py`, which `tests/test_oracle_golden.py` and "
          "`tests/test_oracle_vs_reference.py` pin bit-exactly to the reference's own classes.\n")
    print("* **loss**: worst loss scalar (tolerance 1e-5; 3e-5 on fixtures below 49 k pixels; plus the near-tie allowance of "
          "`tests/test_gpu_parity.py`).  **images**: worst warped image / feature map (1e-5).")
    print("* **grad untrimmed / trimmed**: worst depth / feature gradient before and after discarding the K largest-error cells "
          "(K = 0.4 % of a tensor, 1 % on `scene`); **strict** cases (band-limited `waves` frames) must meet 1e-4 UNTRIMMED, "
          "up to the cells of a few isolated one-pixel events (`cells>tol`, counted).  **pose**: worst pose gradient.")
    print("* **ref-vs-ref floor**: the same gradient error of the REFERENCE run as eager PyTorch on the same GPU against its own "
          "CPU run (benchmark-configuration cases): the kernels sit on that floor.\n")
    print("| case | strict | live px | flips | loss | images | grad untrimmed | grad trimmed | cells>tol | pose | ref-vs-ref floor (untrimmed / trimmed) |")
    print("|---|---|---:|---:|---:|---:|---:|---:|---:|---:|---|")
    for d in recs:
        r = d["report"]
        g = lambda pre, suf="": [v for k, v in r.items() if k.startswith(pre) and (k.endswith(suf) if suf else True) and isinstance(v, (int, float))]
        loss = max(g("loss ") or [0])
        img = max(g("out ") or [0])
        grads = {k: v for k, v in r.items() if k.startswith("grad ") and "cam_T_cam" not in k}
        un = max([v for k, v in grads.items() if not k.endswith("trimmed") and "cells" not in k and "floor" not in k] or [0])
        tr = max([v for k, v in grads.items() if k.endswith(" trimmed") and "floor" not in k] or [0])
        cells = sum(int(v) for k, v in grads.items() if k.endswith("cells>tol"))
        pose = max([v for k, v in r.items() if k.startswith("grad ") and "cam_T_cam" in k] or [0])
        fl_un = max([v for k, v in grads.items() if k.endswith("floor")] or [0])
        fl_tr = max([v for k, v in grads.items() if k.endswith("floor trimmed")] or [0])
        flips = sum(int(v) for k, v in r.items() if k.startswith("flips "))
        live = r.get("live_frac")
        print(f"| {d['case']} | {'yes' if d['strict'] else ''} | {'' if live is None else f'{100 * live:.0f} %'} | {flips} | {loss:.1e} | {img:.1e} | "
              f"{un:.1e} | {tr:.1e} | {cells if d['strict'] else ''} | {pose:.1e} | {f'{fl_un:.1e} / {fl_tr:.1e}' if fl_un else ''} |")


if __name__ == "__main__":
    main(sys.argv[1])
