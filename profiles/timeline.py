"""Kernel timeline of ONE CUDA-graph replay of the bench's loss step (CUPTI through torch.profiler; there is no nsys in
the image): start, duration and stream of every kernel, the idle gaps, and how much of the step each stream covers.
Not a bench number (the profiler adds overhead); it answers "what is on the critical path".

    python profiles/timeline.py smooth|scene [out.json] [lib options as name=value ...]
"""
import importlib
import json
import os
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch                                                        # noqa: E402
from torch.profiler import ProfilerActivity, profile               # noqa: E402

import bench                                                        # noqa: E402


def main():
    frames = sys.argv[1] if len(sys.argv) > 1 else "smooth"
    out = sys.argv[2] if len(sys.argv) > 2 and "=" not in sys.argv[2] else None
    opts = dict(a.split("=") for a in sys.argv[2:] if "=" in a)
    tdl = importlib.import_module(bench.PKG)
    for k, v in opts.items():
        tdl._lib.set_option(k, int(v))
    device = torch.device("cuda", 0)
    torch.cuda.set_device(device)
    torch.cuda.set_stream(torch.cuda.Stream(device))
    B, H, W = 8, 192, 640
    host = bench.make_host_workload(B, H, W, bench.rank_seed(0), frames)
    step = bench.DeviceStep(host, B, H, W, device, True)
    for _ in range(3):
        step.run_eager()
    step.capture()
    for _ in range(10):
        step.replay()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(4):
            step.replay()
        torch.cuda.synchronize()
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "trace.json")
        prof.export_chrome_trace(path)
        trace = json.load(open(path))
    ev = [e for e in trace["traceEvents"] if e.get("ph") == "X" and e.get("cat") in ("kernel", "gpu_memset", "gpu_memcpy")]
    ev.sort(key=lambda e: e["ts"])
    # split into replays: the first kernel of a replay is proj_fwd
    starts = [i for i, e in enumerate(ev) if "proj_fwd" in e["name"]]
    if len(starts) < 3:
        raise SystemExit(f"expected >= 3 replays in the trace, found {len(starts)}")
    a, b = starts[-2], starts[-1]                                    # the last complete replay
    one = ev[a:b]
    t0 = one[0]["ts"]
    period = ev[b]["ts"] - t0
    rows = []
    for e in one:
        name = e["name"].split("(")[0].replace("void ", "").replace("tdl::", "")
        rows.append({"name": name[:48], "stream": e["args"].get("stream"), "start_us": round(e["ts"] - t0, 2),
                     "dur_us": round(e["dur"], 2)})
    # busy intervals (union over streams) and idle gaps
    ivs = sorted((r["start_us"], r["start_us"] + r["dur_us"]) for r in rows)
    busy, cur_s, cur_e, gaps = 0.0, ivs[0][0], ivs[0][1], []
    for s, e in ivs[1:]:
        if s > cur_e:
            busy += cur_e - cur_s
            gaps.append((round(cur_e, 1), round(s - cur_e, 2)))
            cur_s, cur_e = s, e
        else:
            cur_e = max(cur_e, e)
    busy += cur_e - cur_s
    per_stream = {}
    for r in rows:
        per_stream[r["stream"]] = per_stream.get(r["stream"], 0.0) + r["dur_us"]
    summary = {"workload": frames, "options": opts, "replay_period_us": round(period, 1), "kernels": len(rows),
               "busy_union_us": round(busy, 1), "idle_us": round(period - busy, 1),
               "sum_kernel_us": round(sum(r["dur_us"] for r in rows), 1),
               "per_stream_us": {str(k): round(v, 1) for k, v in per_stream.items()},
               "largest_gaps(at_us, len_us)": sorted(gaps, key=lambda g: -g[1])[:8]}
    print(json.dumps(summary))
    for r in rows:
        print(f'{r["start_us"]:8.1f} {r["dur_us"]:7.1f}  s{r["stream"]}  {r["name"]}')
    if out:
        json.dump({"summary": summary, "kernels": rows}, open(out, "w"), indent=0)


if __name__ == "__main__":
    main()
