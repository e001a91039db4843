"""Round-2 experiment 1 (baseline diagnostics on the r1 code): how the timing method and concurrency move the
C2 step, PCIe ceilings, and the N sweep. Writes JSON to gpurun_out/r2_exp1.json."""
import json
import statistics
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from path_gene_multimodal_b200 import _host, synth  # noqa: E402
from path_gene_multimodal_b200.engine import Engine, get_engine, radius_cell  # noqa: E402

dev = torch.device("cuda", 0)
R = 50.0
out = {}


def make(n, seed):
    xy, ty, side = synth.make_points(n, seed)
    return torch.from_numpy(xy).to(dev), torch.from_numpy(ty).to(dev), float(side)


def med(x):
    return statistics.median(x)


def step_fn(eng, d_xy, d_ty, side, cap, out_d):
    eng.grid_build(d_xy, d_ty, None, radius_cell(R), (0.0, 0.0, side, side))
    return eng.radius_graph(R, upper=True, n_types=5, want_dist32=True, want_edges=True, capacity=cap, out=out_d)


def time_steps(fn, pre, reps=20, warm=4):
    ms = []
    for i in range(reps + warm):
        pre()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(i); e1.record()
        torch.cuda.synchronize()
        if i >= warm:
            ms.append(e0.elapsed_time(e1))
    return {"median_ms": med(ms), "min_ms": min(ms), "mean_ms": sum(ms) / len(ms)}


eng = get_engine(0)
flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device=dev)
n = 1_000_000
d_xy, d_ty, side = make(n, 1002)
eng.grid_build(d_xy, d_ty, None, radius_cell(R), (0.0, 0.0, side, side))
g0 = eng.radius_graph(R, upper=True, n_types=5, want_edges=True)
e_und = int(g0["total"]); cap = int(e_und * 1.25) + 1024
keep = {}


def one(i):
    keep["o"] = step_fn(eng, d_xy, d_ty, side, cap, keep.get("o"))


res = {}
res["flush_zero"] = time_steps(one, lambda: flush.zero_())
res["flush_read"] = time_steps(one, lambda: flush.max())
res["no_flush_warm"] = time_steps(one, lambda: None)
# rotating inputs/outputs: 10 distinct slides (10 x (20 MB in + ~70 MB out) >> L2), no explicit flush
sets = [make(n, 2000 + s) for s in range(10)]
outs = [None] * 10


def rot(i):
    s = i % 10
    outs[s] = step_fn(eng, sets[s][0], sets[s][1], sets[s][2], cap + 200000, outs[s])


res["rotate10_no_flush"] = time_steps(rot, lambda: None, reps=30, warm=12)
out["c2_step_timing_method"] = res
print(json.dumps(res), flush=True)

# per-kernel profile under each pre-step treatment
prof = {}
for name, pre in (("flush_zero", lambda: flush.zero_()), ("flush_read", lambda: flush.max()), ("warm", lambda: None)):
    eng.profile(True)
    for i in range(10):
        pre(); one(i)
    recs = eng.profile_records(); eng.profile(False)
    per = {}
    for k, ms in recs:
        per.setdefault(k, []).append(ms)
    prof[name] = {k: round(med(v) * 1e3, 2) for k, v in per.items()}
out["c2_kernels_us"] = prof
print(json.dumps(prof), flush=True)

# back-to-back throughput: K steps enqueued without sync, one stream / two / four streams with own engines
def throughput(n_lanes, steps=40):
    engs = [eng] + [Engine(0) for _ in range(n_lanes - 1)]
    streams = [torch.cuda.Stream(dev) for _ in range(n_lanes)]
    lane_out = [None] * 10
    for s in range(10):
        with torch.cuda.stream(streams[s % n_lanes]):
            lane_out[s] = step_fn(engs[s % n_lanes], sets[s][0], sets[s][1], sets[s][2], cap + 200000, lane_out[s])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for st in streams:
        st.wait_stream(torch.cuda.current_stream())
    for i in range(steps):
        s = i % 10
        with torch.cuda.stream(streams[i % n_lanes]):
            lane_out[s] = step_fn(engs[i % n_lanes], sets[s][0], sets[s][1], sets[s][2], cap + 200000, lane_out[s])
    for st in streams:
        torch.cuda.current_stream().wait_stream(st)
    e1.record(); torch.cuda.synchronize()
    for e in engs[1:]:
        e.close()
    return e0.elapsed_time(e1) / steps


out["c2_back_to_back_ms_per_slide"] = {f"lanes{l}": throughput(l) for l in (1, 2, 4)}
print(json.dumps(out["c2_back_to_back_ms_per_slide"]), flush=True)

# PCIe ceilings, one process: pinned H2D / D2H, one cudaMemcpyAsync per buffer
pc = {}
for mb in (4, 20, 80):
    nb = mb * 1000 * 1000
    h = _host.pinned_empty((nb,), np.uint8)
    d = torch.empty(nb, dtype=torch.uint8, device=dev)
    ht = torch.from_numpy(h)
    for direction in ("h2d", "d2h"):
        ms = []
        for i in range(8):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            if direction == "h2d":
                d.copy_(ht, non_blocking=True)
            else:
                ht.copy_(d, non_blocking=True)
            e1.record(); torch.cuda.synchronize()
            if i >= 2:
                ms.append(e0.elapsed_time(e1))
        pc[f"{direction}_{mb}MB_GBs"] = nb / med(ms) / 1e6
out["pcie"] = pc
print(json.dumps(pc), flush=True)

# N sweep of the C2 step (same density)
sw = {}
for nn in (250_000, 1_000_000, 4_000_000, 16_000_000):
    xy, ty, sd = make(nn, 1002)
    eng.grid_build(xy, ty, None, radius_cell(R), (0.0, 0.0, sd, sd))
    g = eng.radius_graph(R, upper=True, n_types=5, want_edges=True)
    eu = int(g["total"]); cp = int(eu * 1.1) + 1024
    del g
    kk = {}

    def f(i):
        kk["o"] = step_fn(eng, xy, ty, sd, cp, kk.get("o"))

    t = time_steps(f, lambda: flush.zero_(), reps=8, warm=3)
    alg = 48 * nn + 16 * eu
    sw[str(nn)] = {"ms": t["median_ms"], "e_und": eu, "alg_MB": alg / 1e6, "GBs": alg / t["median_ms"] / 1e6, "frac": alg / t["median_ms"] / 1e6 / 6550.7}
    del xy, ty, kk
    torch.cuda.empty_cache()
out["n_sweep"] = sw
print(json.dumps(sw), flush=True)
Path(ROOT / "gpurun_out").mkdir(exist_ok=True)
(ROOT / "gpurun_out" / "r2_exp1.json").write_text(json.dumps(out, indent=1))
