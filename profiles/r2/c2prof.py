"""C2 step timing + per-kernel CUDA-event profile (+ optional N sweep). Usage: c2prof.py [tag] [sweep]"""
import json
import statistics
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from path_gene_multimodal_b200 import synth  # noqa: E402
from path_gene_multimodal_b200.engine import get_engine, radius_cell  # noqa: E402

tag = sys.argv[1] if len(sys.argv) > 1 else "run"
sweep = len(sys.argv) > 2
dev = torch.device("cuda", 0)
R = 50.0
import os
CELL_DIV = float(os.environ.get("PG_CELL_DIV", "1"))   # cells of r / CELL_DIV: the walk then takes the general (2R+1)^2 block path
_rc = radius_cell
radius_cell = lambda r: _rc(r) / CELL_DIV  # noqa: E731
eng = get_engine(0)
flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device=dev)
out = {"tag": tag}


def run(n, reps=20):
    xy, ty, side = synth.make_points(n, 1002)
    d_xy, d_ty = torch.from_numpy(xy).to(dev), torch.from_numpy(ty).to(dev)
    b = (0.0, 0.0, float(side), float(side))
    eng.grid_build(d_xy, d_ty, None, radius_cell(R), b)
    g0 = eng.radius_graph(R, upper=True, n_types=5, want_edges=True)
    e_und = int(g0["total"]); cap = int(e_und * 1.25) + 1024
    keep = {}

    def step():
        eng.grid_build(d_xy, d_ty, None, radius_cell(R), b)
        keep["o"] = eng.radius_graph(R, upper=True, n_types=5, want_dist32=True, want_edges=True, capacity=cap, out=keep.get("o"))

    ms = []
    for i in range(reps + 4):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); step(); e1.record()
        torch.cuda.synchronize()
        if i >= 4:
            ms.append(e0.elapsed_time(e1))
    eng.check_overflow()
    assert int(keep["o"]["row_ptr"][-1]) == e_und
    eng.profile(True)
    for _ in range(10):
        flush.zero_(); step()
    recs = eng.profile_records(); eng.profile(False)
    per = {}
    for k, v in recs:
        per.setdefault(k, []).append(v)
    alg = 48 * n + 16 * e_und
    t = statistics.median(ms)
    return {"n": n, "e_und": e_und, "step_us": round(t * 1e3, 2), "min_us": round(min(ms) * 1e3, 2), "frac_of_6550.7": round(alg / t / 1e6 / 6550.7, 4),
            "kernels_us": {k: round(statistics.median(v) * 1e3, 2) for k, v in per.items()}}


out["c2_1M"] = run(1_000_000)
print(json.dumps(out["c2_1M"]), flush=True)
if sweep:
    for n in (250_000, 4_000_000, 16_000_000):
        out[f"c2_{n}"] = run(n, reps=8)
        print(json.dumps(out[f"c2_{n}"]), flush=True)
(ROOT / "gpurun_out" / f"r2_c2prof_{tag}.json").write_text(json.dumps(out, indent=1))
