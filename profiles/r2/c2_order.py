"""C2 step time as a function of the ORDER of the input rows (same points, same graph up to relabelling):
random (BASELINE's synthetic spec: every nucleus on a random tile), tile order (what the reference's producer emits:
tile after tile, nuclei in label order inside a tile), cell order (pg_grid_export of the radius grid)."""
import json
import statistics
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from path_gene_multimodal_b200 import synth  # noqa: E402
from path_gene_multimodal_b200.engine import get_engine, radius_cell  # noqa: E402

dev = torch.device("cuda", 0)
eng = get_engine(0)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
n = 1_000_000
xy, ty, side = synth.make_points(n, 1002)
b = (0.0, 0.0, float(side), float(side))
orders = {"random": np.arange(n)}
tile = (xy[:, 1] // 508).astype(np.int64) * (side // 508) + (xy[:, 0] // 508).astype(np.int64)
orders["tile"] = np.argsort(tile, kind="stable")
d_xy, d_ty = torch.from_numpy(xy).to(dev), torch.from_numpy(ty).to(dev)
eng.grid_build(d_xy, d_ty, None, radius_cell(50.0), b)
orders["cell"] = eng.grid_export()[2].cpu().numpy().astype(np.int64)
out = {}
for name, perm in orders.items():
    p_xy, p_ty = torch.from_numpy(np.ascontiguousarray(xy[perm])).to(dev), torch.from_numpy(np.ascontiguousarray(ty[perm])).to(dev)
    eng.grid_build(p_xy, p_ty, None, radius_cell(50.0), b)
    e = int(eng.radius_graph(50.0, upper=True, want_edges=True)["total"])
    cap = int(e * 1.25) + 1024
    keep = {}

    def step():
        eng.grid_build(p_xy, p_ty, None, radius_cell(50.0), b)
        keep["o"] = eng.radius_graph(50.0, upper=True, n_types=5, want_dist32=True, want_edges=True, capacity=cap, out=keep.get("o"))

    ms = []
    for i in range(24):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); step(); e1.record(); torch.cuda.synchronize()
        if i >= 4:
            ms.append(e0.elapsed_time(e1))
    eng.profile(True)
    for _ in range(8):
        flush.zero_(); step()
    per = {}
    for k, v in eng.profile_records():
        per.setdefault(k, []).append(v)
    eng.profile(False)
    out[name] = {"edges": e, "step_us": round(statistics.median(ms) * 1e3, 2),
                 "kernels_us": {k: round(statistics.median(v) * 1e3, 2) for k, v in per.items()}}
    print(name, json.dumps(out[name]), flush=True)
(ROOT / "gpurun_out" / "r2_c2_order.json").write_text(json.dumps(out, indent=1))
