"""Per-kernel CUDA-event times (medians, us) of the kNN pipeline on the 1M-nuclei slide:
grid build + query + undirected union + i<j edge list + composition.  python profiles/knn_stages.py [k ...]"""
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch

from path_gene_multimodal_b200 import synth
from path_gene_multimodal_b200.engine import default_knn_cell, get_engine

eng = get_engine(0)
dev = torch.device("cuda", 0)
n = int(1_000_000)
xy, ty, side = synth.make_points(n, synth.SEEDS["C2"])
d_xy, d_ty = torch.from_numpy(xy).to(dev), torch.from_numpy(ty).to(dev)
bounds = (0.0, 0.0, float(side), float(side))
import os
if os.environ.get("PG_SORTED_INPUT") == "1":   # the same points with their rows in cell order (how much of the chain is row-order locality?)
    eng.grid_build(d_xy, d_ty, None, default_knn_cell(n, float(side) ** 2, 8), bounds)
    d_xy, d_ty, _ = eng.grid_export()
    print("input rows in cell order")
for k in [int(a) for a in sys.argv[1:]] or [8, 16]:
    cell = default_knn_cell(n, float(side) ** 2, k)

    def run():
        eng.grid_build(d_xy, d_ty, None, cell, bounds)
        kres = eng.knn(k, dist_dtype=torch.float32)
        u = eng.knn_union(kres["knn_idx"], kres["dist32"], types=d_ty, n_types=5, symmetric_dist=True)
        return u, u

    for _ in range(3):
        run()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        run()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / 5 * 1e3
    eng.profile(True)
    for _ in range(5):
        sym, up = run()
    recs = eng.profile_records()
    eng.profile(False)
    per = {}
    for name, ms in recs:
        per.setdefault(name, []).append(ms)
    tot = 0.0
    print(f"k={k}: E_und={up['edges'].shape[0]}  wall {wall:.3f} ms per pipeline (the host read of the two totals included)")
    for name, v in per.items():
        med = float(np.median(v)) * 1e3
        cnt = len(v) // 5
        tot += med * cnt
        print(f"   {name:32s} x{cnt}  {med:8.1f} us")
    print(f"   sum of kernels {tot:.0f} us")
