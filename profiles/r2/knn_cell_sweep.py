"""kNN grid + query time against the cell size (as a multiple of d_k = sqrt(k / (pi rho))): larger cells = more candidates
in the 3x3 select pass but fewer points retried by the ring pass. 1 M points, CUDA events, median of 10."""
import math
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
from path_gene_multimodal_b200 import synth  # noqa: E402
from path_gene_multimodal_b200.engine import get_engine  # noqa: E402

eng = get_engine(0)
dev = torch.device("cuda", 0)
n = 1_000_000
xy, ty, side = synth.make_points(n, synth.SEEDS["C2"])
d_xy, d_ty = torch.from_numpy(xy).to(dev), torch.from_numpy(ty).to(dev)
bounds = (0.0, 0.0, float(side), float(side))
rho = n / float(side) ** 2
for k in (8, 16, 5):
    row = []
    for f in (0.9, 1.0, 1.1, 1.15, 1.2, 1.3, 1.4, 1.5):
        cell = f * math.sqrt(k / (math.pi * rho))
        ms = []
        out = None
        for i in range(13):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            eng.grid_build(d_xy, d_ty, None, cell, bounds)
            out = eng.knn(k, dist_dtype=torch.float32)
            e1.record()
            torch.cuda.synchronize()
            if i >= 3:
                ms.append(e0.elapsed_time(e1))
        row.append((f, round(float(np.median(ms)) * 1e3, 1)))
    print(f"k={k}: " + "  ".join(f"{f}:{t}us" for f, t in row), flush=True)
