"""Short C2 driver for ncu captures: 3 steps of grid build + radius graph on the 1M-nuclei slide."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from path_gene_multimodal_b200 import synth  # noqa: E402
from path_gene_multimodal_b200.engine import get_engine, radius_cell  # noqa: E402

dev = torch.device("cuda", 0)
eng = get_engine(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
xy, ty, side = synth.make_points(n, 1002)
d_xy, d_ty = torch.from_numpy(xy).to(dev), torch.from_numpy(ty).to(dev)
b = (0.0, 0.0, float(side), float(side))
eng.grid_build(d_xy, d_ty, None, radius_cell(50.0), b)
g0 = eng.radius_graph(50.0, upper=True, n_types=5, want_edges=True)
cap = int(int(g0["total"]) * 1.25) + 1024
o = None
for _ in range(3):
    eng.grid_build(d_xy, d_ty, None, radius_cell(50.0), b)
    o = eng.radius_graph(50.0, upper=True, n_types=5, want_dist32=True, want_edges=True, capacity=cap, out=o)
torch.cuda.synchronize()
eng.check_overflow()
print("ok", int(o["row_ptr"][-1]))
