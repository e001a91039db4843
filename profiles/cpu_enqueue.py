"""How long the host needs to enqueue one C2 step (6 kernels through ctypes) vs how long the GPU needs to run it."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
from path_gene_multimodal_b200 import synth
from path_gene_multimodal_b200.engine import get_engine, radius_cell
eng = get_engine(0); dev = torch.device('cuda', 0)
n = 1_000_000
xy, ty, side = synth.make_points(n, synth.SEEDS['C2'])
d_xy, d_ty = torch.from_numpy(xy).to(dev), torch.from_numpy(ty).to(dev)
bounds = (0.0, 0.0, float(side), float(side))
cell = radius_cell(50.0)
out = {}
def step():
    global out
    eng.grid_build(d_xy, d_ty, None, cell, bounds)
    out = eng.radius_graph(50.0, upper=True, want_dist32=True, want_edges=True, capacity=2_000_000, out=out)
for _ in range(5): step()
torch.cuda.synchronize()
# host enqueue time with an idle, never-blocking queue: few steps at a time
ts = []
for _ in range(20):
    torch.cuda.synchronize()
    t0 = time.perf_counter(); step(); ts.append(time.perf_counter() - t0)
print('host enqueue per step: median %.1f us, min %.1f us' % (np.median(ts) * 1e6, min(ts) * 1e6))
# GPU time, back to back without L2 flush
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record()
for _ in range(50): step()
e1.record(); torch.cuda.synchronize()
print('back-to-back (L2 warm) per step: %.1f us' % (e0.elapsed_time(e1) * 1e3 / 50))
