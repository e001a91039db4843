"""kNN query time vs grid cell size (in units of the expected k-th neighbour distance), per kernel."""
import sys, math
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
from path_gene_multimodal_b200 import synth
from path_gene_multimodal_b200.engine import get_engine
eng = get_engine(0); dev = torch.device('cuda', 0)
n = 1_000_000
xy, ty, side = synth.make_points(n, synth.SEEDS['C2'])
d_xy, d_ty = torch.from_numpy(xy).to(dev), torch.from_numpy(ty).to(dev)
bounds = (0.0, 0.0, float(side), float(side))
rho = n / float(side) ** 2
fs = [float(a) for a in sys.argv[1:]] or [1.0, 1.15, 1.3, 1.5, 1.8, 2.2]
for k in (5, 8, 16):
    dk = math.sqrt(k / (math.pi * rho))
    ref = None
    for f in fs:
        cell = f * dk
        res = {}
        for it in range(3):
            eng.grid_build(d_xy, d_ty, None, cell, bounds)
            res = eng.knn(k, dist_dtype=torch.float32, out=res)
        torch.cuda.synchronize()
        eng.profile(True)
        for it in range(5):
            eng.grid_build(d_xy, d_ty, None, cell, bounds)
            res = eng.knn(k, dist_dtype=torch.float32, out=res)
        recs = eng.profile_records(); eng.profile(False)
        per = {}
        for name, ms in recs: per.setdefault(name, []).append(ms)
        same = True if ref is None else bool(torch.equal(ref, res['knn_idx']))
        if ref is None: ref = res['knn_idx'].clone()
        txt = ', '.join(f'{k_}: {np.median(v)*1e3:.0f}' for k_, v in per.items() if 'knn' in k_ and 'prepare' not in k_)
        print(f'k={k} cell={f:.2f} d_k ({cell:.1f} px): {txt} us  same={same}')
