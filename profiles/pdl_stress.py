import sys, os
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from path_gene_multimodal_b200 import synth
from path_gene_multimodal_b200.engine import get_engine, radius_cell
eng = get_engine(0); dev = torch.device('cuda', 0)
n = 1_000_000
xy, ty, side = synth.make_points(n, synth.SEEDS['C2'])
d_xy, d_ty = torch.from_numpy(xy).to(dev), torch.from_numpy(ty).to(dev)
bounds = (0.0, 0.0, float(side), float(side))
eng.grid_build(d_xy, d_ty, None, radius_cell(50.0), bounds)
g0 = eng.radius_graph(50.0, upper=True, want_edges=True)
e = int(g0['total']); ref_edges = g0['edges'].clone(); ref_deg = g0['degree'].clone(); ref_nbr = g0['nbr_count'].clone()
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
out = {}
bad = 0
for it in range(30):
    flush.zero_()
    eng.grid_build(d_xy, d_ty, None, radius_cell(50.0), bounds)
    out = eng.radius_graph(50.0, upper=True, want_dist32=True, want_edges=True, capacity=int(e * 1.25) + 1024, out=out)
    if it % 3 == 2:
        tot = int(out['row_ptr'][-1])
        ok = tot == e and torch.equal(out['edges'][:e], ref_edges) and torch.equal(out['degree'], ref_deg) and torch.equal(out['nbr_count'], ref_nbr)
        if not ok:
            bad += 1
            print('iter', it, 'total', tot, 'expected', e, 'deg_eq', torch.equal(out['degree'], ref_deg), 'nbr_eq', torch.equal(out['nbr_count'], ref_nbr),
                  'deg_diff', int((out['degree'] != ref_deg).sum()))
print('PG_PDL_MASK', os.environ.get('PG_PDL_MASK'), 'bad', bad)
if bad:
    d = (out['degree'] - ref_deg).cpu().numpy()
    idx = np.nonzero(d)[0]
    print('n wrong', len(idx), 'diff hist', np.unique(d[idx], return_counts=True))
    info = eng.grid_info()
    cx = np.floor((xy[idx, 0] - info['x0']) / info['cell']).astype(int); cy = np.floor((xy[idx, 1] - info['y0']) / info['cell']).astype(int)
    print('cy%32 hist', np.bincount(cy % 32, minlength=32))
    print('cx range', cx.min(), cx.max(), 'strip hist', np.bincount(cy // 32)[:40])
    print('first rows', idx[:10], 'cx', cx[:10], 'cy', cy[:10])
