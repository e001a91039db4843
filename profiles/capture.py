"""Short driver for `ncu --set full`: each hot-path kernel a few times at its bench shape (C2 radius pipeline
on 1M nuclei, C3 map+morphology on 2M x 32, kNN k=8 on 1M). Not a benchmark - nothing printed here is a result."""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from path_gene_multimodal_b200 import synth  # noqa: E402
from path_gene_multimodal_b200.engine import default_knn_cell, get_engine, radius_cell  # noqa: E402

eng = get_engine(0)
dev = torch.device("cuda", 0)
n = 1_000_000
xy, ty, side = synth.make_points(n, synth.SEEDS["C2"])
d_xy, d_ty = torch.from_numpy(xy).to(dev), torch.from_numpy(ty).to(dev)
bounds = (0.0, 0.0, float(side), float(side))
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
out = {}
for _ in range(reps):
    flush.zero_()
    eng.grid_build(d_xy, d_ty, None, radius_cell(50.0), bounds)
    out = eng.radius_graph(50.0, upper=True, want_dist32=True, want_edges=True, capacity=2_000_000, out=out)
kres = {}
for _ in range(reps):
    flush.zero_()
    eng.grid_build(d_xy, d_ty, None, default_knn_cell(n, float(side) ** 2, 8), bounds)
    kres = eng.knn(8, dist_dtype=torch.float32, out=kres)
sym = eng.knn_union(kres["knn_idx"], kres["dist32"], types=d_ty, n_types=5, symmetric_dist=True)
eng.csr_upper(sym["row_ptr"], sym["col"], sym["w"])          # the stand-alone passes, for any CSR
eng.compose_degree(sym["row_ptr"], sym["col"], d_ty, 5)
eng.clustering(sym["row_ptr"], sym["col"])
n3 = 2_000_000
off, pxy = synth.make_polygons(n3, synth.SEEDS["C3"], v_fixed=32)
rng = np.random.default_rng(3)
t = 73
tile_x = torch.from_numpy(((np.arange(t * t) % t) * 508).astype(np.int32)).to(dev)
tile_y = torch.from_numpy(((np.arange(t * t) // t) * 508).astype(np.int32)).to(dev)
nuc_tile = torch.from_numpy(rng.integers(0, t * t, size=n3).astype(np.int32)).to(dev)
cen = torch.from_numpy(rng.random((n3, 2)) * 508).to(dev)
bb = torch.from_numpy(rng.integers(0, 508, size=(n3, 4)).astype(np.int32)).to(dev)
d_off, d_p = torch.from_numpy(off).to(dev), torch.from_numpy(pxy).to(dev)
res = {}
for _ in range(reps):
    flush.zero_()
    res = eng.map_morph(d_off, d_p, nuc_tile, tile_x, tile_y, cen, bb, write_polygons=True, out=res)
# K10 / K12 (not on the bench path): node features of 1M nuclei x 10 columns, region properties of a 4096^2 map
feat = torch.rand((10, n), dtype=torch.float64, device=dev)
eng.node_features(feat, d_ty, torch.arange(1, 6, dtype=torch.int32, device=dev))
lab = (torch.arange(4096 * 4096, device=dev, dtype=torch.int32).reshape(4096, 4096) // 16 % 4096 // 16 +
       (torch.arange(4096, device=dev, dtype=torch.int32) // 16)[:, None] * 256 + 1).contiguous()
eng.raster_props(lab, 65536)
# round 2: strip partition / multi-range unpack / gid maps (C5 exchange steps), neighbour coordinates, count narrowing
gid = torch.arange(n, dtype=torch.int32, device=dev)
recs, totals = eng.strip_partition(d_xy, d_ty, gid, [float(side) * q / 8 for q in range(1, 8)])
xy_all = torch.empty((2 * n, 2), dtype=torch.float64, device=dev)
ty_all = torch.empty((2 * n,), dtype=torch.int32, device=dev)
gid_all = torch.empty((2 * n,), dtype=torch.int32, device=dev)
eng.halo_unpack_multi(recs, n, 0, 0, [(0.0, float(side) / 2), (float(side) / 2, float(side) + 1.0)], xy_all, ty_all, gid_all, n)
eng.gid_maps(gid, d_ty, n, n)
eng.knn_neighbor_coords(kres["knn_idx"], d_xy)
eng.narrow_counts(out["nbr_count"], torch.uint8)
# round 2 (late): halo exchange over peer memory - here with two local slabs standing in for the peers - and the
# half-pixel vertex staging of the cohort tables
cap, world = 1 << 18, 2
slabs = [torch.zeros(world * cap * 24 + 16, dtype=torch.uint8, device=dev) for _ in range(world)]
ptrs = torch.tensor([s_.data_ptr() for s_ in slabs], dtype=torch.int64, device=dev)
eng.halo_push(d_xy, d_ty, gid, float(side) * 0.1, float(side) * 0.9, ptrs.data_ptr(), world, 0, cap)
eng.halo_unpack_slab(slabs[1], world, 1, cap, [(-1.0, float(side) + 1.0)], xy_all, ty_all, gid_all, n)
q16 = torch.randint(-2000, 2000, (20_000_000, 2), dtype=torch.int16, device=dev)
eng.widen_halfpx(q16)
torch.cuda.synchronize()
eng.check_overflow()
print("capture ok", eng.launches)
