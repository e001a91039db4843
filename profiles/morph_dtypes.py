import numpy as np, torch, sys
sys.path.insert(0, "/root/repo")
from path_gene_multimodal_b200 import synth
from path_gene_multimodal_b200.engine import get_engine
eng = get_engine(0); dev = torch.device("cuda", 0)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
for name, n, vfix, dt in (("C3 f32", 2_000_000, 32, np.float32), ("C3 f64", 2_000_000, 32, np.float64), ("ragged f32", 2_000_000, None, np.float32), ("ragged f64", 2_000_000, None, np.float64)):
    off, xy = synth.make_polygons(n, 1003, v_fixed=vfix, dtype=dt)
    rng = np.random.default_rng(3)
    t = 73
    tile_x = torch.from_numpy(((np.arange(t * t) % t) * 508).astype(np.int32)).to(dev)
    tile_y = torch.from_numpy(((np.arange(t * t) // t) * 508).astype(np.int32)).to(dev)
    nuc_tile = torch.from_numpy(rng.integers(0, t * t, size=n).astype(np.int32)).to(dev)
    cen = torch.from_numpy(rng.random((n, 2)) * 508).to(dev)
    bb = torch.from_numpy(rng.integers(0, 508, size=(n, 4)).astype(np.int32)).to(dev)
    d_off, d_p = torch.from_numpy(off).to(dev), torch.from_numpy(xy).to(dev)
    res = {}
    ms = []
    for i in range(7):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); res = eng.map_morph(d_off, d_p, nuc_tile, tile_x, tile_y, cen, bb, write_polygons=True, out=res); e1.record()
        torch.cuda.synchronize()
        if i >= 2: ms.append(e0.elapsed_time(e1))
    m = xy.shape[0]
    b = 2 * xy.itemsize * 2 * m + 84 * n
    t_ms = float(np.median(ms))
    print(f"{name}: {t_ms*1e3:.0f} us, {b/1e9:.3f} GB -> {b/t_ms/1e6:.0f} GB/s = {b/t_ms/1e6/6550.7*100:.0f} % of peak")
    del d_p, res
