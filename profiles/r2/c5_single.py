"""The C5 stage of bench.py on one GPU (PG_C5_SORT=0/1 to compare the strip's row order)."""
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from path_gene_multimodal_b200.engine import get_engine  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
out = bench.c5_stage(get_engine(0), dev, None, 0, 1, n=int(sys.argv[1]) if len(sys.argv) > 1 else 20_000_000)
print(json.dumps({k: out[k] for k in ("partition_ms", "radius_ms", "knn_union_ms", "total_ms", "bit_identical_to_single_gpu", "checksums")}))
