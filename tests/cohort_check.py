"""C4 (BASELINE configs[3]) by hand (not collected by pytest; bench.py --gpus N runs the same stage as stages.c4_cohort):

    python tests/cohort_check.py                                   # 1 GPU
    torchrun --nproc-per-node 8 tests/cohort_check.py              # 8 GPUs, slides dealt out by nuclei count

PG_SLIDES (64) slides of PG_SLIDE_N (500000) nuclei, PG_LANES (4) in flight per GPU, PG_STAGING float | halfpx16 (polygon
vertices as float32 pixels or int16 half-pixels); prints the stage's JSON."""
import json
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench  # noqa: E402


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)
    out = bench.c4_stage(dev, local, dist, rank, world, n_slides=int(os.environ.get("PG_SLIDES", 64)),
                         n=int(os.environ.get("PG_SLIDE_N", 500_000)), lanes=int(os.environ.get("PG_LANES", 4)),
                         staging=os.environ.get("PG_STAGING", "float"))
    if rank == 0:
        print(json.dumps(out))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
