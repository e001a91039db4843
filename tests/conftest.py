import json
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))
GOLDEN = Path(__file__).resolve().parent / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


def pytest_collection_modifyitems(config, items):
    """A plain `pytest` on a machine without a GPU skips the gpu-marked tests instead of failing at the first one
    (the product itself still fails loudly without a GPU: tests/test_host_cpu.py checks that)."""
    try:
        import torch

        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device: gpu-marked tests run on the B200 box (pytest -m gpu)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_graph():
    return np.load(GOLDEN / "graph_small.npz")


@pytest.fixture(scope="session")
def golden_add_wsi():
    return json.loads((GOLDEN / "add_wsi_ref.json").read_text())


@pytest.fixture(scope="session")
def known_answers():
    return json.loads((GOLDEN / "notebook_known_answers.json").read_text())


@pytest.fixture(scope="session")
def engine():
    """The process-wide libpathgraph engine on cuda:0 (GPU tests only; fails loudly without a GPU)."""
    from path_gene_multimodal_b200.engine import get_engine

    return get_engine(0)


def frames_from_golden(g):
    """(nuc_df, tiles_df, expected_df) rebuilt from tests/golden/add_wsi_ref.json with exact floats."""
    import pandas as pd

    def frame(d):
        return pd.DataFrame(d["data"], columns=d["columns"], index=d["index"])

    nuc, tiles, out = frame(g["nuc_df"]), frame(g["tiles_df"]), frame(g["out"])
    unhex = float.fromhex
    nuc["polygon"] = [None if p is None else [[unhex(x), unhex(y)] for x, y in p] for p in g["polygon_hex"]]
    nuc["centroid"] = [[unhex(a), unhex(b)] for a, b in g["centroid_hex"]]
    for c, vals in g["out_hex"].items():
        out[c] = [unhex(v) for v in vals]
    out["wsi_polygon"] = [None if p is None else [[unhex(x), unhex(y)] for x, y in p] for p in g["wsi_polygon_hex"]]
    out["polygon"] = nuc["polygon"]
    out["centroid"] = nuc["centroid"]
    return nuc, tiles, out
