"""Probe (GPU box and build container): are the reference's third-party geometry engines importable?
SURVEY 8(c): if shapely / skimage / libpysal import, golden vectors are generated from them; else parity
for a4 / a5 / f3 stays pinned on restatements only. Output is committed under profiles/."""
import importlib
import json
import platform
import sys

out = {"python": sys.version.split()[0], "platform": platform.platform()}
for name in ("shapely", "skimage", "libpysal", "geopandas", "torch_geometric", "scipy", "networkx", "pandas", "pyarrow", "cv2", "numba"):
    try:
        m = importlib.import_module(name)
        out[name] = getattr(m, "__version__", "importable")
    except Exception as e:  # noqa: BLE001
        out[name] = f"MISSING ({type(e).__name__}: {e})"
print(json.dumps(out, indent=1))
