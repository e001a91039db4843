"""Turn an `ncu --set full` report into the two committed artefacts:
    profiles/<tag>_ncu_full_summary.csv   one row per profiled launch, the metrics DESIGN.md cites
    profiles/ncu_traffic.json             DRAM bytes per launch per kernel (median over its launches), read by bench.py

    python profiles/summarize.py gpurun_out/prof_<tag>.ncu-rep <tag>
Needs `ncu` on PATH (reading a report needs no GPU)."""
import csv
import io
import json
import re
import statistics
import subprocess
import sys
from pathlib import Path

METRICS = [
    "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "gpu__time_duration.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
]
UNIT_TO_BYTES = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
UNIT_TO_US = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}


def short(name: str) -> str:
    name = re.sub(r"^void\s+", "", name)
    name = re.sub(r"<unnamed>::|\(anonymous namespace\)::", "", name)
    return re.sub(r"\(.*$", "", name).strip()


def main():
    rep, tag = Path(sys.argv[1]), sys.argv[2]
    root = Path(__file__).resolve().parent
    raw = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, body = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    keep = [m for m in METRICS if m in col]
    out = io.StringIO()
    w = csv.writer(out)
    w.writerow(["ID", "Kernel Name"] + keep)
    w.writerow(["", ""] + [units[col[m]] for m in keep])
    traffic = {}
    for r in body:
        name = short(r[col["Kernel Name"]])
        w.writerow([r[col["ID"]], name] + [r[col[m]] for m in keep])
        base = re.sub(r"<.*$", "", name)
        rd = float(r[col["dram__bytes_read.sum"]]) * UNIT_TO_BYTES[units[col["dram__bytes_read.sum"]]]
        wr = float(r[col["dram__bytes_write.sum"]]) * UNIT_TO_BYTES[units[col["dram__bytes_write.sum"]]]
        us = float(r[col["gpu__time_duration.sum"]]) * UNIT_TO_US[units[col["gpu__time_duration.sum"]]]
        traffic.setdefault(base, []).append((rd + wr, us))
    (root / f"{tag}_ncu_full_summary.csv").write_text(out.getvalue())
    js = {"source": f"profiles/{tag}_ncu_full_summary.csv (ncu --set full --clock-control none, cold caches: ncu flushes L2 "
                    "before every replay)",
          "unit": "bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum)", "kernels": {}}
    for k, v in traffic.items():
        js["kernels"][k] = {"traffic": int(statistics.median(t for t, _ in v)),
                            "duration_us": round(statistics.median(u for _, u in v), 2), "launches": len(v)}
    (root / "ncu_traffic.json").write_text(json.dumps(js, indent=1))
    for k, v in js["kernels"].items():
        print(f"{k:32s} {v['traffic'] / 1e6:10.2f} MB  {v['duration_us']:9.2f} us  x{v['launches']}")


if __name__ == "__main__":
    main()
