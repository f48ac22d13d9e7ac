"""DRAM traffic / throughput of the grouped DGEMM under different work-item orders (L2 rasterisation),
schedulers and TMA L2-promotion settings. Each variant: one timed bench run (TFLOP/s) and one ncu run
(dram bytes per launch). python scripts/raster_sweep.py out.json [N]"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out_path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/raster_sweep.json"
N = sys.argv[2] if len(sys.argv) > 2 else "16384"
VARIANTS = [
    {"TADEV_RASTER_S": "0"},
    {"TADEV_RASTER_S": "12"},
    {"TADEV_RASTER_S": "12", "TADEV_WAVE_SYNC": "1"},
    {"TADEV_RASTER_S": "12", "TADEV_WAVE_SYNC": "4"},
    {"TADEV_RASTER_S": "12", "TADEV_WAVE_SYNC": "16"},
    {"TADEV_RASTER_S": "0", "TADEV_WAVE_SYNC": "4"},
    {"TADEV_RASTER_S": "12", "TADEV_RASTER_ROWMAJOR": "1", "TADEV_WAVE_SYNC": "4"},
]
if os.environ.get("RASTER_VARIANTS"):
    VARIANTS = json.loads(os.environ["RASTER_VARIANTS"])
results = []
for v in VARIANTS:
    env = dict(os.environ)
    env.update(v)
    base = [sys.executable, os.path.join(ROOT, "bench.py"), "--n", N, "--no-e2e", "--no-cpu"]
    r = subprocess.run(base + ["--steps", "3", "--warmup", "3"], env=env, capture_output=True, text=True)
    tf = None
    for ln in r.stdout.splitlines():
        if ln.startswith("{"):
            tf = json.loads(ln)["value"]
    csvp = "/tmp/raster_ncu.csv"
    subprocess.run(["ncu", "--metrics", "gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum", "--clock-control", "none", "--print-units", "base",
                    "-k", "regex:gemm_grouped", "--csv", "--log-file", csvp] + base + ["--steps", "1", "--warmup", "1"],
                   env=env, capture_output=True, text=True)
    rd = wr = 0.0
    nl = 0
    hdr = None
    for row in csv.reader(open(csvp)):
        if len(row) > 5 and row[0] == "ID":
            hdr = row
            continue
        if hdr and len(row) == len(hdr):
            d = dict(zip(hdr, row))
            val = float(d["Metric Value"].replace(",", ""))
            if d["Metric Name"] == "dram__bytes_read.sum":
                rd += val
                nl += 1
            elif d["Metric Name"] == "dram__bytes_write.sum":
                wr += val
    rec = {"variant": v, "n": int(N), "tflops": tf, "launches": nl, "dram_read_gb_per_launch": rd / max(nl, 1) / 1e9,
           "dram_write_gb_per_launch": wr / max(nl, 1) / 1e9}
    print(json.dumps(rec), flush=True)
    results.append(rec)
json.dump(results, open(out_path, "w"), indent=1)
