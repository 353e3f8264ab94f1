"""Runs only the >L2 segment-SpMM roofline case of bench.py (for `ncu --set full`)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402
import bignn_b200 as B  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 8_000_000
peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json'))) if os.path.exists(
    os.path.join(ROOT, 'MEASURED_PEAKS.json')) else {}
B._lib.load()
print(json.dumps(bench.spmm_roofline(torch, B, rows, peaks, 'cuda:0')))
