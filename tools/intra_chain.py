#!/usr/bin/env python
"""Micro-benchmark of the intra wavefront: rows of TUs that depend on their left neighbour only -> time per dependent step."""
import sys, os, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from openhevc_b200 import FrameEngine, worklist as W

def run(log2, rows, w=3840, h=256, bd=10):
    n = 1 << log2
    assert rows * n <= h
    per_row = w // n
    recs = np.zeros(rows * per_row, W.intra_dt)
    k = 0
    for x in range(per_row):          # level order: column by column
        for r in range(rows):
            recs[k] = (x * n, r * n, 0, log2, 1, (W.INF_LEFT if x else 0) | W.INF_FILTER, 0, 0, (0, 0), W.NO_RESID)
            k += 1
    blob = W.build_blob(w, h, 1, bd, 6, 0, intra=recs)
    eng = FrameEngine(w, h, 1, bd, n_slots=2)
    eng.set_profiling(True)
    best = 1e9
    for _ in range(5):
        eng.submit(blob); ms = eng.stage_ms()["intra"]; best = min(best, ms)
    eng.close()
    return best * 1e3 / per_row

for log2 in (2, 3, 4, 5):
    for rows in (1, 4):
        print(f"n={1<<log2:2d} rows={rows:3d}: {run(log2, rows):7.3f} us per dependent step")
