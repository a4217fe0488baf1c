#!/usr/bin/env python
"""Where does one dependent intra step spend its time?  Single chain of 4x4 TUs, per-TU %globaltimer stamps."""
import sys, os, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from openhevc_b200 import FrameEngine, worklist as W
log2 = int(sys.argv[1]) if len(sys.argv) > 1 else 2
w, h, bd, n = 3840, 256, 10, 1 << log2
per_row = w // n
recs = np.zeros(per_row, W.intra_dt)
for x in range(per_row):
    recs[x] = (x * n, 0, 0, log2, 1, (W.INF_LEFT if x else 0) | W.INF_FILTER, 0, 0, (0, 0), W.NO_RESID)
blob = W.build_blob(w, h, 1, bd, 6, 0, intra=recs)
eng = FrameEngine(w, h, 1, bd, n_slots=2)
tr = torch.zeros(per_row * 8, dtype=torch.int64, device="cuda")
eng.lib.b200_debug_set_intra_trace.argtypes = [C.c_void_p]
assert eng.lib.b200_debug_set_intra_trace(tr.data_ptr()) == 0
for _ in range(3):
    eng.submit(blob); eng.sync()
raw = tr.cpu().numpy().reshape(per_row, 8)
cyc = (raw[:, 7] - raw[:, 6]).astype(np.float64)
t = raw[:, :6].astype(np.float64)
t -= t[0, 0]
names = ["grab", "deps_ready", "gathered", "predicted", "fenced", "published"]
k = slice(100, per_row - 10)
print("per-step (published[k+1]-published[k]) mean ns:", np.diff(t[k, 5]).mean())
for a in range(1, 6):
    print(f"  {names[a-1]:>10s} -> {names[a]:<10s}: {np.mean(t[k, a] - t[k, a-1]):9.1f} ns")
print("  published[k] -> deps_ready[k+1]:", np.mean(t[101:per_row-9, 1] - t[100:per_row-10, 5]), "ns")
print("compute phase cycles (clock64):", cyc[100:-10].mean(), "-> effective SM MHz:", cyc[100:-10].mean() / np.mean(t[k, 3] - t[k, 2]) * 1e3)
print("globaltimer granularity sample:", np.unique(np.diff(np.sort(t[:, 5])))[:6])
eng.close()
