#!/usr/bin/env python
"""Profiling helper: executes the 9 distinct pictures of the bench stream `reps` times on one GPU with
device-resident work lists (for `ncu`: capture the last repetition) and prints per-picture stage times."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c3_4k_main10_ra")
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--only", type=int, default=-1, help="execute only this blob index")
    args = ap.parse_args()
    from openhevc_b200 import FrameEngine
    from openhevc_b200 import frame_parallel as FP
    from openhevc_b200.synth import smooth_frame
    wl = bench.WORKLOADS[args.workload]
    eng = FrameEngine(wl["width"], wl["height"], wl["cfi"], wl["bit_depth"], n_slots=FP.N_SLOTS, n_arenas=16)
    blobs, stats = bench.make_blobs(wl, out_alloc=eng.pinned)
    eng.upload_slot(FP.anchor_slot(-1), smooth_frame(wl["width"], wl["height"], wl["cfi"], wl["bit_depth"], 7))
    for b, blob in enumerate(blobs):
        eng.upload(blob, b)
    eng.sync()
    pics = {p.blob: p for g in range(FP.INTRA_PERIOD_GOPS) for p in FP.gop_pictures(g)}
    order = [FP.BLOB_ANCHOR_I, FP.BLOB_ANCHOR_P] + list(range(2, FP.N_BLOBS))
    if args.only >= 0:
        order = [args.only]
    eng.set_profiling(True)
    names = [n for n, _ in FP.blob_specs()]
    for rep in range(args.reps):
        for b in order:
            eng.execute(b, pics[b].cur_slot, pics[b].ref_slots)
            ms = eng.stage_ms()
            if rep == args.reps - 1:
                st = stats[b]
                print(json.dumps({"picture": names[b], "ms": {k: round(v, 4) for k, v in ms.items()}, "n_intra": st["n_intra"], "n_tu": st["n_tu"],
                                  "n_mc_tiles": st["n_mc_tiles"], "bytes": {k[6:]: st[k] for k in st if k.startswith("bytes_")}}))
    eng.sync()
    print("launches", eng.launch_count())
    eng.close()


if __name__ == "__main__":
    main()
