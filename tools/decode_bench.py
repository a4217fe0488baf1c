#!/usr/bin/env python
"""decode_bench.py — SURVEY.md §8d(i): the REAL decoder on a real Annex-B stream, end to end.

Times `oracle/_ref/decode_ref` (the unmodified reference, pure-C path) and `oracle/_ref/decode_b200` (the same decoder
with the six hook lines of INTEGRATION.md, pixel reconstruction on the GPU) on the same stream, with 1 thread and with
frame threads (`hevc -p N -f 1`), MD5 work off inside the timed loop.  Both arms include the CABAC parse on the host;
the hooked arm additionally includes the upload of every work list and the read-back of every picture.

    python tools/decode_bench.py [stream.hevc] [--threads 1,8] [--repeat 2]

Prints one JSON line.  Test / measurement infrastructure: it executes binaries under oracle/_ref (built here by
`__graft_entry__.build()`, shipped prebuilt to the GPU box).
"""
import argparse
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")


def run(binary, stream, threads, repeat, passes):
    best = None
    for _ in range(repeat):
        r = subprocess.run([os.path.join(REF, binary), stream, str(threads), "time", str(passes)], capture_output=True, text=True, timeout=1800)
        m = re.search(r"frames (\d+) time ([\d.]+) fps ([\d.]+) first_frame_s ([\d.]+) steady_fps ([\d.]+)", r.stdout)
        if r.returncode or not m:
            return {"error": (r.stderr or r.stdout)[-300:], "rc": r.returncode}
        frames, sec = int(m.group(1)), float(m.group(2))
        if best is None or sec < best["sec"]:
            best = {"frames": frames, "sec": sec, "fps": frames / sec, "first_frame_s": float(m.group(4)), "steady_fps": float(m.group(5))}
    return best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("stream", nargs="?", default=os.path.join(REF, "streams", "c3_4k_33.hevc"))
    ap.add_argument("--threads", default="1,%d" % min(len(os.sched_getaffinity(0)), 16))
    ap.add_argument("--repeat", type=int, default=2)
    ap.add_argument("--passes", type=int, default=3, help="decode the file this many times back to back; steady_fps excludes the first pass")
    ap.add_argument("--only", default="", help="ref | b200")
    a = ap.parse_args()
    if not os.path.exists(a.stream):
        a.stream = os.path.join(ROOT, "tests", "golden", "streams", "c3_3840x2160_10b_lowdelay.hevc")
    out = {"stream": os.path.basename(a.stream), "bytes": os.path.getsize(a.stream), "host_cores": os.cpu_count(), "passes": a.passes, "runs": []}
    for t in [int(x) for x in a.threads.split(",")]:
        for arm, binary in (("reference", "decode_ref"), ("b200", "decode_b200")):
            if a.only and not arm.startswith(a.only):
                continue
            if not os.path.exists(os.path.join(REF, binary)):
                continue
            r = run(binary, a.stream, t, a.repeat, a.passes)
            r.update(arm=arm, frame_threads=t)
            out["runs"].append(r)
    print(json.dumps(out))


if __name__ == "__main__":
    sys.exit(main())
