"""Damaged streams against the whole host + device stack under AddressSanitizer, without a GPU.

Builds sanitizer versions of the shim (hevcdsp_init_b200.c), the recorder and the emulated device library (tests/emul/warp/),
pre-loads them in front of oracle/_ref/decode_b200, and decodes bit-flipped copies of committed streams with one thread, four
frame threads and (streams with entry points) four slice workers.  Any sanitizer report, non-zero exit or time-out is printed.
TEST INFRASTRUCTURE (needs /root/reference headers for the shim build).

    python tools/fuzz/corrupt_sweep.py [--seeds 5]
"""
import argparse
import os
import random
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REFD = os.path.join(ROOT, "oracle", "_ref")
OUT = "/tmp/b200_corrupt_sweep"
STREAMS = ["b_416x240_10b_weighted", "ra_416x240_8b", "cip_416x240_8b_lowdelay", "tiles_416x240_10b_nolf", "c444_416x240_8b_ra", "ccp_416x240_8b_ra",
           "tqb_416x240_10b_lowdelay", "wpp_416x240_8b_lowdelay", "amp_416x240_10b_ra", "pcm_416x240_10b_lfoff"]


def build():
    os.makedirs(OUT, exist_ok=True)
    subprocess.run(["make", "-C", os.path.join(ROOT, "tests", "emul")], check=True, capture_output=True)
    san = ["-fsanitize=address", "-fno-omit-frame-pointer"]
    subprocess.run(["g++", "-O1", "-g", "-std=c++17", "-fPIC", *san, "-c", os.path.join(ROOT, "openhevc_b200", "csrc", "recorder.cpp"), "-o", OUT + "/recorder.o"], check=True)
    objs = [os.path.join(REFD, "emul_obj", o) for o in ("asan_kernels.o", "asan_engine.o", "asan_warp_emul.o")]
    subprocess.run(["g++", "-shared", "-Wl,-Bsymbolic", "-fsanitize=address", "-o", OUT + "/libb200hevc_emul_asan.so", *objs, OUT + "/recorder.o", "-lpthread"], check=True)
    subprocess.run(["gcc", "-O1", "-g", "-std=gnu99", "-fPIC", "-w", "-shared", *san, "-I" + os.path.join(REFD, "gen"), "-I/root/reference", "-I" + os.path.join(ROOT, "include"),
                    "-o", OUT + "/libb200hevc_shim.so", os.path.join(ROOT, "openhevc_b200", "csrc", "shim", "hevcdsp_init_b200.c")], check=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seeds", type=int, default=5)
    a = ap.parse_args()
    build()
    rt = subprocess.run(["gcc", "-print-file-name=libasan.so"], capture_output=True, text=True).stdout.strip()
    env = dict(os.environ, LD_PRELOAD=f"{rt} {OUT}/libb200hevc_shim.so {OUT}/libb200hevc_emul_asan.so",
               ASAN_OPTIONS="detect_leaks=0:detect_stack_use_after_return=0:halt_on_error=1")
    runs = bad = 0
    for name in STREAMS:
        src = open(os.path.join(ROOT, "tests", "golden", "streams", name + ".hevc"), "rb").read()
        for seed in range(a.seeds):
            rnd = random.Random(seed * 13 + 5)
            data = bytearray(src)
            for _ in range(4):
                data[rnd.randrange(len(data) // 4, len(data))] ^= 1 << rnd.randrange(8)
            path = OUT + "/damaged.hevc"
            open(path, "wb").write(data)
            for threads in (("1", "4w") if name.startswith(("wpp", "tiles")) else ("1", "4")):
                runs += 1
                try:
                    r = subprocess.run([os.path.join(REFD, "decode_b200"), path, threads], capture_output=True, text=True, timeout=300, env=env)
                except subprocess.TimeoutExpired:
                    bad += 1
                    print(name, seed, threads, "TIMEOUT")
                    continue
                if "ERROR: AddressSanitizer" in r.stderr or r.returncode:
                    bad += 1
                    i = r.stderr.find("ERROR: AddressSanitizer")
                    print(name, seed, threads, "rc", r.returncode, r.stderr[i:i + 800] if i >= 0 else r.stderr[-300:])
    print(f"{runs} runs, {bad} bad")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
