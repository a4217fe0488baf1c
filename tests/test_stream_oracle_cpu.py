"""CPU suite, end to end without a GPU: the REAL decoder (reference + the hooks of INTEGRATION.md) parses the committed
Annex-B streams with the B200 tables installed in record-only mode (B200_SHIM_DUMP: every picture's work list is written
to a file, nothing is sent to a device); the CPU oracle executes the dumped work lists picture after picture on its own
DPB; the per-plane MD5s must equal those of the UNMODIFIED reference decoder committed next to the streams.

What this pins: the shim's pointer -> (slot, plane, x, y) mapping, the recorder and the wire format (incl. sparse
coefficients, MC tile splitting, intra level ordering, deblock / SAO grids, constrained_intra_pred bitmap), and the
oracle itself against the reference decoder on real syntax -- the GPU suite then only has to show kernel == oracle."""
import ctypes as C
import glob
import hashlib
import os
import subprocess
import tempfile

import numpy as np
import pytest

import oracle_lib
from openhevc_b200 import worklist as W

HERE = os.path.dirname(os.path.abspath(__file__))
REFDIR = os.path.join(os.path.dirname(HERE), "oracle", "_ref")
STREAMS = sorted(glob.glob(os.path.join(HERE, "golden", "streams", "*.hevc")))
SMALL = [s for s in STREAMS if os.path.getsize(s) < 400_000]          # the 1080p / 4K streams run in the GPU suite


@pytest.mark.parametrize("dbd", ["derive", "check"])
@pytest.mark.parametrize("threads", ["1", "4", "4w", "2x"])
@pytest.mark.parametrize("stream", SMALL, ids=[os.path.basename(s) for s in SMALL])
def test_recorded_work_lists_through_the_oracle_equal_the_reference_decoder(stream, threads, dbd):
    """threads: "1"; "4" = four frame threads (pictures recorded concurrently, dumped in decode order by the shim's ticket);
    "4w" = four slice-thread workers (WPP rows, or tiles) recording one picture together (merged by b200_rec_merge) -- streams with
    entry points only; "2x" = frame threads whose pictures are decoded by two slice threads each (hevc -f 4, pthread.c:57-71:
    several pictures in progress, each recorded by several workers -- b200_worker_begin tells a worker which one it is on).
    dbd: "derive" (the default of the drop-in) = the work lists carry the INPUTS of the deblocking control
    (transform-tree leaves, QP map, slice offsets; motion and cbf are in the MC / TU records) and the oracle derives boundary
    strengths, tc and beta itself (orc_dbd_derive, restating hevc_filter.c:345-581, 583-700, 805-941); "check" (B200_DBD=2) = the
    reference's own filter calls are recorded as well and the oracle requires the two grids to be identical (error -10)."""
    if dbd == "check" and threads != "1":
        pytest.skip("the comparison runs single-threaded")
    binary = os.path.join(REFDIR, "decode_b200")
    if not os.path.exists(binary):
        pytest.skip("oracle/_ref/decode_b200 not built (needs /root/reference)")
    if threads in ("4w", "2x") and not os.path.basename(stream).startswith(("wpp_", "tiles_")):
        pytest.skip("no entry points: slice threads fall back to one thread")
    # the arbiter is the UNMODIFIED decoder run the same way.  Single thread: that is the committed MD5 file.  With threads
    # the reference may differ from itself: it never clears s->is_pcm between pictures (hevc.c:147, only allocated zeroed), so
    # on transquant-bypass streams the flags a context sees depend on which pictures it decoded before -- i.e. on the
    # thread count.  The drop-in reads the same array and must reproduce whatever the reference does with it.  (On streams
    # shorter than the thread count the reference's flush logic, main_hm/main.c:283, also drops the delayed pictures.)
    ref = subprocess.run([os.path.join(REFDIR, "decode_ref"), stream, threads], capture_output=True, text=True, timeout=600)
    want = [l for l in ref.stdout.splitlines() if l.startswith("frame ")]
    committed = open(stream[:-5] + ".md5").read().splitlines()
    if threads == "1":
        assert want == committed
    elif threads in ("4w", "2x") and len(want) == len(committed):
        # slice threads never change what a picture is; the unmodified decoder's own slice-threaded runs can (a race in its host
        # pixel path, seen on 4K WPP streams: DESIGN.md 6) -- the single-threaded run is the arbiter whenever all pictures came out
        want = committed
    if not want:
        pytest.skip("the reference outputs no picture of this stream with that many threads")
    with tempfile.TemporaryDirectory() as d:
        r = subprocess.run([binary, stream, threads, "quiet"], capture_output=True, text=True, timeout=600,
                           env=dict(os.environ, B200_SHIM_DUMP=d, B200_SHIM_STATS="1"))
        assert r.returncode == 0, r.stderr[-2000:]
        gen_path = stream[:-5] + ".gen.txt"
        if os.path.exists(gen_path) and threads not in ("4", "2x"):      # B200_SHIM_STATS: the table calls per picture == what the generator wrote
            import re                                        # (with frame threads the lines come in completion order)
            gen = [tuple(int(v) for v in m.groups()) for m in re.finditer(r"intra_pred (\d+) transform_add (\d+) prediction units (\d+)", open(gen_path).read())]
            got = [tuple(int(v) for v in m.groups()) for m in re.finditer(r"b200 picture \d+: intra_pred (\d+) transform_add (\d+) mc (\d+)", r.stderr)]
            assert len(got) == len(gen)
            for (gi, gt, gm), (wi, wt, wp) in zip(got, gen):
                assert (gi, gt) == (wi, wt) and gm == 3 * wp
        blobs = sorted(glob.glob(os.path.join(d, "pic_*.blob")))
        assert len(blobs) >= len(want), f"{len(blobs)} work lists for {len(want)} pictures"
        lib = oracle_lib.lib()
        slots = {}
        results = []
        dummy = np.zeros(1, np.uint16)
        for k, path in enumerate(blobs):
            blob = np.fromfile(path, np.uint8)
            hdr, _ = W.parse_blob(blob)
            w, h, cfi, bd = int(hdr["width"]), int(hdr["height"]), int(hdr["chroma_format_idc"]), int(hdr["bit_depth"])
            cur = int(hdr["cur_slot"])
            # grey reference pictures the decoder generated for this picture (generate_missing_ref -> b200_frame_fill)
            if os.path.exists(path[:-5] + ".fill"):
                for line in open(path[:-5] + ".fill"):
                    sl, val = [int(v) for v in line.split()]
                    slots[sl] = [np.full(W.plane_dims(w, h, cfi, p)[::-1], val, np.uint16) for p in range(3)]
            # a new picture in a slot starts from the decoder's fresh frame; the oracle overwrites every sample anyway
            slots[cur] = [np.zeros(W.plane_dims(w, h, cfi, p)[::-1], np.uint16) for p in range(3)]
            for i in range(int(hdr["n_ref"])):
                assert int(hdr["ref_slot"][i]) in slots, f"picture {k} references an empty DPB slot"
            n_slots = 32
            flat = [slots[s][p] if s in slots else dummy for s in range(n_slots) for p in range(3)]
            ptrs = (C.c_void_p * len(flat))(*[a.ctypes.data for a in flat])
            assert lib.orc_execute_blob(blob.ctypes.data_as(C.c_void_p), ptrs, n_slots) == 0
            md5 = [hashlib.md5((pl.astype(np.uint8) if bd == 8 else pl.astype("<u2")).tobytes()).hexdigest() for pl in slots[cur]]
            results.append((int(hdr["poc"]), f"{w}x{h} bd{bd} " + " ".join(md5)))
        # the decoder outputs in POC order (one coded video sequence per stream); with low-delay streams that is the decode order
        results.sort(key=lambda t: t[0])
        for k, (poc, line) in enumerate(results[:len(want)]):
            assert f"frame {k} " + line == want[k], f"picture {k} (POC {poc}) of {os.path.basename(stream)}"
