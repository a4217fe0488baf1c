#!/usr/bin/env python
"""Multi-GPU parity check (run under torchrun, one rank per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/verify_multi_gpu.py
Every rank runs the product schedule (openhevc_b200/frame_parallel.py: GOP ownership, NCCL anchor broadcast through the
engine's slot-hazard protocol -- ownership by intra period, one anchor per period sent to the next GPU --, 8 compute lanes) and reads every picture it decodes back; rank 0 then decodes the same
stream sequentially with the CPU oracle (test infrastructure) and compares the MD5 of every picture of every rank.
Also run by hand with --nproc-per-node 1 (lanes only)."""
import argparse
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gops", type=int, default=8, help="GOPs per rank (whole intra periods of 4)")
    ap.add_argument("--size", default="832x480")
    ap.add_argument("--bit-depth", type=int, default=10)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    from openhevc_b200 import FrameEngine, _lib
    from openhevc_b200 import frame_parallel as FP
    from openhevc_b200 import worklist as W
    from openhevc_b200.synth import FrameSynth, smooth_frame
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    w, h = [int(v) for v in args.size.split("x")]
    cfi, bd = 1, args.bit_depth
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = _lib.load()
    cfg = _lib.B200Config(local, w, h, cfi, bd, 6, FP.N_SLOTS, 16, 0, None, 0, 0, 0)
    slot_bytes = int(lib.b200_dpb_bytes(cfg)) // FP.N_SLOTS
    dpb = torch.zeros(FP.N_SLOTS * slot_bytes, dtype=torch.uint8, device=f"cuda:{local}")
    eng = FrameEngine(w, h, cfi, bd, n_slots=FP.N_SLOTS, n_arenas=16, device=local, ext_frame_mem=dpb.data_ptr(), ext_frame_bytes=dpb.numel())
    blobs = []
    for i, (name, n_ref) in enumerate(FP.blob_specs()):
        blob, _ = FrameSynth(w, h, cfi, bd, seed=900 + i, refs=list(range(n_ref)), cur_slot=2, p_intra=0.1 if n_ref else 1.0).generate()
        blobs.append(blob)
    start = smooth_frame(w, h, cfi, bd, 7)
    eng.upload_slot(FP.anchor_slot(-1), start)
    for b, blob in enumerate(blobs):
        eng.upload(blob, b)
    eng.sync()

    class Checked(FP.GpuBackend):
        def __init__(self, *a):
            super().__init__(*a)
            self.out, self.g, self.last_anchor = {}, -1, -1

        def decode(self, pic, g=None):
            super().decode(pic, g)
            self.out[(g, pic.blob)] = self.eng.readback(pic.cur_slot, self.eng.new_host_frame(pinned=True), sync=False)

    be = Checked(eng, dpb, slot_bytes, world, list(range(FP.N_BLOBS)))
    FP.run_schedule(be, rank, world, args.gops)
    eng.sync()
    torch.cuda.synchronize()
    mine = {f"{g}:{b}": hashlib.md5(b"".join(p.tobytes() for p in planes)).hexdigest() for (g, b), planes in be.out.items()}
    allr = [None] * world
    if world > 1:
        dist.all_gather_object(allr, mine)
    else:
        allr = [mine]
    ok = True
    if rank == 0:
        import oracle_lib
        got = {}
        for d in allr:
            got.update(d)
        ref_dpb = [[np.zeros_like(p) for p in start] for _ in range(FP.N_SLOTS)]
        ref_dpb[FP.anchor_slot(-1)] = start
        want = {}
        for g in range(-(-args.gops // FP.INTRA_PERIOD_GOPS) * FP.INTRA_PERIOD_GOPS * world):
            for pic in FP.gop_pictures(g):
                blob = blobs[pic.blob].copy()
                hdr = blob[:256].view(W.header_dt)
                hdr["cur_slot"] = pic.cur_slot
                hdr["n_ref"] = len(pic.ref_slots)
                hdr["ref_slot"][0][:len(pic.ref_slots)] = pic.ref_slots
                out = oracle_lib.execute(blob, ref_dpb)
                ref_dpb[pic.cur_slot] = [p.copy() for p in out]
                want[f"{g}:{pic.blob}"] = hashlib.md5(b"".join(p.astype(eng.dtype).tobytes() for p in out)).hexdigest()
        bad = sorted(k for k in want if got.get(k) != want[k])
        ok = not bad and set(got) == set(want)
        print(json.dumps({"verify_multi_gpu": "ok" if ok else "MISMATCH", "world": world, "pictures": len(want), "size": args.size,
                          "bit_depth": bd, "bad": bad[:8]}), flush=True)
    eng.close()
    if world > 1:
        dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
