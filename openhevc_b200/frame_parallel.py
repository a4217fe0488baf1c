"""Frame-parallel decoding of a random-access (hierarchical-B, GOP 8) stream over N GPUs.

Mirror of the reference's frame threads (libavcodec/pthread_frame.c; context i <-> GPU i mod N) at the
only granularity at which this path shards: whole pictures.  GOP g is owned by rank
g mod N.  Inside a GOP the seven B pictures only reference pictures of the same GOP and the two
surrounding anchors, so the ONE exchange step is the anchor picture: when the owner has finished
anchor A_g (the analogue of ff_thread_report_progress(&ref->tf, INT_MAX), hevc.c:4026) it is
broadcast into the same DPB slot on every peer (NCCL over NVLink; one 24.9 MB message per 8 pictures
at 4K Main10) on a side stream, overlapping the B pictures.  Non-anchor pictures never leave their GPU.

The schedule is written against a tiny backend interface so that the very same code runs
  * on GPUs  (GpuBackend: libb200hevc + torch.distributed/NCCL, CUDA events between streams), and
  * in the CPU test-suite (tests/test_frame_parallel_cpu.py: gloo, world_size 2, pictures "decoded" by the
    test oracle) -- the host logic, slot rotation and collective order are identical.
"""
from dataclasses import dataclass
from typing import List

N_ANCHOR_SLOTS = 16            # anchors rotate through slots 0..15: a slot is reused 16 GOPs later
N_B_SLOTS = 14                 # two sets of seven: consecutive GOPs do not wait for each other's B pictures (WAR on the slots)
N_SLOTS = N_ANCHOR_SLOTS + N_B_SLOTS
INTRA_PERIOD_GOPS = 4          # every 4th anchor is an I picture (intra period 32 pictures)

# blob indices of the 9 distinct pictures of the periodic stream
BLOB_ANCHOR_P, BLOB_ANCHOR_I, BLOB_B4, BLOB_B2, BLOB_B6, BLOB_B1, BLOB_B3, BLOB_B5, BLOB_B7 = range(9)
N_BLOBS = 9


@dataclass
class Picture:
    blob: int
    cur_slot: int
    ref_slots: List[int]
    anchor: bool


def anchor_slot(g):
    return g % N_ANCHOR_SLOTS


def gop_pictures(g):
    """decode-order pictures of GOP g (POC order 8g+{8,4,2,6,1,3,5,7}) with their DPB placement"""
    a_prev, a_cur = anchor_slot(g - 1), anchor_slot(g)
    b = [N_ANCHOR_SLOTS + 7 * (g & 1) + k for k in range(7)]     # b4, b2, b6, b1, b3, b5, b7
    intra = g % INTRA_PERIOD_GOPS == 0
    return [
        Picture(BLOB_ANCHOR_I if intra else BLOB_ANCHOR_P, a_cur, [] if intra else [a_prev], True),
        Picture(BLOB_B4, b[0], [a_prev, a_cur], False),
        Picture(BLOB_B2, b[1], [a_prev, b[0]], False),
        Picture(BLOB_B6, b[2], [b[0], a_cur], False),
        Picture(BLOB_B1, b[3], [a_prev, b[1]], False),
        Picture(BLOB_B3, b[4], [b[1], b[0]], False),
        Picture(BLOB_B5, b[5], [b[0], b[2]], False),
        Picture(BLOB_B7, b[6], [b[2], a_cur], False),
    ]


def blob_specs():
    """(name, n_refs) of the distinct work lists the stream is made of"""
    return [("anchor_P", 1), ("anchor_I", 0)] + [(n, 2) for n in ("b4", "b2", "b6", "b1", "b3", "b5", "b7")]


def run_schedule(backend, rank, world, n_gops_per_rank):
    """Issue the whole schedule (asynchronously on a GPU backend).  Returns pictures decoded by this rank."""
    total = n_gops_per_rank * world
    decoded = 0
    for g in range(total):
        owner = g % world
        pics = gop_pictures(g)
        if owner == rank:
            if pics[0].ref_slots:
                backend.wait_anchor(g - 1)            # A_{g-1} must have arrived (or been decoded) here
            backend.decode(pics[0])
            backend.anchor_decoded(g)
            decoded += 1
        if world > 1:
            backend.broadcast_anchor(g, anchor_slot(g), owner)      # collective: every rank, same order
        if owner == rank:
            backend.wait_anchor(g - 1)
            for p in pics[1:]:
                backend.decode(p)
            backend.gop_done(g)
            decoded += len(pics) - 1
    return decoded


class GpuBackend:
    """libb200hevc engine + NCCL.  The DPB lives in a torch tensor so that torch.distributed can address slots.
    Ordering on the device is the engine's slot-hazard tracking (include/b200hevc.h: b200_slot_begin/end_access):
    the broadcast is a reader of the anchor's slot on its owner and a writer of the slot everywhere else, on the
    communication stream, so it overlaps the B pictures of the previous GOP on the compute lanes."""

    def __init__(self, engine, dpb_tensor, slot_bytes, world, arenas_of_blob):
        import torch
        self.torch = torch
        self.eng, self.dpb, self.slot_bytes, self.world = engine, dpb_tensor, slot_bytes, world
        self.arena = arenas_of_blob
        self.compute = torch.cuda.ExternalStream(engine.lib.b200_stream(engine.h))
        self.comm = torch.cuda.Stream() if world > 1 else None
        self.bcast_bytes = 0

    def decode(self, pic):
        self.eng.execute(self.arena[pic.blob], pic.cur_slot, pic.ref_slots)

    def anchor_decoded(self, g):
        pass                                           # the engine recorded the slot's "written" event

    def wait_anchor(self, g):
        pass                                           # pictures wait for the writers of their reference slots

    def broadcast_anchor(self, g, slot, owner):
        torch = self.torch
        import torch.distributed as dist
        t = self.dpb[slot * self.slot_bytes:(slot + 1) * self.slot_bytes]
        write = dist.get_rank() != owner
        with torch.cuda.stream(self.comm):
            self.eng.slot_begin_access(slot, self.comm.cuda_stream, write)
            dist.broadcast(t, src=owner)
            self.eng.slot_end_access(slot, self.comm.cuda_stream, write)
        self.bcast_bytes += self.slot_bytes

    def gop_done(self, g):
        pass
