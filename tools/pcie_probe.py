#!/usr/bin/env python
"""Pinned-memory copy bandwidth of this box (the ceiling of bench.py's e2e number): D2H, H2D, both at once."""
import torch

n = 256 << 20
dev = torch.empty(n, dtype=torch.uint8, device="cuda")
dev2 = torch.empty(n, dtype=torch.uint8, device="cuda")
h1 = torch.empty(n, dtype=torch.uint8).pin_memory()
h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def t(fn, reps=8):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    s1.wait_event(e0); s2.wait_event(e0)
    for _ in range(reps):
        fn()
    e = torch.cuda.Event(); e.record(s1); torch.cuda.current_stream().wait_event(e)
    e = torch.cuda.Event(); e.record(s2); torch.cuda.current_stream().wait_event(e)
    e1.record(); torch.cuda.synchronize()
    return reps * n / (e0.elapsed_time(e1) * 1e-3) / 1e9


def d2h():
    with torch.cuda.stream(s1):
        h1.copy_(dev, non_blocking=True)


def h2d():
    with torch.cuda.stream(s2):
        dev2.copy_(h2, non_blocking=True)


def both():
    d2h(); h2d()


print("torch pin_memory():   D2H %.1f GB/s  H2D %.1f GB/s  both: %.1f GB/s each direction" % (t(d2h), t(h2d), t(both)))

# the same with the library's own staging allocator (b200_host_alloc: placed on the GPU's NUMA node)
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from openhevc_b200 import engine as E, _lib   # noqa: E402
_b1, _b2 = E.PinnedBuffer(_lib.load(), n), E.PinnedBuffer(_lib.load(), n)
h1, h2 = torch.from_numpy(_b1.array), torch.from_numpy(_b2.array)
print("b200_host_alloc():    D2H %.1f GB/s  H2D %.1f GB/s  both: %.1f GB/s each direction" % (t(d2h), t(h2d), t(both)))
