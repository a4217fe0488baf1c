#!/usr/bin/env python
"""How long does page-locking host memory take on this box?  (start-up cost of the hooked decoder: pinned frames and recorder blobs)"""
import ctypes as C, time, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from openhevc_b200 import _lib
lib = _lib.load()
lib.b200_host_alloc.restype = C.c_void_p
t = time.time(); p = lib.b200_host_alloc(C.c_uint64(4096)); print("first call (context creation) %.3f s" % (time.time() - t))
for mb in (8, 17, 25, 64, 256, 1024):
    t = time.time(); q = lib.b200_host_alloc(C.c_uint64(mb << 20)); dt = time.time() - t
    print("cudaHostAlloc %5d MB: %7.1f ms  (%.2f GB/s)" % (mb, dt * 1e3, (mb / 1024) / dt))
    t = time.time(); lib.b200_host_free(C.c_void_p(q)); print("   free %.1f ms" % ((time.time() - t) * 1e3))
libc = C.CDLL("libc.so.6")
libc.aligned_alloc.restype = C.c_void_p
for mb in (25, 256):
    n = mb << 20
    buf = libc.aligned_alloc(C.c_size_t(4096), C.c_size_t(n))
    C.memset(C.c_void_p(buf), 0, n)
    t = time.time(); rc = lib.b200_host_register(C.c_void_p(buf), C.c_uint64(n)); dt = time.time() - t
    print("cudaHostRegister %4d MB (touched): %7.1f ms rc %d (%.2f GB/s)" % (mb, dt * 1e3, rc, (mb / 1024) / dt))
