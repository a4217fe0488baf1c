// common.cuh — shared device-side declarations for the B200 HEVC reconstruction kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/b200hevc_worklist.h"

struct PlaneDesc {
    uint8_t *base;   // device pointer to sample (0,0)
    int      pitch;  // bytes per row
    int      w, h;   // samples
};
struct FrameDesc {
    PlaneDesc p[3];
};
struct RefTable {
    uint64_t w[2];      // B200BlobHeader.ref_slot[16] packed little-endian (possibly overridden at execute time)
};

__device__ __forceinline__ int clip3i(int v, int lo, int hi) { return min(max(v, lo), hi); }
__device__ __forceinline__ int clip16i(int v) { return min(max(v, -32768), 32767); }

// plane descriptor of a kernel-parameter FrameDesc by register selects: indexing the parameter struct with a runtime
// plane number would make the compiler copy it to local memory (an L2 round trip per access for a cold warp)
__device__ __forceinline__ PlaneDesc plane_of(const FrameDesc &f, int plane)
{
    PlaneDesc d;
    d.base = plane == 0 ? f.p[0].base : plane == 1 ? f.p[1].base : f.p[2].base;
    d.pitch = plane == 0 ? f.p[0].pitch : plane == 1 ? f.p[1].pitch : f.p[2].pitch;
    d.w = plane == 0 ? f.p[0].w : plane == 1 ? f.p[1].w : f.p[2].w;
    d.h = plane == 0 ? f.p[0].h : plane == 1 ? f.p[1].h : f.p[2].h;
    return d;
}

template <typename PIX>
__device__ __forceinline__ PIX *px_ptr(const PlaneDesc &p, int x, int y)
{
    return reinterpret_cast<PIX *>(p.base + (size_t)y * p.pitch) + x;
}

// launchers (kernels.cu) -- all asynchronous on `st`; return number of kernels launched
int launch_mc(cudaStream_t st, const B200McRec *recs, int count, int n_big, const FrameDesc &cur, const FrameDesc *dpb_dev, const RefTable &rt, int bd);
int launch_residual(cudaStream_t st, const B200TuRec *const recs[4], const int counts[4], const int16_t *pool, int16_t *parked, const FrameDesc &cur, int bd);
int launch_intra(cudaStream_t st, const B200IntraRec *recs, int count, const int16_t *pool, const FrameDesc &cur, int bd,
                 uint2 *edges[3], const int edge_stride[3], uint32_t *counter);
int launch_deblock(cudaStream_t st, const uint16_t *grid, const B200DbkLayout &L, const FrameDesc &cur, int bd);
int launch_sao(cudaStream_t st, const B200SaoRec *grid, const FrameDesc &src, const FrameDesc &dst, int bd,
               int log2_ctb, int ctb_w, int ctb_h, int chroma_format_idc);
int set_intra_trace(unsigned long long *p);
int launch_fill(cudaStream_t st, const FrameDesc &f, int bd, int value);
