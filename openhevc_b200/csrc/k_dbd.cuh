// k_dbd.cuh — deblocking parameters derived on the device.
//
// The reference derives them on the host, between the parse and the filter calls:
//   ff_hevc_deblocking_boundary_strengths()  hevc_filter.c:805-941   bs per 4-sample edge segment from the motion field
//                                                                     (boundary_strength(), :583-700) and the coded-block flags
//   deblocking_filter_CTB()                  hevc_filter.c:345-581   tc / beta per edge from the QP map and the slice offsets
// 19-26 % of the hooked decoder's host time on lightly coded 4K content.  Everything those two functions read is either already in
// the picture's work list -- the motion of every prediction block (luma MC records: position, mv = source position + fraction,
// reference picture = DPB slot), the luma coded-block flags (a luma transform record exists exactly where cbf_luma is set,
// hevc.c:1568-1576 / hevc_cabac.c:1949) -- or small (B200DbdHeader: one word per transform-tree leaf, the QP map, two offsets
// per CTB).  Four kernels rebuild the reference's state and write the SAME dense grids the recorded path fills from the
// reference's own filter calls (b200_rec_deblock), so K4 is unchanged and the two paths can be compared entry by entry
// (B200_DBD_CHECK).
//   K4a k_dbd_raster   MC / TU records -> per-4x4-unit motion map (16 B) and cbf map
//   K4b k_dbd_bs       leaves -> boundary strengths of vertical / horizontal edges (one warp per leaf)
//   K4c k_dbd_params   bs + QP + offsets -> grid entries (one thread per 8-sample edge position and pass)
#pragma once

struct DbdArgs {
    const uint8_t *blob;    // device copy of the picture's blob
    B200Section mc, tu[4];
    const uint32_t *leaf; int n_leaf;
    const int8_t *qp; int log2_min_cb, min_cb_w;
    const int8_t *ctb;      // beta_offset, tc_offset per CTB
    const uint8_t *pcm; int log2_min_pu, min_pu_w, min_pu_h;     // pcm == nullptr: no PCM-loop-filter-off / bypass blocks in this stream
    int cb_qp_offset, cr_qp_offset;
    int width, height, cfi, log2_ctb, ctb_w;
    RefTable rt;
};

__device__ __forceinline__ int dbd_slot(const RefTable &rt, int i)
{
    const uint64_t w = (i & 8) ? rt.w[1] : rt.w[0];
    return (int)((w >> (8 * (i & 7))) & 0xff);
}

// K4a: one thread per record.  MC records of the luma plane give the unit its motion; luma TU records (not PCM) its cbf.
__global__ void __launch_bounds__(256) k_dbd_raster(DbdArgs a, DbdMaps m, const uint32_t *__restrict__ gate)
{
    if (__ldg(gate + 1)) return;
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < a.mc.count) {
        const int4 *p = reinterpret_cast<const int4 *>(a.blob + a.mc.off) + 2 * (size_t)i;
        const int4 ra = __ldg(p), rb = __ldg(p + 1);
        const int plane = (ra.y >> 16) & 0xff;
        if (plane == 0) {
            const int x = ra.x & 0xffff, y = (unsigned)ra.x >> 16, w = ra.y & 0xff, h = (ra.y >> 8) & 0xff, flags = (unsigned)ra.y >> 24;
            const int sx0 = (int16_t)(ra.z & 0xffff), sy0 = (int16_t)((unsigned)ra.z >> 16), sx1 = (int16_t)(ra.w & 0xffff), sy1 = (int16_t)((unsigned)ra.w >> 16);
            const int ref0 = rb.x & 0xff, ref1 = (rb.x >> 8) & 0xff, f0 = (rb.x >> 16) & 0xff, f1 = (unsigned)rb.x >> 24;
            const bool bi = flags & B200_MCF_BI;
            // the source position is destination + (mv >> 2) and the fraction mv & 3 (hevc.c:1656-1663): mv back from the two
            const int mv0x = ((sx0 - x) << 2) | (f0 & 3), mv0y = ((sy0 - y) << 2) | ((f0 >> 4) & 3);
            const int mv1x = ((sx1 - x) << 2) | (f1 & 3), mv1y = ((sy1 - y) << 2) | ((f1 >> 4) & 3);
            uint4 v;
            v.x = (uint32_t)(mv0x & 0xffff) | ((uint32_t)(mv0y & 0xffff) << 16);
            v.y = bi ? (uint32_t)(mv1x & 0xffff) | ((uint32_t)(mv1y & 0xffff) << 16) : 0u;
            v.z = (uint32_t)dbd_slot(a.rt, ref0) | (bi ? (uint32_t)dbd_slot(a.rt, ref1) << 8 : 0u) | (bi ? 3u << 16 : 1u << 16);
            v.w = 0;
            for (int uy = y >> 2; uy < ((y + h + 3) >> 2) && uy < m.uh; uy++)
                for (int ux = x >> 2; ux < ((x + w + 3) >> 2) && ux < m.uw; ux++) m.mot[(size_t)uy * m.uw + ux] = v;
        }
    }
#pragma unroll
    for (int s = 0; s < 4; s++) {
        if (i >= a.tu[s].count) continue;
        const int4 raw = __ldg(reinterpret_cast<const int4 *>(a.blob + a.tu[s].off) + i);
        const int x = raw.x & 0xffff, y = (unsigned)raw.x >> 16, plane = raw.y & 0xff, kind = (raw.y >> 16) & 0xff;
        if (plane != 0 || kind == B200_TU_PCM) continue;
        const int u = 1 << s;
        for (int uy = y >> 2; uy < (y >> 2) + u && uy < m.uh; uy++)
            for (int ux = x >> 2; ux < (x >> 2) + u && ux < m.uw; ux++) m.cbf[(size_t)uy * m.uw + ux] = 1;
    }
}

__device__ __forceinline__ int dbd_far(uint32_t a, uint32_t b)           // |dx| >= 4 || |dy| >= 4 of two packed motion vectors
{
    const int dx = (int16_t)(a & 0xffff) - (int16_t)(b & 0xffff), dy = (int16_t)(a >> 16) - (int16_t)(b >> 16);
    return abs(dx) >= 4 || abs(dy) >= 4;
}
// boundary_strength(), hevc_filter.c:583-700 (the build compares reference pictures by POC, hevc.h:73; a POC names one picture of
// the DPB, i.e. one slot): both blocks inter coded
__device__ __forceinline__ int dbd_bs_motion(const uint4 c, const uint4 n)
{
    const int pc = (c.z >> 16) & 3, pn = (n.z >> 16) & 3;
    const int c0 = c.z & 0xff, c1 = (c.z >> 8) & 0xff, n0 = n.z & 0xff, n1 = (n.z >> 8) & 0xff;
    if (pc == 3 && pn == 3) {
        if (c0 == n0 && c0 == c1 && n0 == n1)
            return (dbd_far(n.x, c.x) || dbd_far(n.y, c.y)) && (dbd_far(n.y, c.x) || dbd_far(n.x, c.y));
        if (n0 == c0 && n1 == c1) return dbd_far(n.x, c.x) || dbd_far(n.y, c.y);
        if (n1 == c0 && n0 == c1) return dbd_far(n.y, c.x) || dbd_far(n.x, c.y);
        return 1;
    }
    if (pc != 3 && pn != 3) return c0 == n0 ? dbd_far(c.x, n.x) : 1;
    return 1;
}
__device__ __forceinline__ int dbd_bs_edge(const DbdMaps &m, int ux, int uy, int px, int py)     // current unit, neighbour unit
{
    const size_t ic = (size_t)uy * m.uw + ux, in = (size_t)py * m.uw + px;
    const uint4 c = m.mot[ic], n = m.mot[in];
    if (!((c.z >> 16) & 3) || !((n.z >> 16) & 3)) return 2;
    if (m.cbf[ic] | m.cbf[in]) return 1;
    return dbd_bs_motion(c, n);
}

// K4b: one warp per leaf (a transform block or a coding block without residual), hevc_filter.c:805-941
__global__ void __launch_bounds__(256) k_dbd_bs(DbdArgs a, DbdMaps m, const uint32_t *__restrict__ gate)
{
    if (__ldg(gate + 1)) return;
    const int li = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (li >= a.n_leaf) return;
    const uint32_t L = __ldg(a.leaf + li);
    const int ux0 = L & 0xfff, uy0 = (L >> 12) & 0xfff, log2 = ((L >> 24) & 7) + 2, nu = 1 << (log2 - 2);
    if (ux0 + nu > m.uw + 15 || uy0 + nu > m.uh + 15) return;                      // (validated on the host as well)
    if ((L >> 28) & 1)                                                             // top edge, :832-866
        for (int i = lane; i < nu; i += 32)
            if (ux0 + i < m.uw && uy0 > 0 && uy0 < m.uh) m.bsh[(size_t)uy0 * m.uw + ux0 + i] = (uint8_t)dbd_bs_edge(m, ux0 + i, uy0, ux0 + i, uy0 - 1);
    if ((L >> 29) & 1)                                                             // left edge, :870-904
        for (int i = lane; i < nu; i += 32)
            if (uy0 + i < m.uh && ux0 > 0 && ux0 < m.uw) m.bsv[(size_t)(uy0 + i) * m.uw + ux0] = (uint8_t)dbd_bs_edge(m, ux0, uy0 + i, ux0 - 1, uy0 + i);
    // prediction-block boundaries inside the leaf (:906-940): only for leaves larger than the smallest prediction block and not intra
    if (log2 <= a.log2_min_pu || ux0 >= m.uw || uy0 >= m.uh) return;
    if (!((m.mot[(size_t)uy0 * m.uw + ux0].z >> 16) & 3)) return;
    const int ne = nu >> 1;                       // edges on the 8-sample grid per row / column of units, the first one excluded
    for (int k = lane; k < nu * (ne - 1); k += 32) {
        const int i = k % nu, e = k / nu + 1;     // unit along the edge, edge number (at 8 * e samples)
        // horizontal edge at y0 + 8e, column x0 + 4i: the neighbour above is the row the reference's running `top` points at --
        // row y0 + 7 for the first edge, the row of the previous edge afterwards
        {
            const int ux = ux0 + i, uy = uy0 + 2 * e, py = e == 1 ? uy0 + 1 : uy0 + 2 * (e - 1);
            if (ux < m.uw && uy < m.uh) {
                const uint4 c = m.mot[(size_t)uy * m.uw + ux], n = m.mot[(size_t)py * m.uw + ux];
                m.bsh[(size_t)uy * m.uw + ux] = (uint8_t)(((c.z >> 16) & 3) && ((n.z >> 16) & 3) ? dbd_bs_motion(c, n) : 1);
            }
        }
        {
            const int uy = uy0 + i, ux = ux0 + 2 * e, px = e == 1 ? ux0 + 1 : ux0 + 2 * (e - 1);
            if (ux < m.uw && uy < m.uh) {
                const uint4 c = m.mot[(size_t)uy * m.uw + ux], n = m.mot[(size_t)uy * m.uw + px];
                m.bsv[(size_t)uy * m.uw + ux] = (uint8_t)(((c.z >> 16) & 3) && ((n.z >> 16) & 3) ? dbd_bs_motion(c, n) : 1);
            }
        }
    }
}

__constant__ uint8_t c_tctable[54] = { 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 1, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4,
                                       5, 5, 6, 6, 7, 8, 9, 10, 11, 13, 14, 16, 18, 20, 22, 24 };
__constant__ uint8_t c_betatable[52] = { 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 20, 22, 24, 26, 28, 30, 32, 34, 36,
                                         38, 40, 42, 44, 46, 48, 50, 52, 54, 56, 58, 60, 62, 64 };
__constant__ uint8_t c_qp_c[14] = { 29, 30, 31, 32, 33, 33, 34, 34, 35, 35, 36, 36, 37, 37 };

__device__ __forceinline__ int dbd_qpy(const DbdArgs &a, int x, int y) { return a.qp[(x >> a.log2_min_cb) + (y >> a.log2_min_cb) * a.min_cb_w]; }
__device__ __forceinline__ int dbd_pcm(const DbdArgs &a, int x, int y)           // get_pcm(), hevc_filter.c:324-338: non-zero = leave the samples alone
{
    if (x < 0 || y < 0) return 1;
    const int xp = x >> a.log2_min_pu, yp = y >> a.log2_min_pu;
    if (xp >= a.min_pu_w || yp >= a.min_pu_h) return 1;
    return a.pcm[yp * a.min_pu_w + xp] != 0;
}
__device__ __forceinline__ int dbd_tc(int qp, int bs, int tc_offset)            // TC_CALC, hevc_filter.c:340-343
{
    return c_tctable[clip3i(qp + 2 * (bs - 1) + ((tc_offset >> 1) << 1), 0, 53)];
}
__device__ __forceinline__ int dbd_chroma_tc(const DbdArgs &a, int qp_y, int c_idx, int tc_offset)    // chroma_tc(), hevc_filter.c:62-88
{
    const int qp_i = clip3i(qp_y + (c_idx == 1 ? a.cb_qp_offset : a.cr_qp_offset), 0, 57);
    int qp;
    if (a.cfi == 1) qp = qp_i < 30 ? qp_i : qp_i > 43 ? qp_i - 6 : c_qp_c[qp_i - 30];
    else qp = clip3i(qp_i, 0, 51);
    return c_tctable[clip3i(qp + 2 + tc_offset, 0, 53)];
}
// what b200_rec_deblock() writes for one filter call of the reference: two 4-sample segments along the edge
__device__ __forceinline__ void dbd_emit(const DbdMaps &m, const B200DbkLayout &L, int plane, int vertical, int x, int y, int pw, int ph, int beta,
                                         int tc0, int tc1, int nop0, int nop1, int noq0, int noq1)
{
    uint16_t *g = m.grid + L.off[plane][vertical ? 0 : 1];
    const int gs = (int)L.stride[plane][vertical ? 0 : 1];
#pragma unroll
    for (int j = 0; j < 2; j++) {
        const int sx = vertical ? x : x + 4 * j, sy = vertical ? y + 4 * j : y;
        if (sx >= pw || sy >= ph) continue;
        const int idx = vertical ? (sy >> 2) * gs + (sx >> 3) : (sy >> 3) * gs + (sx >> 2);
#ifdef B200_DBD_FAULT
        if (plane == 0 && !vertical && (idx % 97) == 5) beta ^= 1;     /* fault injection for the test of the check mode */
#endif
        g[idx] = B200_DBK_PACK(j ? tc1 : tc0, plane ? 0 : beta, j ? nop1 : nop0, j ? noq1 : noq0);
    }
}

// K4c: deblocking_filter_CTB(), hevc_filter.c:345-581, turned inside out: one thread per 8x8 luma position (x, y) and pass
// (0 vertical luma, 1 vertical chroma, 2 horizontal luma, 3 horizontal chroma).  The CTB whose call of deblocking_filter_CTB
// visits the position supplies the offsets: the CTB the position lies in for vertical edges; for horizontal edges the loop
// of a CTB starts 8 (chroma: 8 << hshift) samples to the left of it and ends as much before its right border, and the first
// position takes beta_offset (luma) / the first segment's tc_offset (chroma) from the CTB to the left (:477, 518).
__global__ void __launch_bounds__(256) k_dbd_params(DbdArgs a, DbdMaps m, B200DbkLayout L, const uint32_t *__restrict__ gate)
{
    if (__ldg(gate + 1)) return;
    const int X = blockIdx.x * blockDim.x + threadIdx.x, Y = blockIdx.y, pass = blockIdx.z;
    const int x = 8 * X, y = 8 * Y;
    if (x >= a.width || y >= a.height) return;
    const int hs = a.cfi != 3, vs = a.cfi == 1, h = 1 << hs, v = 1 << vs;
    const int ctb = 1 << a.log2_ctb;
    const bool chroma = pass & 1, horiz = pass >> 1;
    if (chroma && ((x & (8 * h - 1)) || (y & (8 * v - 1)))) return;
    const int sx = chroma ? 8 * h : 8, sy = chroma ? 8 * v : 8;          // step of the reference's loops
    const int qx = chroma ? 4 * h : 4, qy = chroma ? 4 * v : 4;          // second segment
    const int pwc = a.width >> hs, phc = a.height >> vs;
    if (!horiz) {
        if (x == 0) return;
        const int ci = (y >> a.log2_ctb) * a.ctb_w + (x >> a.log2_ctb);
        const int beta_offset = a.ctb[2 * ci], tc_offset = a.ctb[2 * ci + 1];
        const int bs0 = m.bsv[(size_t)(y >> 2) * m.uw + (x >> 2)];
        const int bs1 = y + qy < a.height ? m.bsv[(size_t)((y + qy) >> 2) * m.uw + (x >> 2)] : 0;
        int nop0 = 0, nop1 = 0, noq0 = 0, noq1 = 0;
        if (!chroma) {
            if (!(bs0 | bs1)) return;
            const int qp = (dbd_qpy(a, x - 1, y) + dbd_qpy(a, x, y) + 1) >> 1;
            const int beta = c_betatable[clip3i(qp + beta_offset, 0, 51)];
            if (a.pcm) { nop0 = dbd_pcm(a, x - 1, y); nop1 = dbd_pcm(a, x - 1, y + 4); noq0 = dbd_pcm(a, x, y); noq1 = dbd_pcm(a, x, y + 4); }
            dbd_emit(m, L, 0, 1, x, y, a.width, a.height, beta, bs0 ? dbd_tc(qp, bs0, tc_offset) : 0, bs1 ? dbd_tc(qp, bs1, tc_offset) : 0, nop0, nop1, noq0, noq1);
        } else {
            if (bs0 != 2 && bs1 != 2) return;
            const int qp0 = (dbd_qpy(a, x - 1, y) + dbd_qpy(a, x, y) + 1) >> 1;
            const int yq = y + qy < a.height ? y + qy : y;
            const int qp1 = (dbd_qpy(a, x - 1, yq) + dbd_qpy(a, x, yq) + 1) >> 1;
            if (a.pcm) { nop0 = dbd_pcm(a, x - 1, y); nop1 = dbd_pcm(a, x - 1, y + qy); noq0 = dbd_pcm(a, x, y); noq1 = dbd_pcm(a, x, y + qy); }
            for (int c = 1; c <= 2; c++)
                dbd_emit(m, L, c, 1, x >> hs, y >> vs, pwc, phc, 0, bs0 == 2 ? dbd_chroma_tc(a, qp0, c, tc_offset) : 0, bs1 == 2 ? dbd_chroma_tc(a, qp1, c, tc_offset) : 0,
                         nop0, nop1, noq0, noq1);
        }
    } else {
        if (y == 0) return;
        int cx = (x + sx) >> a.log2_ctb;
        const bool last_ctb = cx >= a.ctb_w;
        if (last_ctb) cx = a.ctb_w - 1;
        const bool first = !last_ctb && ((x + sx) & (ctb - 1)) == 0;      // the loop of CTB cx starts here, one step left of its border
        const int ci = (y >> a.log2_ctb) * a.ctb_w + cx;
        const int cur_beta = a.ctb[2 * ci], cur_tc = a.ctb[2 * ci + 1];
        const int left_beta = first ? a.ctb[2 * (ci - 1)] : cur_beta, left_tc = first ? a.ctb[2 * (ci - 1) + 1] : cur_tc;
        const int bs0 = m.bsh[(size_t)(y >> 2) * m.uw + (x >> 2)];
        const int bs1 = x + qx < a.width ? m.bsh[(size_t)(y >> 2) * m.uw + ((x + qx) >> 2)] : 0;
        int nop0 = 0, nop1 = 0, noq0 = 0, noq1 = 0;
        if (!chroma) {
            if (!(bs0 | bs1)) return;
            const int qp = (dbd_qpy(a, x, y - 1) + dbd_qpy(a, x, y) + 1) >> 1;
            const int beta = c_betatable[clip3i(qp + left_beta, 0, 51)];
            if (a.pcm) { nop0 = dbd_pcm(a, x, y - 1); nop1 = dbd_pcm(a, x + 4, y - 1); noq0 = dbd_pcm(a, x, y); noq1 = dbd_pcm(a, x + 4, y); }
            dbd_emit(m, L, 0, 0, x, y, a.width, a.height, beta, bs0 ? dbd_tc(qp, bs0, cur_tc) : 0, bs1 ? dbd_tc(qp, bs1, cur_tc) : 0, nop0, nop1, noq0, noq1);
        } else {
            if (bs0 != 2 && bs1 != 2) return;
            const int qp0 = bs0 == 2 ? (dbd_qpy(a, x, y - 1) + dbd_qpy(a, x, y) + 1) >> 1 : 0;
            const int xq = x + qx < a.width ? x + qx : x;
            const int qp1 = bs1 == 2 ? (dbd_qpy(a, xq, y - 1) + dbd_qpy(a, xq, y) + 1) >> 1 : 0;
            if (a.pcm) { nop0 = dbd_pcm(a, x, y - 1); nop1 = dbd_pcm(a, x + qx, y - 1); noq0 = dbd_pcm(a, x, y); noq1 = dbd_pcm(a, x + qx, y); }
            for (int c = 1; c <= 2; c++)
                dbd_emit(m, L, c, 0, x >> hs, y >> vs, pwc, phc, 0, bs0 == 2 ? dbd_chroma_tc(a, qp0, c, left_tc) : 0, bs1 == 2 ? dbd_chroma_tc(a, qp1, c, cur_tc) : 0,
                         nop0, nop1, noq0, noq1);
        }
    }
    (void)sy;
}

// B200_DBD_CHECK: the derived grid against the one recorded from the reference's own filter calls
__global__ void __launch_bounds__(256) k_dbd_compare(const uint16_t *__restrict__ derived, const uint16_t *__restrict__ recorded, uint32_t n, uint32_t *gate)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && derived[i] != recorded[i]) { atomicOr(gate + 3, 1u << 30); latch_host(gate, (1u << 30) | (i & 0xfffffff)); }
}
