/*
 * replay_ref.c — TEST INFRASTRUCTURE / CPU BASELINE: executes a B200 work-list blob with the
 * UNMODIFIED reference's own function tables (HEVCDSPContext / HEVCPredContext / VideoDSPContext
 * from oracle/_ref/libohevc_ref.so, built by build_ref.sh from /root/reference).
 *
 * It plays the role of the reference's call sites (hevc.c luma_mc_* / chroma_mc_*,
 * hevc_cabac.c:1868-1949, hevc_filter.c deblocking_filter_CTB / sao_filter_CTB) for a recorded
 * picture: same table functions, same arguments, host planes in AVFrame layout.  Used
 *   - by tests: the frame-level golden result (reference arithmetic, not the restatement),
 *   - by bench.py: `cpu_baseline` / `--impl reference` = the reference's C path on the host cores,
 *     frame-parallel over pthreads like the reference's own frame threads (pthread_frame.c).
 * Compiled against the reference headers where they lie; nothing is copied into the repo.
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include "libavcodec/hevc.h"
#include "libavcodec/hevcdsp.h"
#include "libavcodec/hevcpred.h"
#include "libavcodec/videodsp.h"
#include "libavcodec/get_bits.h"
#include <dlfcn.h>
#include "../include/b200hevc_worklist.h"

extern const uint8_t ff_hevc_pel_weight[65];

typedef struct RefCtx {
    HEVCContext *s;
    HEVCSPS *sps;
    HEVCPPS *pps;
    HEVCLocalContext *lc;
    AVFrame *fr;
    VideoDSPContext vdsp;
    int *zs;
    int bd, B, cfi, W, H;
    int pw[3], ph[3];
    uint64_t stage_ns[5];        /* time spent in the table calls of each stage: MC, residual, intra (+ its residuals), deblock, SAO */
    HEVCFrame refframe;          /* s->ref of a constrained_intra_pred picture: only tab_mvf[].pred_flag is read (hevcpred_template.c:35-40) */
    MvField *mvf;
} RefCtx;

static RefCtx *ref_ctx_new(const B200BlobHeader *h)
{
    RefCtx *c = calloc(1, sizeof(*c));
    c->s = calloc(1, sizeof(HEVCContext)); c->sps = calloc(1, sizeof(HEVCSPS)); c->pps = calloc(1, sizeof(HEVCPPS));
    c->lc = calloc(1, sizeof(HEVCLocalContext)); c->fr = calloc(1, sizeof(AVFrame));
    c->bd = h->bit_depth; c->B = c->bd > 8 ? 2 : 1; c->cfi = h->chroma_format_idc; c->W = h->width; c->H = h->height;
    for (int p = 0; p < 3; p++) b200_plane_dims(c->W, c->H, c->cfi, p, &c->pw[p], &c->ph[p]);
    HEVCContext *s = c->s;
    s->sps = c->sps; s->pps = c->pps; s->HEVClc = c->lc; s->frame = c->fr;
    ff_hevc_pred_init(&s->hpc, c->bd);
    ff_hevc_dsp_init(&s->hevcdsp, c->bd);
    ff_videodsp_init(&c->vdsp, c->bd);
    HEVCSPS *sps = c->sps;
    sps->width = c->W; sps->height = c->H; sps->log2_ctb_size = h->log2_ctb_size; sps->log2_min_tb_size = 2; sps->log2_min_pu_size = 2;
    sps->tb_mask = (1 << (h->log2_ctb_size - 2)) - 1;
    sps->min_pu_width = c->W / 4; sps->min_pu_height = c->H / 4; sps->chroma_array_type = c->cfi; sps->pixel_shift = c->bd > 8;
    sps->hshift[1] = sps->hshift[2] = c->cfi != 3; sps->vshift[1] = sps->vshift[2] = c->cfi == 1;
    sps->bit_depth = c->bd; sps->chroma_format_idc = c->cfi;
    const int tw = sps->tb_mask + 2;
    c->zs = malloc(sizeof(int) * tw * tw);
    for (int i = 0; i < tw * tw; i++) c->zs[i] = -1;     /* every neighbour "earlier in z-scan": the record flags are already final */
    c->pps->min_tb_addr_zs_tab = c->zs; c->pps->min_tb_addr_zs = c->zs + tw + 1;
    return c;
}
static void ref_ctx_free(RefCtx *c)
{
    free(c->mvf); free(c->zs); free(c->fr); free(c->lc); free(c->pps); free(c->sps); free(c->s); free(c);
}

typedef struct HostFrame { uint8_t *p[3]; int stride[3]; } HostFrame;

static uint64_t now_ns(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return (uint64_t)t.tv_sec * 1000000000ull + (uint64_t)t.tv_nsec; }

typedef struct Parked { uint32_t off; const B200TuRec *t; } Parked;
static int cmp_park(const void *a, const void *b)
{
    const Parked *x = a, *y = b;
    return x->off < y->off ? -1 : x->off > y->off;
}

/* ---- K1: inter, following hevc.c:1641-1949 ------------------------------------------------------ */
static void replay_mc(RefCtx *c, const B200BlobHeader *hd, const B200McRec *m, HostFrame *cur, HostFrame *dpb)
{
    HEVCDSPContext *d = &c->s->hevcdsp;
    const int B = c->B, chroma = !!(m->flags & B200_MCF_CHROMA), p = m->plane;
    const int before = chroma ? 1 : 3, after = chroma ? 2 : 4, extra = before + after;
    const int pw = c->pw[p], ph = c->ph[p], w = m->w, h = m->h;
    DECLARE_ALIGNED(16, int16_t, tmp[MAX_PB_SIZE * MAX_PB_SIZE]);
    uint8_t *emu[2] = { c->lc->edge_emu_buffer, c->lc->edge_emu_buffer2 };
    const int nl = (m->flags & B200_MCF_BI) ? 2 : 1;
    uint8_t *src[2]; ptrdiff_t ss[2]; int mx[2], my[2];
    for (int l = 0; l < nl; l++) {
        HostFrame *rf = &dpb[hd->ref_slot[l ? m->ref1 : m->ref0]];
        int sx = l ? m->sx1 : m->sx0, sy = l ? m->sy1 : m->sy0, fr = l ? m->frac1 : m->frac0;
        mx[l] = fr & 15; my[l] = fr >> 4;
        ss[l] = rf->stride[p];
        src[l] = rf->p[p] + (ptrdiff_t)sy * ss[l] + (ptrdiff_t)sx * B;
        if (sx < before || sy < after || sx >= pw - w - after || sy >= ph - h - after) {   /* hevc.c:1660-1663 */
            const int es = EDGE_EMU_BUFFER_STRIDE * B;
            c->vdsp.emulated_edge_mc(emu[l], src[l] - before * ss[l] - before * B, es, ss[l], w + extra, h + extra, sx - before, sy - before, pw, ph);
            src[l] = emu[l] + before * es + before * B; ss[l] = es;
        }
    }
    uint8_t *dst = cur->p[p] + (ptrdiff_t)m->y * cur->stride[p] + m->x * B;
    const ptrdiff_t ds = cur->stride[p];
    const int idx = ff_hevc_pel_weight[w], wt = !!(m->flags & B200_MCF_WEIGHTED);
#define CALL(T) \
    if (nl == 1) { \
        if (!wt) d->put_hevc_##T##_uni[idx][!!my[0]][!!mx[0]](dst, ds, src[0], ss[0], h, mx[0], my[0], w); \
        else d->put_hevc_##T##_uni_w[idx][!!my[0]][!!mx[0]](dst, ds, src[0], ss[0], h, m->denom, m->w0, m->o0, mx[0], my[0], w); \
    } else { \
        d->put_hevc_##T[idx][!!my[0]][!!mx[0]](tmp, MAX_PB_SIZE, src[0], ss[0], h, mx[0], my[0], w); \
        if (!wt) d->put_hevc_##T##_bi[idx][!!my[1]][!!mx[1]](dst, ds, src[1], ss[1], tmp, MAX_PB_SIZE, h, mx[1], my[1], w); \
        else d->put_hevc_##T##_bi_w[idx][!!my[1]][!!mx[1]](dst, ds, src[1], ss[1], tmp, MAX_PB_SIZE, h, m->denom, m->w0, m->w1, m->o0, m->o1, mx[1], my[1], w); \
    }
    if (chroma) { CALL(epel) } else { CALL(qpel) }
#undef CALL
}

/* ---- K2: residual, following hevc_cabac.c:1868-1949 ------------------------------------------------ */
static void replay_tu_residual(RefCtx *c, const B200TuRec *t, int16_t *coeffs)
{
    HEVCDSPContext *d = &c->s->hevcdsp;
    switch (t->kind) {
    case B200_TU_IDCT: d->idct[t->log2 - 2](coeffs, t->col_limit); break;
    case B200_TU_DC:   d->idct_dc[t->log2 - 2](coeffs); break;
    case B200_TU_DST:  d->idct_4x4_luma(coeffs); break;
    case B200_TU_SKIP: d->transform_skip(coeffs, t->log2); break;
    default: break;
    }
    if (t->flags & B200_TUF_RDPCM) d->transform_rdpcm(coeffs, t->log2, !!(t->flags & B200_TUF_RDPCM_VERT));
}

/* ---- K3: intra through the reference's intra_pred[] (hevcpred_template.c:30-344) ---------------------- */
static void replay_intra(RefCtx *c, const B200IntraRec *r)
{
    HEVCLocalContext *lc = c->lc;
    HEVCSPS *sps = c->sps;
    const int hs = sps->hshift[r->plane], vs = sps->vshift[r->plane];
    const int x0 = r->x << hs, y0 = r->y << vs;
    lc->na.cand_up_left = !!(r->flags & B200_INF_UP_LEFT); lc->na.cand_up = !!(r->flags & B200_INF_UP);
    lc->na.cand_up_right = !!(r->flags & B200_INF_UP_RIGHT); lc->na.cand_left = !!(r->flags & B200_INF_LEFT);
    lc->na.cand_bottom_left = !!(r->flags & B200_INF_BOTTOM_LEFT);
    lc->tu.intra_pred_mode = lc->tu.intra_pred_mode_c = r->mode;
    if (r->plane == 0 || c->cfi == 3) sps->spsRext.intra_smoothing_disabled_flag = !(r->flags & B200_INF_FILTER);
    sps->sps_strong_intra_smoothing_enable_flag = !!(r->flags & B200_INF_STRONG);
    const int tw = sps->tb_mask + 2;
    int *cur = &c->pps->min_tb_addr_zs[((y0 >> 2) & sps->tb_mask) * tw + ((x0 >> 2) & sps->tb_mask)];
    *cur = 0;                                      /* current block later than every neighbour (-1) */
    c->s->hpc.intra_pred[r->log2 - 2](c->s, x0, y0, r->plane);
    *cur = -1;
}

static HostFrame frame_alloc(const RefCtx *c, int pad)
{
    HostFrame f;
    for (int p = 0; p < 3; p++) {
        f.stride[p] = ((c->pw[p] + 2 * pad) * c->B + 63) & ~63;
        uint8_t *base = calloc((size_t)f.stride[p] * (c->ph[p] + 2 * pad) + 64, 1);
        f.p[p] = base + (size_t)pad * f.stride[p] + pad * c->B;
    }
    return f;
}
static void frame_free(const RefCtx *c, HostFrame *f, int pad)
{
    for (int p = 0; p < 3; p++) free(f->p[p] - (size_t)pad * f->stride[p] - pad * c->B);
}

/* planes[slot*3+c], strides[slot*3+c]: host planes in the reference's AVFrame layout */
static int execute(RefCtx *c, const uint8_t *blob, uint8_t **planes, const int64_t *strides, int n_slots)
{
    const B200BlobHeader *h = (const B200BlobHeader *)blob;
    HEVCDSPContext *d = &c->s->hevcdsp;
    const int B = c->B;
    HostFrame dpb[64];
    if (n_slots > 64 || h->cur_slot >= n_slots) return -2;
    for (int s = 0; s < n_slots; s++) for (int p = 0; p < 3; p++) { dpb[s].p[p] = planes[3 * s + p]; dpb[s].stride[p] = (int)strides[3 * s + p]; }
    HostFrame *cur = &dpb[h->cur_slot];
    for (int p = 0; p < 3; p++) { c->fr->data[p] = cur->p[p]; c->fr->linesize[p] = cur->stride[p]; }
    /* constrained_intra_pred: the motion field the reference looks the PU types up in, from the blob's bitmap */
    c->pps->constrained_intra_pred_flag = 0;
    if ((h->flags & B200_FRAME_CIP) && h->cip.count >= 4) {
        const uint32_t *cw = (const uint32_t *)(blob + h->cip.off);
        const int pw_ = (int)cw[1], ph_ = (int)cw[2];
        if (h->cip.count < B200_CIP_WORDS(pw_, ph_)) return -6;
        c->sps->log2_min_pu_size = cw[0]; c->sps->min_pu_width = pw_; c->sps->min_pu_height = ph_;
        free(c->mvf);
        c->mvf = calloc((size_t)pw_ * ph_ + 1, sizeof(MvField));
        for (int i = 0; i < pw_ * ph_; i++) c->mvf[i].pred_flag = ((cw[4 + (i >> 5)] >> (i & 31)) & 1) ? PF_INTRA : PF_L0;
        if (!c->s->ref) c->s->ref = &c->refframe;       /* the drop-in run has s->ref = its DPB entry already */
        c->s->ref->tab_mvf = c->mvf;
        c->pps->constrained_intra_pred_flag = 1;
    }

    uint64_t tick = now_ns(), tock;
#define STAGE_END(k) do { tock = now_ns(); c->stage_ns[k] += tock - tick; tick = tock; } while (0)
    const B200McRec *mc = (const B200McRec *)(blob + h->sec[B200_SEC_MC].off);
    for (uint32_t i = 0; i < h->sec[B200_SEC_MC].count; i++) replay_mc(c, h, &mc[i], cur, dpb);
    STAGE_END(0);

    const int16_t *pool = (const int16_t *)(blob + h->sec[B200_SEC_COEFF].off);
    DECLARE_ALIGNED(32, int16_t, coeffs[32 * 32]);
    /* residuals of intra TUs are applied right after their prediction, as hls_transform_unit does (hevc.c:1212-1291):
     * index the parked TU records by pool offset */
    size_t npark = 0;
    for (int sidx = B200_SEC_TU4; sidx <= B200_SEC_TU32; sidx++) {
        const B200TuRec *tu = (const B200TuRec *)(blob + h->sec[sidx].off);
        for (uint32_t i = 0; i < h->sec[sidx].count; i++) npark += !!(tu[i].flags & B200_TUF_PARK);
    }
    Parked *park = malloc((npark + 1) * sizeof(*park));
    npark = 0;
    for (int sidx = B200_SEC_TU4; sidx <= B200_SEC_TU32; sidx++) {
        const B200TuRec *tu = (const B200TuRec *)(blob + h->sec[sidx].off);
        for (uint32_t i = 0; i < h->sec[sidx].count; i++) {
            const B200TuRec *t = &tu[i];
            uint32_t po = 0;
            const int16_t *data = b200_tu_data(t, pool, &po);
            if (t->flags & B200_TUF_PARK) { park[npark].off = po; park[npark++].t = t; continue; }
            const int n = 1 << t->log2, p = t->plane;
            uint8_t *dst = cur->p[p] + (ptrdiff_t)t->y * cur->stride[p] + t->x * B;
            if (t->kind == B200_TU_PCM) {       /* hls_pcm_sample (hevc.c:1587-1623): put_pcm reads the samples from the bitstream */
                uint8_t bits[32 * 32 * 2 + 16];
                GetBitContext gb;
                memset(bits, 0, sizeof(bits));
                for (int k = 0, bp = 0; k < n * n; k++)
                    for (int bit = c->bd - 1; bit >= 0; bit--, bp++)
                        if ((data[k] >> bit) & 1) bits[bp >> 3] |= 0x80 >> (bp & 7);
                init_get_bits(&gb, bits, n * n * c->bd);
                d->put_pcm(dst, cur->stride[p], n, n, &gb, c->bd);
                continue;
            }
            b200_tu_expand(t, data, coeffs);
            replay_tu_residual(c, t, coeffs);
            d->transform_add[t->log2 - 2](dst, coeffs, cur->stride[p]);
        }
    }
    qsort(park, npark, sizeof(*park), cmp_park);
    STAGE_END(1);
    const B200IntraRec *ir = (const B200IntraRec *)(blob + h->sec[B200_SEC_INTRA].off);
    for (uint32_t i = 0; i < h->sec[B200_SEC_INTRA].count; i++) {
        const B200IntraRec *r = &ir[i];
        replay_intra(c, r);
        if (r->resid_off != B200_NO_RESID) {
            size_t lo = 0, hi = npark;
            while (lo + 1 < hi) { size_t mid = (lo + hi) / 2; if (park[mid].off <= r->resid_off) lo = mid; else hi = mid; }
            if (!npark || park[lo].off != r->resid_off) { free(park); return -5; }
            const B200TuRec *t = park[lo].t;
            uint32_t po = 0;
            b200_tu_expand(t, b200_tu_data(t, pool, &po), coeffs);
            replay_tu_residual(c, t, coeffs);
            d->transform_add[r->log2 - 2](cur->p[r->plane] + (ptrdiff_t)r->y * cur->stride[r->plane] + r->x * B, coeffs, cur->stride[r->plane]);
        }
    }
    free(park);
    STAGE_END(2);

    if (h->sec[B200_SEC_DBK].count) {              /* every vertical edge of the picture, then every horizontal one */
        B200DbkLayout L;
        b200_dbk_layout(c->W, c->H, c->cfi, &L);
        const uint16_t *grid = (const uint16_t *)(blob + h->sec[B200_SEC_DBK].off);
        for (int p = 0; p < 3; p++)
            for (int dir = 0; dir < 2; dir++) {
                const uint16_t *g = grid + L.off[p][dir];
                const int gs = L.stride[p][dir];
                const int nx = dir == 0 ? c->pw[p] / 8 : c->pw[p] / 8, ny = c->ph[p] / 8;
                for (int by = 0; by < ny; by++)
                    for (int bx = 0; bx < nx; bx++) {
                        const int x = 8 * bx, y = 8 * by;
                        uint16_t e0, e1;
                        if (dir == 0) { if (!x) continue; e0 = g[(y >> 2) * gs + (x >> 3)]; e1 = g[((y >> 2) + 1) * gs + (x >> 3)]; }
                        else { if (!y) continue; e0 = g[(y >> 3) * gs + (x >> 2)]; e1 = g[(y >> 3) * gs + (x >> 2) + 1]; }
                        if (!((e0 | e1) & B200_DBK_PRESENT)) continue;
                        int tc[2] = { (e0 & B200_DBK_PRESENT) ? B200_DBK_TC(e0) : 0, (e1 & B200_DBK_PRESENT) ? B200_DBK_TC(e1) : 0 };
                        uint8_t no_p[2] = { B200_DBK_NOP(e0), B200_DBK_NOP(e1) }, no_q[2] = { B200_DBK_NOQ(e0), B200_DBK_NOQ(e1) };
                        const int beta = (e0 & B200_DBK_PRESENT) ? B200_DBK_BETA(e0) : B200_DBK_BETA(e1);
                        uint8_t *pix = cur->p[p] + (ptrdiff_t)y * cur->stride[p] + x * B;
                        if (p == 0) (dir == 0 ? d->hevc_v_loop_filter_luma : d->hevc_h_loop_filter_luma)(pix, cur->stride[p], beta, tc, no_p, no_q);
                        else (dir == 0 ? d->hevc_v_loop_filter_chroma : d->hevc_h_loop_filter_chroma)(pix, cur->stride[p], tc, no_p, no_q);
                    }
            }
    }
    STAGE_END(3);
    if (h->sec[B200_SEC_SAO].count) {              /* hevc_filter.c:255-319 with a whole-picture copy as sao_frame */
        const B200SaoRec *sg = (const B200SaoRec *)(blob + h->sec[B200_SEC_SAO].off);
        const int ctb = 1 << h->log2_ctb_size, cw = (c->W + ctb - 1) >> h->log2_ctb_size, ch = (c->H + ctb - 1) >> h->log2_ctb_size;
        HostFrame cp = frame_alloc(c, 1);
        for (int p = 0; p < 3; p++) {
            for (int y = 0; y < c->ph[p]; y++) memcpy(cp.p[p] + (ptrdiff_t)y * cp.stride[p], cur->p[p] + (ptrdiff_t)y * cur->stride[p], (size_t)c->pw[p] * B);
            const int hs = p && c->cfi != 3, vs = p && c->cfi == 1;
            for (int cy = 0; cy < ch; cy++)
                for (int cx = 0; cx < cw; cx++) {
                    const B200SaoRec *r = &sg[(p * ch + cy) * cw + cx];
                    if (r->type == B200_SAO_NONE) continue;
                    const int x0 = (cx << h->log2_ctb_size) >> hs, y0 = (cy << h->log2_ctb_size) >> vs;
                    int w = ctb >> hs, hh = ctb >> vs;
                    if (w > c->pw[p] - x0) w = c->pw[p] - x0;
                    if (hh > c->ph[p] - y0) hh = c->ph[p] - y0;
                    SAOParams sp; memset(&sp, 0, sizeof(sp));
                    sp.band_position[p] = r->param; sp.eo_class[p] = r->param; sp.type_idx[p] = r->type;
                    for (int k = 0; k < 5; k++) sp.offset_val[p][k] = r->offset_val[k];
                    int borders[4] = { r->borders & 1, (r->borders >> 1) & 1, (r->borders >> 2) & 1, (r->borders >> 3) & 1 };
                    uint8_t ve[2] = { r->edges & 1, (r->edges >> 1) & 1 }, he[2] = { (r->edges >> 2) & 1, (r->edges >> 3) & 1 };
                    uint8_t de[4] = { (r->edges >> 4) & 1, (r->edges >> 5) & 1, (r->edges >> 6) & 1, (r->edges >> 7) & 1 };
                    uint8_t *dst = cur->p[p] + (ptrdiff_t)y0 * cur->stride[p] + x0 * B, *src = cp.p[p] + (ptrdiff_t)y0 * cp.stride[p] + x0 * B;
                    if (r->type == B200_SAO_BAND) d->sao_band_filter(dst, src, cur->stride[p], cp.stride[p], &sp, borders, w, hh, p);
                    else d->sao_edge_filter[r->variant ? 1 : 0](dst, src, cur->stride[p], cp.stride[p], &sp, borders, w, hh, p, ve, he, de);
                }
        }
        frame_free(c, &cp, 1);
    }
    STAGE_END(4);
#undef STAGE_END
    return 0;
}

int ref_execute_blob(const uint8_t *blob, uint8_t **planes, const int64_t *strides, int n_slots)
{
    const B200BlobHeader *h = (const B200BlobHeader *)blob;
    if (h->magic != B200_BLOB_MAGIC || h->version != B200_BLOB_VERSION) return -1;
    RefCtx *c = ref_ctx_new(h);
    int rc = execute(c, blob, planes, strides, n_slots);
    ref_ctx_free(c);
    return rc;
}

/* ---- the same call sequence through the B200 drop-in: the tables are re-populated by libb200hevc_shim.so
 * (ff_hevcdsp_init_b200 & co.), the calls are recorded, b200_frame_end runs the GPU, b200_frame_readback returns
 * the picture.  This is the reference-side half of INTEGRATION.md exercised without a bitstream. */
int ref_execute_blob_b200(const uint8_t *blob, uint8_t **planes, const int64_t *strides, int n_slots, const char *shim_path, char *err, int errlen)
{
    const B200BlobHeader *h = (const B200BlobHeader *)blob;
    if (h->magic != B200_BLOB_MAGIC || h->version != B200_BLOB_VERSION || n_slots > 32 || h->cur_slot >= n_slots) return -1;
    void *so = dlopen(shim_path, RTLD_NOW | RTLD_GLOBAL);
    if (!so) { snprintf(err, errlen, "dlopen: %s", dlerror()); return -10; }
    void (*init_dsp)(HEVCDSPContext *, int) = dlsym(so, "ff_hevcdsp_init_b200");
    void (*init_pred)(HEVCPredContext *, int) = dlsym(so, "ff_hevcpred_init_b200");
    void (*init_vdsp)(VideoDSPContext *, int) = dlsym(so, "ff_videodsp_init_b200");
    int (*fbegin)(HEVCContext *) = dlsym(so, "b200_frame_begin");
    int (*fend)(HEVCContext *) = dlsym(so, "b200_frame_end");
    int (*fread)(HEVCContext *, AVFrame *) = dlsym(so, "b200_frame_readback");
    int (*fup)(HEVCContext *, AVFrame *) = dlsym(so, "b200_frame_upload_ref");
    int (*fwait)(HEVCContext *, AVFrame *) = dlsym(so, "b200_output_wait");
    const char *(*ferr)(void) = dlsym(so, "b200_shim_error");
    if (!init_dsp || !init_pred || !init_vdsp || !fbegin || !fend || !fread || !fup || !ferr || !fwait) { snprintf(err, errlen, "shim lacks an entry point"); return -11; }
    RefCtx *c = ref_ctx_new(h);
    init_dsp(&c->s->hevcdsp, c->bd);          /* exactly what ff_hevc_dsp_init would do last (hevcdsp.c:1326) */
    init_pred(&c->s->hpc, c->bd);
    init_vdsp(&c->vdsp, c->bd);
    AVFrame *fr = calloc(n_slots, sizeof(AVFrame));
    for (int s = 0; s < n_slots; s++) {
        for (int p = 0; p < 3; p++) { fr[s].data[p] = planes[3 * s + p]; fr[s].linesize[p] = (int)strides[3 * s + p]; }
        c->s->DPB[s].frame = &fr[s];
    }
    c->s->ref = &c->s->DPB[h->cur_slot];
    c->s->poc = h->poc;
    AVFrame *keep = c->fr;
    c->s->frame = &fr[h->cur_slot];
    c->fr = &fr[h->cur_slot];
    int rc = 0;
    for (int i = 0; i < h->n_ref && !rc; i++) rc = fup(c->s, &fr[h->ref_slot[i]]);
    if (!rc) rc = fbegin(c->s);
    if (!rc) rc = execute(c, blob, planes, strides, n_slots);
    if (!rc) rc = fend(c->s); else fend(c->s);
    if (!rc) rc = fread(c->s, &fr[h->cur_slot]);
    if (!rc) rc = fwait(c->s, &fr[h->cur_slot]);       /* hevc_decode_frame's hook where the picture leaves the decoder */
    if (rc) snprintf(err, errlen, "%s", ferr());
    c->fr = keep;
    free(fr);
    ref_ctx_free(c);
    return rc;
}

/* ---- CPU baseline: `iters` pictures per thread, frame-parallel like the reference's frame threads --------- */
static uint64_t g_stage_ns[5];                   /* thread-time per stage of the last ref_bench() run, summed over the threads */
static pthread_mutex_t g_stage_mu = PTHREAD_MUTEX_INITIALIZER;
void ref_bench_stage_seconds(double out[5]) { for (int k = 0; k < 5; k++) out[k] = 1e-9 * (double)g_stage_ns[k]; }

typedef struct Job { const uint8_t *const *blobs; int n_blobs; uint8_t **planes; const int64_t *strides; int n_slots; int iters; int tid; int rc; pthread_barrier_t *bar; } Job;

static void *worker(void *arg)
{
    Job *j = arg;
    const B200BlobHeader *h0 = (const B200BlobHeader *)j->blobs[0];
    RefCtx *c = ref_ctx_new(h0);
    /* private DPB copy per thread: threads decode independent replicas of the pictures */
    uint8_t *pl[64 * 3]; int64_t st[64 * 3];
    HostFrame own[64];
    for (int s = 0; s < j->n_slots; s++) {
        own[s] = frame_alloc(c, 0);
        for (int p = 0; p < 3; p++) {
            for (int y = 0; y < c->ph[p]; y++) memcpy(own[s].p[p] + (ptrdiff_t)y * own[s].stride[p], j->planes[3 * s + p] + (ptrdiff_t)y * j->strides[3 * s + p], (size_t)c->pw[p] * c->B);
            pl[3 * s + p] = own[s].p[p]; st[3 * s + p] = own[s].stride[p];
        }
    }
    j->rc = 0;
    pthread_barrier_wait(j->bar);               /* setup (private DPB copies) is outside the timed region */
    for (int i = 0; i < j->iters && !j->rc; i++) j->rc = execute(c, j->blobs[(i + j->tid) % j->n_blobs], pl, st, j->n_slots);
    for (int s = 0; s < j->n_slots; s++) frame_free(c, &own[s], 0);
    pthread_mutex_lock(&g_stage_mu);
    for (int k = 0; k < 5; k++) g_stage_ns[k] += c->stage_ns[k];
    pthread_mutex_unlock(&g_stage_mu);
    ref_ctx_free(c);
    return NULL;
}

/* returns seconds of wall time for n_threads x iters pictures, or a negative error */
double ref_bench(const uint8_t *const *blobs, int n_blobs, uint8_t **planes, const int64_t *strides, int n_slots, int n_threads, int iters)
{
    if (n_threads < 1 || n_threads > 256) return -1;
    memset(g_stage_ns, 0, sizeof(g_stage_ns));
    pthread_t th[256]; Job jobs[256];
    struct timespec t0, t1;
    pthread_barrier_t bar;
    pthread_barrier_init(&bar, NULL, n_threads + 1);
    for (int t = 0; t < n_threads; t++) {
        jobs[t] = (Job){ blobs, n_blobs, planes, strides, n_slots, iters, t, 0, &bar };
        pthread_create(&th[t], NULL, worker, &jobs[t]);
    }
    pthread_barrier_wait(&bar);
    clock_gettime(CLOCK_MONOTONIC, &t0);
    int rc = 0;
    for (int t = 0; t < n_threads; t++) { pthread_join(th[t], NULL); if (jobs[t].rc) rc = jobs[t].rc; }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    pthread_barrier_destroy(&bar);
    if (rc) return rc;
    return (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
}
