/*
 * hevcdsp_init_b200.c — the drop-in boundary: re-populates openHEVC's own function tables with
 * recorders that feed libb200hevc.so.  The B200 analogue of libavcodec/x86/hevcdsp_init.c:
 *
 *   ff_hevcdsp_init_b200(HEVCDSPContext *, bit_depth)   <- hook after hevcdsp.c:1326-1327
 *   ff_hevcpred_init_b200(HEVCPredContext *, bit_depth) <- hook after hevcpred.c:84
 *   ff_videodsp_init_b200(VideoDSPContext *, bpc)       <- hook after videodsp.c:51-58
 *   b200_frame_begin / b200_frame_end / b200_frame_readback <- hevc.c:3245 / 3446 / 4145
 *
 * Every installed function has exactly the signature of the slot it replaces (hevcdsp.h:45-105,
 * hevcpred.h:32-40, videodsp.h:66-70) and never touches pixels: it maps the pointers it is given
 * back to (DPB slot, plane, x, y) through the planes registered at frame begin and appends one record.
 * Compiled against the reference's headers where they lie (-I/root/reference); it contains no
 * reference code.  Threading: recorder state is thread-local.  Frame threads (-f 1): every thread owns
 * the picture it decodes, pictures are submitted to the GPU in decode order through a ticket.  Slice / WPP / tile worker
 * threads of ONE picture (-f 2, execute2 jobs, hevc.c:3082): a worker attaches itself to the picture in progress on its
 * first table call, records into its own B200Rec, and b200_frame_end folds the workers into the owner's recorder
 * (b200_rec_merge).  Frame and slice threads combined (-f 4) are rejected: a table call carries no context, so a worker
 * cannot tell which of several pictures in progress it belongs to.  One decoder instance per process.  4:4:4 cross-component prediction is
 * rejected with an error from b200_frame_end.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "libavcodec/hevc.h"
#include "libavcodec/hevcdsp.h"
#include "libavcodec/hevcpred.h"
#include "libavcodec/videodsp.h"
#include "libavcodec/get_bits.h"
#include "b200hevc.h"
#include "b200hevc_tables.h"

#define MAX_REG 33
#define MAX_WORKERS 64
#define MAX_ACTIVE 64

typedef struct RegPlane { const uint8_t *base; ptrdiff_t linesize, size; uint64_t inv; int slot, plane, w, h; } RegPlane;

#include <pthread.h>

/* shared by every decoding thread: the device context, the geometry, and the ticket that keeps pictures in decode
 * order on the GPU (frame threads reach b200_frame_end out of order; hevc_frame_start is called in decode order) */
static struct {
    B200Ctx *ctx;
    B200Config cfg;
    int bd, B, cfi;
    int pw[3], ph[3];
    pthread_mutex_t mu;
    pthread_cond_t cv;
    unsigned next_ticket, turn;
    unsigned gen;                       /* bumped when the context is re-created for a new geometry */
    struct ShimThread *active[MAX_ACTIVE];   /* owners of the pictures between b200_frame_begin and b200_frame_end */
    int n_active;
    /* B200_SHIM_DUMP=<dir>: record only.  Every picture's work list is written to <dir>/pic_NNNNN.blob (decode order) and
     * nothing is sent to a GPU -- no device is needed, the decoder's output pictures stay untouched.  tests/ replay the dumps
     * through the CPU oracle and compare with the unmodified decoder: the recorder + wire format + oracle, end to end. */
    const char *dump_dir;
    int dump_no, configured, env_read;
    int in_flight;                      /* pictures begun whose packet has not ended yet (b200_frame_begin .. b200_frame_readback) */
} G = { .mu = PTHREAD_MUTEX_INITIALIZER, .cv = PTHREAD_COND_INITIALIZER };

/* per thread: the owner of a picture (the thread that runs hevc_frame_start .. the end of decode_nal_unit for it; with
 * frame threads one per picture in flight, pthread_frame.c) or a slice / WPP worker attached to an owner's picture */
typedef struct ShimThread {
    B200Rec *rec;
    unsigned rec_gen;
    unsigned ticket;
    RegPlane reg[MAX_REG * 3];
    int n_reg, reg_hit;
    const uint8_t *cur_base[3]; ptrdiff_t cur_ls[3], cur_size[3]; uint64_t cur_inv[3]; int cur_slot;
    uint8_t ref_slot[16]; int n_ref;
    /* pending in-place transform on the per-thread coefficient scratch (hevc.h:1063) */
    const int16_t *pend_ptr; int pend_kind, pend_flags, pend_col_limit;
    /* first list of a bi-predicted block: put_hevc_*pel() into the caller's tmp[] (hevc.c:1761) */
    const int16_t *first_tmp; B200McRec first;
    /* emulated_edge_mc results, keyed by destination buffer (lc->edge_emu_buffer / edge_emu_buffer2) */
    struct { const uint8_t *buf; ptrdiff_t ls; int slot, plane, x, y; } emu[2];
    int emu_next;
    int err; char errmsg[256];
    int in_frame;                                        /* 0 no picture, 1 owner of the picture in progress, 2 attached worker */
    int poc;
    unsigned frame_seq;                                  /* owner: bumped by every b200_frame_begin */
    struct ShimThread *workers[MAX_WORKERS]; int n_workers;   /* owner: workers that recorded part of this picture */
    struct ShimThread *att; unsigned att_seq;            /* worker: the owner it is attached to */
    int n_tu, n_intra, n_pu, n_dbk, n_sao, frame_no;   /* B200_SHIM_STATS=1: table calls per picture (stderr) */
    /* cross-component prediction (4:4:4 range extension): the owner's decoder context, this thread's local context (found by
     * the coefficient pointer), and the luma transform block the chroma blocks of the same TU refer to */
    HEVCContext *s; int ccp; HEVCLocalContext *lc;
    int counted;                                         /* this thread's picture is part of G.in_flight */
    int pic_cip, pic_tqb;                                /* the picture in progress needs the PU-type / is_pcm hand-over at its end */
    struct { int x, y, log2, kind, flags, cl, parked; uint32_t park; } last_y;
    uint8_t fill_slot[16]; int n_fill;                  /* generate_missing_ref (hevc_refs.c:538): grey references this picture needs */
} ShimThread;
/* One heap block per thread, reached through an 8-byte initial-exec TLS pointer: a plain `static __thread ShimThread` in a
 * shared library costs a __tls_get_addr call per table call (2.4% of the hooked decoder's CPU time plus the PLT), and the
 * struct itself is too large for the static TLS surplus if libOpenHevc is dlopen()ed.  Blocks of threads that exited are
 * recycled, never freed: workers keep pointers to their owner's block (att), and frame_seq must stay monotonic. */
static __thread struct ShimThread *g_self __attribute__((tls_model("initial-exec")));
static pthread_key_t g_key;
static pthread_once_t g_key_once = PTHREAD_ONCE_INIT;
static pthread_mutex_t g_pool_mu = PTHREAD_MUTEX_INITIALIZER;
static ShimThread *g_pool[256]; static int g_pool_n;
static ShimThread g_oom = { .err = B200_ENOMEM, .errmsg = "out of memory (per-thread state of the B200 shim)" };

static void shim_thread_exit(void *p)
{
    ShimThread *t = p;
    if (t == &g_oom) return;
    pthread_mutex_lock(&g_pool_mu);
    t->in_frame = 0;
    if (g_pool_n < 256) g_pool[g_pool_n++] = t;         /* else: leaked, 256 exited threads are already parked */
    pthread_mutex_unlock(&g_pool_mu);
}
static void shim_key_make(void) { pthread_key_create(&g_key, shim_thread_exit); }
static __attribute__((noinline)) ShimThread *shim_self_slow(void)
{
    pthread_once(&g_key_once, shim_key_make);
    pthread_mutex_lock(&g_pool_mu);
    ShimThread *t = g_pool_n ? g_pool[--g_pool_n] : NULL;
    pthread_mutex_unlock(&g_pool_mu);
    if (t) { t->err = 0; t->n_fill = 0; t->n_workers = 0; t->att = NULL; }
    else t = calloc(1, sizeof(*t));
    if (!t) t = &g_oom;
    g_self = t;
    pthread_setspecific(g_key, t);
    return t;
}
static inline ShimThread *shim_self(void)
{
    ShimThread *t = g_self;
    return __builtin_expect(t != NULL, 1) ? t : shim_self_slow();
}
#define g (*shim_self())

static void fail(int code, const char *msg)
{
    if (!g.err) { g.err = code; snprintf(g.errmsg, sizeof(g.errmsg), "%s", msg); }
}
const char *b200_shim_error(void) { return g.err ? g.errmsg : (G.ctx ? b200_last_error(G.ctx) : ""); }

/* A table call on a thread that owns no picture: a slice / WPP / tile worker (execute2 job).  Attach it to the picture
 * in progress: own recorder, own reference table, a copy of the owner's plane registry. */
static int attach_slow(void)
{
    pthread_mutex_lock(&G.mu);
    int rc = 0;
    if (G.n_active != 1) {
        fail(B200_ENOTSUP, G.n_active ? "table call from a worker thread with several pictures in progress (frame + slice threads combined)"
                                      : "table call outside b200_frame_begin / b200_frame_end");
        rc = -1;
    } else {
        ShimThread *o = G.active[0];
        if (g.rec && g.rec_gen != G.gen) { b200_rec_destroy(g.rec); g.rec = NULL; }
        if (!g.rec) {
            g.rec_gen = G.gen;
            if (b200_rec_create(&G.cfg, &g.rec)) { fail(B200_ENOMEM, "b200_rec_create failed (worker)"); rc = -1; }
        }
        if (!rc && o->n_workers == MAX_WORKERS) { fail(B200_ENOTSUP, "too many worker threads"); rc = -1; }
        if (!rc) {
            memcpy(g.reg, o->reg, sizeof(g.reg)); g.n_reg = o->n_reg;
            for (int p = 0; p < 3; p++) { g.cur_base[p] = o->cur_base[p]; g.cur_ls[p] = o->cur_ls[p]; g.cur_size[p] = o->cur_size[p]; g.cur_inv[p] = o->cur_inv[p]; }
            g.cur_slot = o->cur_slot; g.poc = o->poc;
            g.s = o->s; g.ccp = o->ccp; g.lc = NULL; g.last_y.log2 = 0;
            g.n_ref = 0; g.pend_ptr = NULL; g.first_tmp = NULL; g.emu[0].buf = g.emu[1].buf = NULL;
            g.n_tu = g.n_intra = g.n_pu = g.n_dbk = g.n_sao = 0;
            if (b200_rec_begin(g.rec, g.cur_slot, g.poc)) { fail(B200_ESTATE, "b200_rec_begin failed (worker)"); rc = -1; }
            else { o->workers[o->n_workers++] = &g; g.att = o; g.att_seq = o->frame_seq; g.in_frame = 2; }
        }
    }
    pthread_mutex_unlock(&G.mu);
    return rc;
}
static inline int attached(void)
{
    if (g.in_frame == 1) return 1;
    if (g.err) return 0;
    if (g.in_frame == 2 && g.att->in_frame == 1 && g.att->frame_seq == g.att_seq) return 1;
    return attach_slow() == 0;
}

/* ---- pointer -> (slot, plane, x, y) ---------------------------------------------------------------- */
/* offset / linesize without a division (several hundred thousand calls per 4K picture): inv = floor(2^48 / ls) + 1 is exact
 * for off * ls < 2^48, i.e. any plane below 4 GB with a line below 64 KB; inv == 0 (odd geometry) falls back to dividing */
static inline uint64_t row_inverse(ptrdiff_t ls) { return ls > 0 && ls < 65536 ? (1ull << 48) / (uint64_t)ls + 1 : 0; }
static inline uint32_t row_of(uint32_t off, uint32_t ls, uint64_t inv)
{
    return inv ? (uint32_t)(((unsigned __int128)off * inv) >> 48) : off / ls;
}

static int locate_cur(const uint8_t *p, int *plane, int *x, int *y)
{
    if (!attached()) return -1;
    ShimThread *t = &g;
    for (int c = 0; c < 3; c++) {
        const ptrdiff_t off = p - t->cur_base[c];
        if (off >= 0 && off < t->cur_size[c]) {
            const uint32_t o = (uint32_t)off, ls = (uint32_t)t->cur_ls[c], row = row_of(o, ls, t->cur_inv[c]);
            *plane = c; *y = (int)row; *x = (int)((o - row * ls) >> (G.B - 1));
            return 0;
        }
    }
    fail(B200_EINVAL, "destination pointer is not inside the current picture");
    return -1;
}

static int locate_ref(const uint8_t *p, int plane_hint, int *slot, int *x, int *y)
{
    if (!attached()) return -1;
    ShimThread *t = &g;
    for (int e = 0; e < 2; e++)          /* source inside an edge-emulation buffer? (hevc.c:1673) */
        if (t->emu[e].buf && p >= t->emu[e].buf && p < t->emu[e].buf + t->emu[e].ls * (MAX_PB_SIZE + 7)) {
            ptrdiff_t off = p - t->emu[e].buf;
            *slot = t->emu[e].slot; *y = t->emu[e].y + (int)(off / t->emu[e].ls); *x = t->emu[e].x + (int)(off % t->emu[e].ls) / G.B;
            return t->emu[e].plane;
        }
    /* the plane of the previous hit first: consecutive blocks mostly read the same reference picture */
    for (int k = -1; k < t->n_reg; k++) {
        const int i = k < 0 ? t->reg_hit : k;
        if (i >= t->n_reg) continue;
        const RegPlane *r = &t->reg[i];
        ptrdiff_t off = p - r->base;
        if (off >= 0 && off < r->size && (plane_hint < 0 || r->plane == plane_hint)) {
            const uint32_t o = (uint32_t)off, ls = (uint32_t)r->linesize, row = row_of(o, ls, r->inv);
            *slot = r->slot; *y = (int)row; *x = (int)((o - row * ls) >> (G.B - 1));
            t->reg_hit = i;
            return r->plane;
        }
    }
    fail(B200_EINVAL, "source pointer is not inside a registered reference picture");
    return -1;
}

static int ref_index(int slot)
{
    for (int i = 0; i < g.n_ref; i++) if (g.ref_slot[i] == slot) return i;
    if (g.n_ref == 16) { fail(B200_ENOTSUP, "more than 16 reference pictures"); return 0; }
    g.ref_slot[g.n_ref] = (uint8_t)slot;
    return g.n_ref++;
}

/* ---- residual slots --------------------------------------------------------------------------------- */
static void rec_idct(int16_t *c, int col_limit) { g.pend_ptr = c; g.pend_kind = B200_TU_IDCT; g.pend_flags = 0; g.pend_col_limit = col_limit; }
static void rec_idct_dc(int16_t *c) { g.pend_ptr = c; g.pend_kind = B200_TU_DC; g.pend_flags = 0; g.pend_col_limit = 0; }
static void rec_idct_4x4_luma(int16_t *c) { g.pend_ptr = c; g.pend_kind = B200_TU_DST; g.pend_flags = 0; g.pend_col_limit = 0; }
static void rec_transform_skip(int16_t *c, int16_t log2_size) { (void)log2_size; g.pend_ptr = c; g.pend_kind = B200_TU_SKIP; g.pend_flags = 0; g.pend_col_limit = 0; }
static void rec_transform_rdpcm(int16_t *c, int16_t log2_size, int mode)
{
    (void)log2_size;
    if (g.pend_ptr != c) { g.pend_ptr = c; g.pend_kind = B200_TU_BYPASS; g.pend_col_limit = 0; }
    g.pend_flags = B200_TUF_RDPCM | (mode ? B200_TUF_RDPCM_VERT : 0);
}
/* Cross-component prediction (hevc.c:1295-1360, hevc_cabac.c:1942-1948).  Between the chroma block's idct() call and its
 * transform_add() call the reference adds (res_scale_val * luma residual) >> 3 to the chroma coefficients ON THE HOST, reading
 * the luma block it has transformed in place -- which, with the transforms recorded instead of executed, still holds the
 * dequantised luma COEFFICIENTS.  The term it added is therefore known exactly (same arrays, same arithmetic) and is taken
 * out again; the device gets the chroma block as the decoder parsed it, the luma block a second time (parked only), and a
 * record that combines the two residuals there.  The local context (res_scale_val, cross_pf, the luma array) is found through
 * the coefficient pointer: the tables carry no context argument. */
static HEVCLocalContext *local_context_of(const int16_t *coeffs)
{
    ShimThread *t = &g;
    if (t->lc && coeffs >= t->lc->tu.coeffs[0] && coeffs < t->lc->tu.coeffs[0] + 2 * MAX_TB_SIZE * MAX_TB_SIZE) return t->lc;
    HEVCContext *s = t->s;
    for (int i = -1; s && i < MAX_NB_THREADS; i++) {
        HEVCLocalContext *lc = i < 0 ? s->HEVClc : s->HEVClcList[i];
        if (lc && coeffs >= lc->tu.coeffs[0] && coeffs < lc->tu.coeffs[0] + 2 * MAX_TB_SIZE * MAX_TB_SIZE) return t->lc = lc;
    }
    return NULL;
}
static int rec_cross_component(int plane, int x, int y, int log2, int16_t *coeffs, int kind, int flags, int cl, int pending)
{
    ShimThread *t = &g;
    HEVCLocalContext *lc = local_context_of(coeffs);
    if (!lc) { fail(B200_EINVAL, "cross_component_prediction: coefficient block outside every local context"); return -1; }
    if (!lc->tu.cross_pf || !lc->tu.res_scale_val) return 0;                 /* nothing was added on the host */
    if (t->last_y.log2 != log2 || t->last_y.x != x || t->last_y.y != y) { fail(B200_ESTATE, "cross_component_prediction without the luma block of the same TU"); return -1; }
    const int16_t *cy = lc->tu.coeffs[0];
    const int scale = lc->tu.res_scale_val, nn = 1 << (2 * log2);
    int16_t own[MAX_TB_SIZE * MAX_TB_SIZE];
    int has_c = pending;
    for (int i = 0; i < nn; i++) {
        own[i] = (int16_t)(coeffs[i] - ((scale * cy[i]) >> 3));
        has_c |= own[i] != 0;
    }
    int rc = 0;
    uint32_t off_c = 0;
    if (!t->last_y.parked) {
        rc = b200_rec_tu_parked(t->rec, 0, x, y, log2, t->last_y.kind, t->last_y.flags, t->last_y.cl, cy, &t->last_y.park);
        t->last_y.parked = 1;
    }
    if (!rc && has_c) rc = b200_rec_tu_parked(t->rec, plane, x, y, log2, kind, flags, cl, own, &off_c);
    if (!rc) rc = b200_rec_ccp(t->rec, plane, x, y, log2, scale, t->last_y.park, has_c, off_c);
    if (rc) { fail(rc, "recording a cross-component prediction block failed"); return -1; }
    return 1;
}

static void rec_transform_add(uint8_t *dst, int16_t *coeffs, ptrdiff_t stride, int log2)
{
    int plane, x, y;
    (void)stride;
    if (locate_cur(dst, &plane, &x, &y)) return;
    int kind = B200_TU_BYPASS, flags = 0, cl = 0;
    const int pending = g.pend_ptr == coeffs;
    if (pending) { kind = g.pend_kind; flags = g.pend_flags; cl = g.pend_col_limit; }
    g.pend_ptr = NULL;
    g.n_tu++;
    if (g.ccp) {
        if (plane == 0) { g.last_y.x = x; g.last_y.y = y; g.last_y.log2 = log2; g.last_y.kind = kind; g.last_y.flags = flags; g.last_y.cl = cl; g.last_y.parked = 0; }
        else if (rec_cross_component(plane, x, y, log2, coeffs, kind, flags, cl, pending)) return;
    }
    int rc = b200_rec_tu(g.rec, plane, x, y, log2, kind, flags, cl, coeffs, -1);
    if (rc) fail(rc, "b200_rec_tu failed");
}
static void rec_add4(uint8_t *d, int16_t *c, ptrdiff_t s) { rec_transform_add(d, c, s, 2); }
static void rec_add8(uint8_t *d, int16_t *c, ptrdiff_t s) { rec_transform_add(d, c, s, 3); }
static void rec_add16(uint8_t *d, int16_t *c, ptrdiff_t s) { rec_transform_add(d, c, s, 4); }
static void rec_add32(uint8_t *d, int16_t *c, ptrdiff_t s) { rec_transform_add(d, c, s, 5); }

static void rec_put_pcm(uint8_t *dst, ptrdiff_t stride, int width, int height, GetBitContext *gb, int pcm_bit_depth)
{
    int plane, x, y;
    int16_t smp[32 * 32];
    (void)stride;
    if (locate_cur(dst, &plane, &x, &y)) return;
    if (width > 32 || height > 64) { fail(B200_ENOTSUP, "pcm block too large"); return; }
    /* square blocks per plane; 4:2:2 chroma (w x 2w) is recorded as two squares */
    for (int part = 0; part < height / width; part++) {
        for (int i = 0; i < width * width; i++) smp[i] = (int16_t)(get_bits(gb, pcm_bit_depth) << (G.bd - pcm_bit_depth));
        int log2 = 0;
        while ((1 << log2) < width) log2++;
        if (log2 < 2) { fail(B200_ENOTSUP, "pcm block smaller than 4x4"); return; }
        int rc = b200_rec_pcm(g.rec, plane, x, y + part * width, log2, smp);
        if (rc) fail(rc, "b200_rec_pcm failed");
    }
}

/* ---- inter prediction slots --------------------------------------------------------------------------- */
static int mc_fill(B200McRec *m, int list, const uint8_t *src, int mx, int my, int chroma_hint)
{
    int slot, sx, sy;
    int plane = locate_ref(src, -1, &slot, &sx, &sy);
    if (plane < 0) return -1;
    (void)chroma_hint;
    const int ri = ref_index(slot);
    /* positions far outside the picture clamp sample by sample: pre-clamp the origin so that it fits int16 */
    const int pw = G.pw[plane], ph = G.ph[plane];
    if (sx < -80) sx = -80; if (sx > pw + 16) sx = pw + 16;
    if (sy < -80) sy = -80; if (sy > ph + 16) sy = ph + 16;
    if (list == 0) { m->ref0 = (uint8_t)ri; m->sx0 = (int16_t)sx; m->sy0 = (int16_t)sy; m->frac0 = (uint8_t)(mx | (my << 4)); }
    else           { m->ref1 = (uint8_t)ri; m->sx1 = (int16_t)sx; m->sy1 = (int16_t)sy; m->frac1 = (uint8_t)(mx | (my << 4)); }
    return plane;
}

static void mc_first(int16_t *dst, uint8_t *src, int height, intptr_t mx, intptr_t my, int width, int chroma)
{
    memset(&g.first, 0, sizeof(g.first));
    g.first.w = (uint8_t)width; g.first.h = (uint8_t)height;
    g.first.flags = (uint8_t)(B200_MCF_BI | (chroma ? B200_MCF_CHROMA : 0));
    if (mc_fill(&g.first, 0, src, (int)mx, (int)my, chroma) < 0) return;
    g.first_tmp = dst;
}
static void mc_emit(B200McRec *m, uint8_t *dst)
{
    int plane, x, y;
    if (locate_cur(dst, &plane, &x, &y)) return;
    m->plane = (uint8_t)plane; m->x = (uint16_t)x; m->y = (uint16_t)y;
    g.n_pu++;
    int rc = b200_rec_mc(g.rec, m);
    if (rc) fail(rc, "b200_rec_mc failed");
}
static void mc_uni(uint8_t *dst, uint8_t *src, int height, int mx, int my, int width, int chroma, int weighted, int denom, int wx, int ox)
{
    B200McRec m;
    memset(&m, 0, sizeof(m));
    m.w = (uint8_t)width; m.h = (uint8_t)height;
    m.flags = (uint8_t)((chroma ? B200_MCF_CHROMA : 0) | (weighted ? B200_MCF_WEIGHTED : 0));
    m.denom = (uint8_t)denom; m.w0 = (int16_t)wx; m.o0 = (int16_t)ox;
    if (mc_fill(&m, 0, src, mx, my, chroma) < 0) return;
    mc_emit(&m, dst);
}
static void mc_bi(uint8_t *dst, uint8_t *src, int16_t *src2, int height, int mx, int my, int width, int chroma, int weighted,
                  int denom, int wx0, int wx1, int ox0, int ox1)
{
    if (src2 != g.first_tmp) { fail(B200_ESTATE, "put_hevc_*_bi without the matching first-list call"); return; }
    B200McRec m = g.first;
    g.first_tmp = NULL;
    if (m.w != width || m.h != height) { fail(B200_ESTATE, "bi-prediction block size mismatch"); return; }
    m.flags |= weighted ? B200_MCF_WEIGHTED : 0;
    m.denom = (uint8_t)denom; m.w0 = (int16_t)wx0; m.w1 = (int16_t)wx1; m.o0 = (int16_t)ox0; m.o1 = (int16_t)ox1;
    if (mc_fill(&m, 1, src, mx, my, chroma) < 0) return;
    mc_emit(&m, dst);
}

#define MC_FAMILY(T, CH) \
static void rec_##T(int16_t *dst, ptrdiff_t ds, uint8_t *src, ptrdiff_t ss, int h, intptr_t mx, intptr_t my, int w) \
{ (void)ds; (void)ss; mc_first(dst, src, h, mx, my, w, CH); } \
static void rec_##T##_uni(uint8_t *dst, ptrdiff_t ds, uint8_t *src, ptrdiff_t ss, int h, intptr_t mx, intptr_t my, int w) \
{ (void)ds; (void)ss; mc_uni(dst, src, h, (int)mx, (int)my, w, CH, 0, 0, 0, 0); } \
static void rec_##T##_uni_w(uint8_t *dst, ptrdiff_t ds, uint8_t *src, ptrdiff_t ss, int h, int denom, int wx, int ox, intptr_t mx, intptr_t my, int w) \
{ (void)ds; (void)ss; mc_uni(dst, src, h, (int)mx, (int)my, w, CH, 1, denom, wx, ox); } \
static void rec_##T##_bi(uint8_t *dst, ptrdiff_t ds, uint8_t *src, ptrdiff_t ss, int16_t *src2, ptrdiff_t s2, int h, intptr_t mx, intptr_t my, int w) \
{ (void)ds; (void)ss; (void)s2; mc_bi(dst, src, src2, h, (int)mx, (int)my, w, CH, 0, 0, 0, 0, 0, 0); } \
static void rec_##T##_bi_w(uint8_t *dst, ptrdiff_t ds, uint8_t *src, ptrdiff_t ss, int16_t *src2, ptrdiff_t s2, int h, int denom, int wx0, int wx1, int ox0, int ox1, intptr_t mx, intptr_t my, int w) \
{ (void)ds; (void)ss; (void)s2; mc_bi(dst, src, src2, h, (int)mx, (int)my, w, CH, 1, denom, wx0, wx1, ox0, ox1); }
MC_FAMILY(qpel, 0)
MC_FAMILY(epel, 1)

/* vdsp.emulated_edge_mc: remember where the window really comes from; nothing is copied (the device clamps) */
static void rec_emulated_edge_mc(uint8_t *buf, const uint8_t *src, ptrdiff_t buf_linesize, ptrdiff_t src_linesize,
                                 int block_w, int block_h, int src_x, int src_y, int w, int h)
{
    (void)block_w; (void)block_h; (void)w; (void)h;
    if (!attached()) return;
    const uint8_t *base = src - ((ptrdiff_t)src_y * src_linesize + (ptrdiff_t)src_x * G.B);
    for (int i = 0; i < g.n_reg; i++)
        if (g.reg[i].base == base && g.reg[i].linesize == src_linesize) {
            int e = (g.emu[0].buf == buf) ? 0 : (g.emu[1].buf == buf) ? 1 : (g.emu_next++ & 1);
            g.emu[e].buf = buf; g.emu[e].ls = buf_linesize; g.emu[e].slot = g.reg[i].slot; g.emu[e].plane = g.reg[i].plane;
            g.emu[e].x = src_x; g.emu[e].y = src_y;
            return;
        }
    fail(B200_EINVAL, "emulated_edge_mc source does not belong to a registered reference picture");
}

/* ---- deblocking slots ------------------------------------------------------------------------------------ */
static void rec_dbk(uint8_t *pix, int vertical, int beta, int *tc, uint8_t *no_p, uint8_t *no_q)
{
    int plane, x, y;
    if (locate_cur(pix, &plane, &x, &y)) return;
    g.n_dbk++;
    int rc = b200_rec_deblock(g.rec, plane, vertical, x, y, beta, tc, no_p, no_q);
    if (rc) fail(rc, "b200_rec_deblock failed");
}
static void rec_h_luma(uint8_t *p, ptrdiff_t s, int beta, int *tc, uint8_t *np, uint8_t *nq) { (void)s; rec_dbk(p, 0, beta, tc, np, nq); }
static void rec_v_luma(uint8_t *p, ptrdiff_t s, int beta, int *tc, uint8_t *np, uint8_t *nq) { (void)s; rec_dbk(p, 1, beta, tc, np, nq); }
static void rec_h_chroma(uint8_t *p, ptrdiff_t s, int *tc, uint8_t *np, uint8_t *nq) { (void)s; rec_dbk(p, 0, 0, tc, np, nq); }
static void rec_v_chroma(uint8_t *p, ptrdiff_t s, int *tc, uint8_t *np, uint8_t *nq) { (void)s; rec_dbk(p, 1, 0, tc, np, nq); }

/* ---- SAO slots (dst = picture, src = the host's sao_frame copy, which the device does not need) ---------- */
static void rec_sao(uint8_t *dst, SAOParams *sao, int *borders, int c_idx, int type, int variant, uint8_t *ve, uint8_t *he, uint8_t *de)
{
    int plane, x, y;
    if (locate_cur(dst, &plane, &x, &y)) return;
    B200SaoRec r;
    memset(&r, 0, sizeof(r));
    r.type = (uint8_t)type;
    r.param = type == B200_SAO_BAND ? sao->band_position[c_idx] : sao->eo_class[c_idx];
    r.borders = (uint8_t)((borders[0] ? 1 : 0) | (borders[1] ? 2 : 0) | (borders[2] ? 4 : 0) | (borders[3] ? 8 : 0));
    if (variant) r.edges = (uint8_t)((ve[0] ? 1 : 0) | (ve[1] ? 2 : 0) | (he[0] ? 4 : 0) | (he[1] ? 8 : 0) | (de[0] ? 16 : 0) | (de[1] ? 32 : 0) | (de[2] ? 64 : 0) | (de[3] ? 128 : 0));
    r.variant = (uint8_t)variant;
    for (int k = 0; k < 5; k++) r.offset_val[k] = sao->offset_val[c_idx][k];
    g.n_sao++;
    int rc = b200_rec_sao(g.rec, plane, x, y, &r);
    if (rc) fail(rc, "b200_rec_sao failed");
}
static void rec_sao_band(uint8_t *dst, uint8_t *src, ptrdiff_t sd, ptrdiff_t ss, SAOParams *sao, int *borders, int w, int h, int c_idx)
{ (void)src; (void)sd; (void)ss; (void)w; (void)h; rec_sao(dst, sao, borders, c_idx, B200_SAO_BAND, 0, NULL, NULL, NULL); }
static void rec_sao_edge0(uint8_t *dst, uint8_t *src, ptrdiff_t sd, ptrdiff_t ss, SAOParams *sao, int *borders, int w, int h, int c_idx, uint8_t *ve, uint8_t *he, uint8_t *de)
{ (void)src; (void)sd; (void)ss; (void)w; (void)h; rec_sao(dst, sao, borders, c_idx, B200_SAO_EDGE, 0, ve, he, de); }
static void rec_sao_edge1(uint8_t *dst, uint8_t *src, ptrdiff_t sd, ptrdiff_t ss, SAOParams *sao, int *borders, int w, int h, int c_idx, uint8_t *ve, uint8_t *he, uint8_t *de)
{ (void)src; (void)sd; (void)ss; (void)w; (void)h; rec_sao(dst, sao, borders, c_idx, B200_SAO_EDGE, 1, ve, he, de); }

/* ---- intra slot: the availability derivation of hevcpred_template.c:82-109, then one record ------------------ */
static void rec_intra(HEVCContext *s, int x0, int y0, int log2_size, int c_idx)
{
    HEVCLocalContext *lc = s->HEVClc;
    const HEVCSPS *sps = s->sps;
    if (!attached()) return;
    const int hshift = sps->hshift[c_idx], vshift = sps->vshift[c_idx];
    const int size = 1 << log2_size;
    const int size_l_h = size << hshift, size_l_v = size << vshift;
    const int tbs_h = size_l_h >> sps->log2_min_tb_size, tbs_v = size_l_v >> sps->log2_min_tb_size;
    const int x_tb = (x0 >> sps->log2_min_tb_size) & sps->tb_mask, y_tb = (y0 >> sps->log2_min_tb_size) & sps->tb_mask;
    const int zstride = sps->tb_mask + 2;
    const int *zs = s->pps->min_tb_addr_zs;
    const int cur = zs[y_tb * zstride + x_tb];
    const int bl = lc->na.cand_bottom_left && cur > zs[((y_tb + tbs_v) & sps->tb_mask) * zstride + x_tb - 1];
    const int ur = lc->na.cand_up_right && cur > zs[(y_tb - 1) * zstride + ((x_tb + tbs_h) & sps->tb_mask)];
    int bls = (FFMIN(y0 + 2 * size_l_v, sps->height) - (y0 + size_l_v)) >> vshift;
    int trs = (FFMIN(x0 + 2 * size_l_h, sps->width) - (x0 + size_l_h)) >> hshift;
    int flags = (lc->na.cand_up_left ? B200_INF_UP_LEFT : 0) | (lc->na.cand_up ? B200_INF_UP : 0) | (ur ? B200_INF_UP_RIGHT : 0) |
                (lc->na.cand_left ? B200_INF_LEFT : 0) | (bl ? B200_INF_BOTTOM_LEFT : 0);
    if (!sps->spsRext.intra_smoothing_disabled_flag && (c_idx == 0 || sps->chroma_array_type == 3)) flags |= B200_INF_FILTER;
    if (sps->sps_strong_intra_smoothing_enable_flag) flags |= B200_INF_STRONG;
    if (ur && trs <= 0) { flags &= ~B200_INF_UP_RIGHT; if (!(flags & B200_INF_UP)) fail(B200_ENOTSUP, "up-right available with no sample inside the picture"); }
    if (bl && bls <= 0) { flags &= ~B200_INF_BOTTOM_LEFT; if (!(flags & B200_INF_LEFT)) fail(B200_ENOTSUP, "bottom-left available with no sample inside the picture"); }
    const int mode = c_idx ? lc->tu.intra_pred_mode_c : lc->tu.intra_pred_mode;
    g.n_intra++;
    int rc = b200_rec_intra(g.rec, c_idx, x0 >> hshift, y0 >> vshift, log2_size, mode, flags, trs, bls);
    if (rc) fail(rc, "b200_rec_intra failed");
}
static void rec_intra_2(HEVCContext *s, int x0, int y0, int c) { rec_intra(s, x0, y0, 2, c); }
static void rec_intra_3(HEVCContext *s, int x0, int y0, int c) { rec_intra(s, x0, y0, 3, c); }
static void rec_intra_4(HEVCContext *s, int x0, int y0, int c) { rec_intra(s, x0, y0, 4, c); }
static void rec_intra_5(HEVCContext *s, int x0, int y0, int c) { rec_intra(s, x0, y0, 5, c); }

/* ---- table installation ------------------------------------------------------------------------------------- */
/* The host copies of the picture are dead once the tables record instead of computing: the one place where the reference
 * still moves whole CTBs of them around (copy_CTB in sao_filter_CTB, hevc_filter.c:151-161, 269, 305; 6 % of the hooked
 * decoder's CPU time on a dense 4K stream, 24 % on a lightly coded one) may skip the work. */
static int g_tables_installed;          /* 0 -> 1 once, by whichever thread initialises its tables first */
int b200_host_pixels_unused(void) { return __atomic_load_n(&g_tables_installed, __ATOMIC_RELAXED); }

void ff_hevcdsp_init_b200(HEVCDSPContext *c, const int bit_depth)
{
    (void)bit_depth;
    __atomic_store_n(&g_tables_installed, 1, __ATOMIC_RELAXED);
    c->put_pcm = rec_put_pcm;
    c->transform_add[0] = rec_add4; c->transform_add[1] = rec_add8; c->transform_add[2] = rec_add16; c->transform_add[3] = rec_add32;
    c->transform_skip = rec_transform_skip;
    c->transform_rdpcm = rec_transform_rdpcm;
    c->idct_4x4_luma = rec_idct_4x4_luma;
    for (int i = 0; i < 4; i++) { c->idct[i] = rec_idct; c->idct_dc[i] = rec_idct_dc; }
    c->sao_band_filter = rec_sao_band;
    c->sao_edge_filter[0] = rec_sao_edge0; c->sao_edge_filter[1] = rec_sao_edge1;
    for (int i = 0; i < 10; i++)
        for (int j = 0; j < 2; j++)
            for (int k = 0; k < 2; k++) {
                c->put_hevc_qpel[i][j][k] = rec_qpel; c->put_hevc_qpel_uni[i][j][k] = rec_qpel_uni; c->put_hevc_qpel_uni_w[i][j][k] = rec_qpel_uni_w;
                c->put_hevc_qpel_bi[i][j][k] = rec_qpel_bi; c->put_hevc_qpel_bi_w[i][j][k] = rec_qpel_bi_w;
                c->put_hevc_epel[i][j][k] = rec_epel; c->put_hevc_epel_uni[i][j][k] = rec_epel_uni; c->put_hevc_epel_uni_w[i][j][k] = rec_epel_uni_w;
                c->put_hevc_epel_bi[i][j][k] = rec_epel_bi; c->put_hevc_epel_bi_w[i][j][k] = rec_epel_bi_w;
            }
    c->hevc_h_loop_filter_luma = rec_h_luma; c->hevc_v_loop_filter_luma = rec_v_luma;
    c->hevc_h_loop_filter_chroma = rec_h_chroma; c->hevc_v_loop_filter_chroma = rec_v_chroma;
    c->hevc_h_loop_filter_luma_c = rec_h_luma; c->hevc_v_loop_filter_luma_c = rec_v_luma;
    c->hevc_h_loop_filter_chroma_c = rec_h_chroma; c->hevc_v_loop_filter_chroma_c = rec_v_chroma;
}

void ff_hevcpred_init_b200(HEVCPredContext *c, const int bit_depth)
{
    (void)bit_depth;
    c->intra_pred[0] = rec_intra_2; c->intra_pred[1] = rec_intra_3; c->intra_pred[2] = rec_intra_4; c->intra_pred[3] = rec_intra_5;
    /* pred_planar / pred_dc / pred_angular are only reached through intra_pred: left untouched */
}

void ff_videodsp_init_b200(VideoDSPContext *c, int bpc)
{
    (void)bpc;
    c->emulated_edge_mc = rec_emulated_edge_mc;
}

/* ---- frame life cycle ----------------------------------------------------------------------------------------- */
static int ensure_ctx(const HEVCContext *s)       /* called with G.mu held */
{
    const HEVCSPS *sps = s->sps;
    if (!G.env_read) { G.dump_dir = getenv("B200_SHIM_DUMP"); G.env_read = 1; }     /* once, under G.mu: read without it afterwards */
    if (!((G.ctx || (G.dump_dir && G.configured)) && G.cfg.width == sps->width && G.cfg.height == sps->height && G.cfg.bit_depth == sps->bit_depth &&
          G.cfg.chroma_format_idc == sps->chroma_format_idc && G.cfg.log2_ctb_size == (int)sps->log2_ctb_size)) {
        /* New geometry (a new SPS, hence an IRAP picture: nothing older is referenced any more).  With frame threads older pictures
         * may still be parsing on other threads against the old context, plane sizes and sample width: wait until their packets
         * have ended (they do not depend on this thread), then switch. */
        while (G.in_flight > (g.counted ? 1 : 0)) pthread_cond_wait(&G.cv, &G.mu);      /* (this thread's own packet does not count) */
        G.gen++;
        if (G.ctx) { b200_ctx_destroy(G.ctx); G.ctx = NULL; }
        memset(&G.cfg, 0, sizeof(G.cfg));
        const char *dev = getenv("B200_DEVICE");
        G.cfg.device = dev ? atoi(dev) : 0;
        G.cfg.width = sps->width; G.cfg.height = sps->height; G.cfg.chroma_format_idc = sps->chroma_format_idc;
        G.cfg.bit_depth = sps->bit_depth; G.cfg.log2_ctb_size = sps->log2_ctb_size;
        G.cfg.n_slots = 32;                /* == FF_ARRAY_ELEMS(s->DPB), hevc.h:1207 */
        G.cfg.n_arenas = 8;
        if (!G.dump_dir) {
            int rc = b200_ctx_create(&G.cfg, &G.ctx);
            if (rc) { fail(rc, b200_last_error(NULL)); return rc; }
        }
        G.configured = 1;
        G.bd = sps->bit_depth; G.B = G.bd > 8 ? 2 : 1; G.cfi = sps->chroma_format_idc;
        for (int p = 0; p < 3; p++) b200_plane_dims(sps->width, sps->height, G.cfi, p, &G.pw[p], &G.ph[p]);
    }
    if (g.rec && g.rec_gen != G.gen) { b200_rec_destroy(g.rec); g.rec = NULL; }
    if (!g.rec) {
        g.rec_gen = G.gen;
        int rc = b200_rec_create(&G.cfg, &g.rec);
        if (rc) { fail(rc, "b200_rec_create failed"); return rc; }
    }
    return 0;
}

static void ticket_release(void)                   /* let the next picture (in decode order) submit */
{
    pthread_mutex_lock(&G.mu);
    while (G.turn != g.ticket) pthread_cond_wait(&G.cv, &G.mu);
    G.turn++;
    pthread_cond_broadcast(&G.cv);
    pthread_mutex_unlock(&G.mu);
}

static void deactivate(void)                       /* the picture is no longer open for worker threads */
{
    pthread_mutex_lock(&G.mu);
    g.in_frame = 0;
    for (int i = 0; i < G.n_active; i++)
        if (G.active[i] == &g) { G.active[i] = G.active[--G.n_active]; break; }
    pthread_mutex_unlock(&G.mu);
}

static void finish_abandoned(HEVCContext *s);
static void picture_flags(HEVCContext *s);

int b200_frame_begin(HEVCContext *s)
{
    if (g.err) return g.err;
    if (g.in_frame == 1) finish_abandoned(s);                  /* previous picture of this thread was abandoned */
    if (g.err) return g.err;
    g.in_frame = 0;
    picture_flags(s);
    /* cross-component prediction (4:4:4): host arithmetic between two table calls, undone and redone on the device (rec_cross_component) */
    g.s = s; g.lc = NULL; g.last_y.log2 = 0;
    g.ccp = s->sps->chroma_array_type == 3 && s->pps->cross_component_prediction_enabled_flag;
    pthread_mutex_lock(&G.mu);
    const int erc = ensure_ctx(s);
    if (!erc) {
        g.ticket = G.next_ticket++;
        if (!g.counted) { G.in_flight++; g.counted = 1; }      /* once per packet, however many pictures it starts */
    }
    pthread_mutex_unlock(&G.mu);
    if (erc) return g.err;
    g.n_reg = 0;
    for (int i = 0; i < 32; i++) {
        AVFrame *f = s->DPB[i].frame;
        if (!f || !f->data[0]) continue;
        for (int p = 0; p < 3; p++) {
            RegPlane *r = &g.reg[g.n_reg++];
            r->base = f->data[p]; r->linesize = f->linesize[p]; r->slot = i; r->plane = p; r->w = G.pw[p]; r->h = G.ph[p];
            r->size = r->linesize * r->h; r->inv = row_inverse(r->linesize);
        }
    }
    g.cur_slot = (int)(s->ref - s->DPB);
    for (int p = 0; p < 3; p++) {
        g.cur_base[p] = s->frame->data[p]; g.cur_ls[p] = s->frame->linesize[p];
        g.cur_size[p] = g.cur_ls[p] * G.ph[p]; g.cur_inv[p] = row_inverse(g.cur_ls[p]);
    }
    g.n_ref = 0; g.pend_ptr = NULL; g.first_tmp = NULL; g.emu[0].buf = g.emu[1].buf = NULL;
    g.poc = s->poc; g.n_workers = 0;
    int rc = b200_rec_begin(g.rec, g.cur_slot, s->poc);
    if (rc) { fail(rc, "b200_rec_begin failed"); ticket_release(); return rc; }
    pthread_mutex_lock(&G.mu);
    g.frame_seq++;
    g.in_frame = 1;
    if (G.n_active < MAX_ACTIVE) G.active[G.n_active++] = &g;
    pthread_mutex_unlock(&G.mu);
    return 0;
}


static int frame_end_of(HEVCContext *s, HEVCFrame *ref)
{
    if (g.in_frame != 1) return g.err ? g.err : B200_ESTATE;
    deactivate();                                   /* all execute2 jobs of the picture have returned (hevc.c:3087) */
    int mrc = b200_rec_set_refs(g.rec, g.ref_slot, g.n_ref);
    for (int i = 0; i < g.n_workers; i++) {         /* fold the slice / WPP workers' lists into this recorder */
        ShimThread *w = g.workers[i];
        if (w->err) { fail(w->err, w->errmsg); w->err = 0; }
        if (!mrc) mrc = b200_rec_set_refs(w->rec, w->ref_slot, w->n_ref);
        if (!mrc) mrc = b200_rec_merge(g.rec, w->rec);
        g.n_tu += w->n_tu; g.n_intra += w->n_intra; g.n_pu += w->n_pu; g.n_dbk += w->n_dbk; g.n_sao += w->n_sao;
        w->in_frame = 0;
    }
    if (mrc) fail(mrc, "merging the worker threads' work lists failed");
    if (!g.err && g.pic_cip && ref && ref->tab_mvf) {
        /* the device applies the constrained-intra rules itself (hevcpred_template.c:116-249): hand it the PU types */
        const int pw = s->sps->min_pu_width, ph = s->sps->min_pu_height;
        uint8_t *map = malloc((size_t)pw * ph);
        if (!map) fail(B200_ENOMEM, "constrained_intra_pred map");
        else {
            for (int i = 0; i < pw * ph; i++) map[i] = ref->tab_mvf[i].pred_flag == PF_INTRA;
            int crc = b200_rec_set_cip(g.rec, s->sps->log2_min_pu_size, pw, ph, map);
            if (crc) fail(crc, "b200_rec_set_cip failed");
            free(map);
        }
    }
    if (!g.err && g.pic_tqb && s->is_pcm) {
        /* restore_tqb_pixels (hevc_filter.c:163-193) is pixel work outside the tables: the device redoes it from is_pcm[] */
        int crc = b200_rec_set_tqb(g.rec, s->sps->log2_min_pu_size, s->sps->min_pu_width, s->sps->min_pu_height, s->is_pcm);
        if (crc) fail(crc, "b200_rec_set_tqb failed");
    }
    if (g.err) { ticket_release(); return g.err; }
    if (getenv("B200_SHIM_STATS"))
        fprintf(stderr, "b200 picture %d: intra_pred %d transform_add %d mc %d deblock %d sao %d\n", g.frame_no, g.n_intra, g.n_tu, g.n_pu, g.n_dbk, g.n_sao);
    g.frame_no++; g.n_tu = g.n_intra = g.n_pu = g.n_dbk = g.n_sao = 0;
    const void *blob; uint64_t n;
    int rc = b200_rec_finish(g.rec, &blob, &n);
    /* pictures enter the compute stream in decode order, whatever order the frame threads finish parsing in */
    pthread_mutex_lock(&G.mu);
    while (G.turn != g.ticket) pthread_cond_wait(&G.cv, &G.mu);
    for (int i = 0; i < g.n_fill && !rc; i++) {      /* grey reference pictures, in decode order with everything else */
        const int grey = 1 << (G.bd - 1);
        if (G.dump_dir && !strcmp(G.dump_dir, "-")) continue;
        if (G.dump_dir) {
            char path[1024];
            snprintf(path, sizeof(path), "%s/pic_%05d.fill", G.dump_dir, G.dump_no);
            FILE *f = fopen(path, "a");
            if (f) { fprintf(f, "%d %d\n", g.fill_slot[i], grey); fclose(f); } else rc = B200_EINVAL;
        } else rc = b200_slot_fill(G.ctx, g.fill_slot[i], grey);
    }
    g.n_fill = 0;
    if (!rc && G.dump_dir && !strcmp(G.dump_dir, "-")) G.dump_no++;     /* "-": record and drop (host-side timing without a device) */
    else if (!rc && G.dump_dir) {
        char path[1024];
        snprintf(path, sizeof(path), "%s/pic_%05d.blob", G.dump_dir, G.dump_no++);
        FILE *f = fopen(path, "wb");
        if (!f || fwrite(blob, 1, (size_t)n, f) != (size_t)n) { fail(B200_EINVAL, "B200_SHIM_DUMP: cannot write the work list"); rc = B200_EINVAL; }
        if (f) fclose(f);
    } else if (!rc) rc = b200_frame_submit(G.ctx, blob, n);
    G.turn++;
    pthread_cond_broadcast(&G.cv);
    pthread_mutex_unlock(&G.mu);
    if (!rc && !G.dump_dir) rc = b200_wait_uploads(G.ctx);  /* this thread's recorder memory is reused by its next picture */
    if (rc) fail(rc, G.ctx ? b200_last_error(G.ctx) : "frame_end failed");
    return rc;
}
static void picture_flags(HEVCContext *s)
{
    g.pic_cip = s->pps->constrained_intra_pred_flag;
    g.pic_tqb = s->sps->sao_enabled && (s->pps->transquant_bypass_enable_flag || (s->sps->pcm.loop_filter_disable_flag && s->sps->pcm_enabled_flag));
}
int b200_frame_end(HEVCContext *s)
{
    picture_flags(s);                               /* the parameter sets of the picture that ends (at b200_frame_begin of the NEXT one they may have changed) */
    return frame_end_of(s, s->ref);
}

static int readback_into(HEVCContext *s, AVFrame *frame);
/* A picture that was begun but never ended (corrupt slice data).  It is finished with what was recorded -- the part the
 * reference has reconstructed as well -- and copied into its host frame, which the decoder will still output. */
static void finish_abandoned(HEVCContext *s)
{
    HEVCFrame *old = g.cur_slot >= 0 && g.cur_slot < 32 ? &s->DPB[g.cur_slot] : NULL;
    if (!old || !old->frame || !old->frame->data[0] || old->frame->data[0] != g.cur_base[0]) {      /* the frame is gone: nothing to show */
        deactivate(); ticket_release();
        return;
    }
    if (!frame_end_of(s, old)) readback_into(s, old->frame);
}

/* hevc_refs.c:538-606 generate_missing_ref: a reference the stream does not contain is replaced by a grey picture the host
 * fills with memset; the device slot must hold the same.  Called while the RPS of the NEXT picture of this thread is set up
 * (before its b200_frame_begin), executed in decode order at that picture's b200_frame_end. */
int b200_frame_fill(HEVCContext *s, HEVCFrame *frame)
{
    if (g.err) return g.err;
    const int slot = (int)(frame - s->DPB);
    if (slot < 0 || slot >= 32) { fail(B200_EINVAL, "generate_missing_ref: frame is not in the DPB"); return g.err; }
    if (g.n_fill == 16) { fail(B200_ENOTSUP, "more than 16 missing reference pictures"); return g.err; }
    g.fill_slot[g.n_fill++] = (uint8_t)slot;
    return 0;
}

static int packet_end(HEVCContext *s, AVFrame *frame);
int b200_frame_readback(HEVCContext *s, AVFrame *frame)
{
    const int rc = packet_end(s, frame);
    if (g.counted) {                                /* this thread's picture no longer needs the context it was begun with */
        pthread_mutex_lock(&G.mu);
        g.counted = 0;
        G.in_flight--;
        pthread_cond_broadcast(&G.cv);
        pthread_mutex_unlock(&G.mu);
    }
    return rc;
}
static int packet_end(HEVCContext *s, AVFrame *frame)
{
    if (g.err) return g.err;
    if (!frame) {
        /* The packet ended without a complete picture.  If one was begun, it was abandoned (corrupt slice data: hls_slice_data came
         * back short of the picture, hevc.c:3444-3446, and decode_nal_unit swallowed the error).  It is finished here with what was
         * recorded -- the part the reference has reconstructed as well -- for two reasons: the decoder may hand the frame to the
         * application as soon as this call returns (low-delay output), and with frame threads the pictures behind it wait for its
         * ticket in b200_frame_end while this thread only gets its next packet after one of THEM has been delivered: a dead lock
         * unless the picture is closed now.  (Packets are access units: a picture never continues in the next packet.) */
        if (!(g.in_frame == 1 && s->ref)) return 0;
        const int rc = b200_frame_end(s);
        if (rc) return rc;
        frame = s->ref->frame;
    }
    return readback_into(s, frame);
}
static int readback_into(HEVCContext *s, AVFrame *frame)
{
    if (G.dump_dir) return 0;                 /* record-only run: there is no device picture */
    int slot = -1;
    for (int i = 0; i < 32; i++) if (s->DPB[i].frame && s->DPB[i].frame->data[0] == frame->data[0]) slot = i;
    if (slot < 0) { fail(B200_EINVAL, "readback of a frame that is not in the DPB"); return g.err; }
    void *planes[3] = { frame->data[0], frame->data[1], frame->data[2] };
    int64_t strides[3] = { frame->linesize[0], frame->linesize[1], frame->linesize[2] };
    int rc = b200_slot_readback(G.ctx, slot, planes, strides);
    if (!rc) rc = b200_sync(G.ctx);
    if (rc) fail(rc, b200_last_error(G.ctx));
    return rc;
}

/* reference pictures that exist only on the host (e.g. produced before the hook was active) */
int b200_frame_upload_ref(HEVCContext *s, AVFrame *frame)
{
    pthread_mutex_lock(&G.mu);
    const int erc = ensure_ctx(s);
    pthread_mutex_unlock(&G.mu);
    if (erc) return g.err;
    if (G.dump_dir) return 0;
    int slot = -1;
    for (int i = 0; i < 32; i++) if (s->DPB[i].frame && s->DPB[i].frame->data[0] == frame->data[0]) slot = i;
    if (slot < 0) { fail(B200_EINVAL, "upload of a frame that is not in the DPB"); return g.err; }
    const void *planes[3] = { frame->data[0], frame->data[1], frame->data[2] };
    int64_t strides[3] = { frame->linesize[0], frame->linesize[1], frame->linesize[2] };
    int rc = b200_slot_upload(G.ctx, slot, planes, strides);
    if (!rc) rc = b200_sync(G.ctx);
    if (rc) fail(rc, b200_last_error(G.ctx));
    return rc;
}

void b200_shim_close(void)
{
    if (g.rec) b200_rec_destroy(g.rec);      /* recorders of other threads die with their threads' process */
    if (G.ctx) b200_ctx_destroy(G.ctx);
    G.ctx = NULL; G.next_ticket = G.turn = 0;
    ShimThread *t = &g;
    if (t != &g_oom) { const unsigned seq = t->frame_seq; memset(t, 0, sizeof(*t)); t->frame_seq = seq; }   /* workers compare (att, att_seq) */
}
