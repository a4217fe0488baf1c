/*
 * b200hevc_tables.h — the reference-facing entry points of libb200hevc_shim.so: what a maintainer of
 * openHEVC links against to route the HEVCDSPContext / HEVCPredContext / VideoDSPContext function
 * tables to the B200.  Each one replaces / complements a hook of the reference:
 *
 *   ff_hevcdsp_init_b200   like ff_hevcdsp_init_x86 / _arm, called at the end of ff_hevc_dsp_init
 *                          (reference libavcodec/hevcdsp.c:1326-1327, prototypes hevcdsp.h:173-174)
 *   ff_hevcpred_init_b200  like ff_hevcpred_init_x86 (libavcodec/hevcpred.c:84, hevcpred.h:44)
 *   ff_videodsp_init_b200  like ff_videodsp_init_x86 (libavcodec/videodsp.c:51-58, videodsp.h:94-97)
 *   b200_frame_begin       after hevc_frame_start() has chosen the DPB slot      (libavcodec/hevc.c:3245)
 *   b200_frame_end         when every CTB of the picture has been parsed and filtered (libavcodec/hevc.c:3447-3449, after tiles_filters)
 *   b200_frame_readback    when the packet has been decoded, before the picture is hashed or output (libavcodec/hevc.c:4141);
 *                          frame == NULL: no complete picture came out of the packet -- an abandoned one is closed
 *   b200_output_wait       where hevc_decode_frame hands a picture to its caller (libavcodec/hevc.c:4177-4180, and the flush path
 *                          :4116): the read-back of a picture is issued with the picture and only waited for here -- a picture
 *                          leaves the decoder long after it was decoded (bumping, hevc_refs.c:182-307)
 *   b200_frame_buffer_alloc  allocator for the decoder's frame pool in place of av_buffer_allocz (libavcodec/utils.c:558-561):
 *                          picture planes in page-locked memory, so that read-backs are asynchronous DMAs at full PCIe rate
 *   b200_worker_begin      first statement of hls_decode_entry_wpp / _tiles / _wpp_in_tiles (libavcodec/hevc.c:2764, 2847, 2931):
 *                          the slice / WPP / tile worker thread records for the picture of THIS context (frame + slice threads)
 *   b200_decoder_close     first statement of hevc_decode_free() (libavcodec/hevc.c:4193): the decoder's device context, its
 *                          submission thread and its entry in the shim's instance table are given back
 *   b200_frame_fill        when generate_missing_ref() has filled a grey reference  (libavcodec/hevc_refs.c:538-606)
 *   b200_bs_on_device      first statement of ff_hevc_deblocking_boundary_strengths() (libavcodec/hevc_filter.c:808): the call is
 *                          recorded as one word and the function returns (non-zero result)
 *   b200_deblock_on_device first statement of deblocking_filter_CTB() (libavcodec/hevc_filter.c:347): non-zero = the device
 *                          derives boundary strengths, tc and beta itself, nothing to do on the host
 *   b200_host_pixels_unused  OPTIONAL, performance only: guard at the top of copy_CTB()  (libavcodec/hevc_filter.c:151-161) --
 *                          sao_filter_CTB copies every CTB between the host frame and sao_frame before it calls the SAO
 *                          tables; with the tables on the device nobody reads those host pixels
 *
 * The structs are the reference's own (opaque here); the implementation
 * (openhevc_b200/csrc/shim/hevcdsp_init_b200.c) is compiled against the reference headers.
 * All int functions return 0 or a negative B200_E* code (include/b200hevc.h); the table functions
 * themselves return void like the slots they replace, errors are latched and reported by
 * b200_frame_end / b200_frame_readback, text from b200_shim_error().
 */
#ifndef B200HEVC_TABLES_H
#define B200HEVC_TABLES_H

#ifdef __cplusplus
extern "C" {
#endif

struct HEVCDSPContext;
struct HEVCPredContext;
struct VideoDSPContext;
struct HEVCContext;
struct AVFrame;
struct HEVCFrame;
struct AVBufferRef;

void ff_hevcdsp_init_b200(struct HEVCDSPContext *c, const int bit_depth);
void ff_hevcpred_init_b200(struct HEVCPredContext *c, const int bit_depth);
void ff_videodsp_init_b200(struct VideoDSPContext *c, int bpc);

int  b200_frame_begin(struct HEVCContext *s);
int  b200_frame_end(struct HEVCContext *s);
int  b200_frame_readback(struct HEVCContext *s, struct AVFrame *frame);
int  b200_output_wait(struct HEVCContext *s, struct AVFrame *frame);        /* the frame leaves the decoder: its pixels have landed when this returns */
struct AVBufferRef *b200_frame_buffer_alloc(int size);                      /* frame-pool allocator: pinned host memory */
int  b200_frame_fill(struct HEVCContext *s, struct HEVCFrame *frame);       /* grey reference picture (generate_missing_ref) */
int  b200_frame_upload_ref(struct HEVCContext *s, struct AVFrame *frame);   /* host-only reference picture -> device slot */
int  b200_bs_on_device(struct HEVCContext *s, int x0, int y0, int log2_size);
int  b200_deblock_on_device(void);
int  b200_worker_begin(struct HEVCContext *owner);                          /* first statement of an execute2 job: owner = avctx->priv_data */
int  b200_host_pixels_unused(void);                                         /* 1 once the B200 tables are installed */
void b200_decoder_close(struct HEVCContext *s);                             /* hevc_decode_free: this decoder's device context is released */
void b200_shim_close(void);
const char *b200_shim_error(void);

#ifdef __cplusplus
}
#endif
#endif /* B200HEVC_TABLES_H */
