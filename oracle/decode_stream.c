/*
 * decode_stream.c — TEST INFRASTRUCTURE: headless equivalent of the reference's CLI (main_hm/main.c:115-309
 * without SDL): decodes an Annex-B file through the public libOpenHevc* API (single thread) and prints one line
 * per output picture with the MD5 of each plane.  Linked once against the plain reference build
 * (decode_ref) and once against the build carrying the B200 hooks (decode_b200); identical output == the
 * drop-in is bit-exact on that stream.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include "openHevcWrapper.h"
#include "libavformat/avformat.h"
#include "libavutil/md5.h"

static void plane_md5(const uint8_t *p, int pitch, int w_bytes, int h, char *hex)
{
    struct AVMD5 *m = av_md5_alloc();
    uint8_t d[16];
    av_md5_init(m);
    for (int y = 0; y < h; y++) av_md5_update(m, p + (size_t)y * pitch, w_bytes);
    av_md5_final(m, d);
    for (int i = 0; i < 16; i++) sprintf(hex + 2 * i, "%02x", d[i]);
    av_free(m);
}

int main(int argc, char **argv)
{
    if (argc < 2) { fprintf(stderr, "usage: %s stream.hevc [threads[w] [quiet|time [passes]]]   (N: frame threads, hevc -p N -f 1; Nw: slice / WPP threads, -f 2; Nx: N slice threads inside frame threads, -f 4)\n", argv[0]); return 2; }
    const int threads = argc > 2 ? atoi(argv[2]) : 1;
    const int slice_threads = argc > 2 && strchr(argv[2], 'w') != NULL;
    const int frame_slice = argc > 2 && strchr(argv[2], 'x') != NULL;   /* pthread.c:57-71: frames = cpus / N + 1 */
    const int quiet = argc > 3 && strcmp(argv[3], "md5");      /* "md5": print the per-picture lines even with a pass count */
    const int loops = argc > 4 ? atoi(argv[4]) : 1;             /* decode the file this many times back to back (the stream
                                                                   starts with parameter sets + IDR, so the concatenation is a valid
                                                                   stream); the steady-state fps excludes the first pass (start-up) */
    const int timing = argc > 3 && !strcmp(argv[3], "time");    /* fps run: no MD5 work inside the timed loop */
    OpenHevc_Handle h = libOpenHevcInit(threads > 0 ? threads : 1, frame_slice ? 4 /* frame + slice */ : slice_threads ? 2 /* slice */ : 1 /* frame */);
    if (!h) return 3;
    libOpenHevcSetCheckMD5(h, 0);
    av_register_all();
    AVFormatContext *fmt = avformat_alloc_context();
    if (avformat_open_input(&fmt, argv[1], NULL, NULL) != 0) { fprintf(stderr, "cannot open %s\n", argv[1]); return 4; }
    int vs = av_find_best_stream(fmt, AVMEDIA_TYPE_VIDEO, -1, -1, NULL, 0);
    if (vs < 0) { fprintf(stderr, "no video stream\n"); return 5; }
    libOpenHevcSetDebugMode(h, 0);
    libOpenHevcStartDecoder(h);
    libOpenHevcSetTemporalLayer_id(h, 7);
    libOpenHevcSetActiveDecoders(h, 0);
    libOpenHevcSetViewLayers(h, 0);
    AVPacket pkt;
    int stop = 0, stop_dec = 0, nframes = 0, pass = 0;
    static struct timespec stamp[1 << 16];
    struct timespec t0, t1, tf;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    while (!stop) {
        if (!stop_dec && av_read_frame(fmt, &pkt) < 0) {
            if (++pass < loops) {                                 /* next pass: reopen the file, keep the decoder running */
                avformat_close_input(&fmt);
                fmt = avformat_alloc_context();
                if (avformat_open_input(&fmt, argv[1], NULL, NULL) != 0) return 4;
                continue;
            }
            stop_dec = 1;
        }
        if (stop_dec || pkt.stream_index == vs) {
            int got = libOpenHevcDecode(h, stop_dec ? NULL : pkt.data, stop_dec ? 0 : pkt.size, stop_dec ? 0 : pkt.pts);
            if (got > 0) {
                OpenHevc_Frame f;
                libOpenHevcGetOutput(h, 1, &f);
                const int B = f.frameInfo.nBitDepth > 8 ? 2 : 1;
                const int cw = f.frameInfo.chromat_format == YUV444 ? f.frameInfo.nWidth : f.frameInfo.nWidth / 2;
                const int ch = f.frameInfo.chromat_format == YUV420 ? f.frameInfo.nHeight / 2 : f.frameInfo.nHeight;
                char a[33], b[33], c[33];
                if (!timing) {
                plane_md5((const uint8_t *)f.pvY, f.frameInfo.nYPitch, f.frameInfo.nWidth * B, f.frameInfo.nHeight, a);
                plane_md5((const uint8_t *)f.pvU, f.frameInfo.nUPitch, cw * B, ch, b);
                plane_md5((const uint8_t *)f.pvV, f.frameInfo.nVPitch, cw * B, ch, c);
                if (!quiet) printf("frame %d %dx%d bd%d %s %s %s\n", nframes, f.frameInfo.nWidth, f.frameInfo.nHeight, f.frameInfo.nBitDepth, a, b, c);
                }
                if (nframes < (1 << 16)) clock_gettime(CLOCK_MONOTONIC, &stamp[nframes]);
                nframes++;
            } else if (stop_dec) stop = 1;
        }
        if (!stop_dec) av_free_packet(&pkt);
    }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    double sec = (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
    /* steady state: everything after the last picture of the first pass came out (one pass when there is only one) */
    const int n1 = loops > 1 ? nframes / loops : 1;
    tf = stamp[n1 > 0 && n1 <= (1 << 16) ? n1 - 1 : 0];
    double steady = (t1.tv_sec - tf.tv_sec) + 1e-9 * (t1.tv_nsec - tf.tv_nsec);
    printf("frames %d time %.3f fps %.2f first_frame_s %.3f steady_fps %.2f\n", nframes, sec, nframes / (sec > 0 ? sec : 1),
           nframes ? sec - steady : 0.0, nframes > n1 && steady > 0 ? (nframes - n1) / steady : 0.0);
    libOpenHevcClose(h);
    return 0;
}
