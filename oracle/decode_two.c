/*
 * decode_two.c — TEST INFRASTRUCTURE: several decoders in ONE process through the public libOpenHevc* API (the
 * application side of main_hm/main.c:115-309, twice).  Two decoder instances are open at the same time and are fed
 * alternately, packet by packet, from the same thread; when both streams have ended the first decoder is closed and
 * a third one is opened on the first stream again (close -> open in one process).  One line per output picture,
 * prefixed with the decoder's letter:  "A frame 0 416x240 bd8 <md5 Y> <md5 Cb> <md5 Cr>".
 *   usage: decode_two a.hevc b.hevc [threads[w|x]]
 * Linked against the plain reference build (decode_two_ref) and against the build with the B200 hooks
 * (decode_two_b200): the shim keeps one device context, submission thread and ticket order per decoder instance.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "openHevcWrapper.h"
#include "libavformat/avformat.h"
#include "libavutil/md5.h"

typedef struct Dec { OpenHevc_Handle h; AVFormatContext *fmt; int vs, eof, done, nframes; char tag; } Dec;

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

static int dec_open(Dec *d, const char *path, int threads, int type, char tag)
{
    memset(d, 0, sizeof(*d));
    d->tag = tag;
    d->h = libOpenHevcInit(threads > 0 ? threads : 1, type);
    if (!d->h) return -1;
    libOpenHevcSetCheckMD5(d->h, 0);
    d->fmt = avformat_alloc_context();
    if (avformat_open_input(&d->fmt, path, NULL, NULL) != 0) { fprintf(stderr, "cannot open %s\n", path); return -1; }
    d->vs = av_find_best_stream(d->fmt, AVMEDIA_TYPE_VIDEO, -1, -1, NULL, 0);
    if (d->vs < 0) return -1;
    libOpenHevcSetDebugMode(d->h, 0);
    libOpenHevcStartDecoder(d->h);
    libOpenHevcSetTemporalLayer_id(d->h, 7);
    libOpenHevcSetActiveDecoders(d->h, 0);
    libOpenHevcSetViewLayers(d->h, 0);
    return 0;
}

static void dec_step(Dec *d)                       /* one packet in (or one flush call), at most one picture out */
{
    AVPacket pkt;
    if (d->done) return;
    if (!d->eof && av_read_frame(d->fmt, &pkt) < 0) d->eof = 1;
    if (!d->eof && pkt.stream_index != d->vs) { av_free_packet(&pkt); return; }
    int got = libOpenHevcDecode(d->h, d->eof ? NULL : pkt.data, d->eof ? 0 : pkt.size, d->eof ? 0 : pkt.pts);
    if (got > 0) {
        OpenHevc_Frame f;
        libOpenHevcGetOutput(d->h, 1, &f);
        const int B = f.frameInfo.nBitDepth > 8 ? 2 : 1;
        const int cw = f.frameInfo.chromat_format == YUV444 ? f.frameInfo.nWidth : f.frameInfo.nWidth / 2;
        const int ch = f.frameInfo.chromat_format == YUV420 ? f.frameInfo.nHeight / 2 : f.frameInfo.nHeight;
        char a[33], b[33], c[33];
        plane_md5((const uint8_t *)f.pvY, f.frameInfo.nYPitch, f.frameInfo.nWidth * B, f.frameInfo.nHeight, a);
        plane_md5((const uint8_t *)f.pvU, f.frameInfo.nUPitch, cw * B, ch, b);
        plane_md5((const uint8_t *)f.pvV, f.frameInfo.nVPitch, cw * B, ch, c);
        printf("%c frame %d %dx%d bd%d %s %s %s\n", d->tag, d->nframes++, f.frameInfo.nWidth, f.frameInfo.nHeight, f.frameInfo.nBitDepth, a, b, c);
    } else if (d->eof) d->done = 1;
    if (!d->eof) av_free_packet(&pkt);
}

static void dec_close(Dec *d)
{
    libOpenHevcClose(d->h);
    avformat_close_input(&d->fmt);
}

int main(int argc, char **argv)
{
    if (argc < 3) { fprintf(stderr, "usage: %s a.hevc b.hevc [threads[w|x]]\n", argv[0]); return 2; }
    const int threads = argc > 3 ? atoi(argv[3]) : 1;
    const int type = argc > 3 && strchr(argv[3], 'x') ? 4 : argc > 3 && strchr(argv[3], 'w') ? 2 : 1;
    av_register_all();
    Dec A, B, C;
    if (dec_open(&A, argv[1], threads, type, 'A') || dec_open(&B, argv[2], threads, type, 'B')) return 3;
    while (!A.done || !B.done) { dec_step(&A); dec_step(&B); }
    dec_close(&A);
    if (dec_open(&C, argv[1], threads, type, 'C')) return 3;
    while (!C.done) dec_step(&C);
    dec_close(&C);
    dec_close(&B);
    return 0;
}
