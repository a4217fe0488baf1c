#!/bin/bash
# The two 4K Main10 streams tools/decode_bench.py and tools/hostprof/hosttime.py read (oracle/_ref/streams/, git-ignored,
# travels to the GPU box).  Pure Python CABAC encoding: several minutes for the dense one.
#   c3_4k_33.hevc       dense random content, ~460 KB per picture (~110 Mbit/s at 30 Hz): the parse-bound worst case
#   c3_4k_calm_17.hevc  lightly coded (--calm 1: 75 % skipped CUs, large CUs, sparse residuals), ~46 KB per picture
#                       (~11 Mbit/s at 30 Hz): the bit rate of real 4K streams, where the pixel path dominates the reference
set -e
cd "$(dirname "$0")/.."
mkdir -p oracle/_ref/streams
[ -f oracle/_ref/streams/c3_4k_calm_17.hevc ] || python tools/hevc_stream_gen.py oracle/_ref/streams/c3_4k_calm_17.hevc \
    --width 3840 --height 2160 --bit-depth 10 --frames 17 --pattern RA --calm 1.0 --seed 9 > oracle/_ref/streams/c3_4k_calm_17.gen.txt
[ -f oracle/_ref/streams/c3_4k_33.hevc ] || python tools/hevc_stream_gen.py oracle/_ref/streams/c3_4k_33.hevc \
    --width 3840 --height 2160 --bit-depth 10 --frames 33 --pattern PBBB --seed 33 > oracle/_ref/streams/c3_4k_33.gen.txt
ls -la oracle/_ref/streams/
