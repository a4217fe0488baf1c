#!/bin/bash
# The Annex-B streams bench.py's `stream_e2e` block, tools/decode_bench.py and tests/test_stream_dropin.py (BASELINE-shape
# fixtures) read: oracle/_ref/streams/ (git-ignored, travels to the GPU box).  RECIPES, not bytes: every stream is a fixed
# seed + generator flags (tools/hevc_stream_gen.py, pure Python CABAC encoding: ~1 minute for a lightly coded 65-picture 4K
# stream, ~10 minutes for a dense one), regenerated here when missing; the per-picture MD5s of the UNMODIFIED reference
# decoder (oracle/_ref/decode_ref) are written next to each stream (.md5) and are what the hooked decoder must reproduce.
#   c3_4k_ra8_calm_65.hevc   3840x2160 Main10, hierarchical-B GOP 8, intra period 32, lightly coded (--calm 1: 75 % skipped CUs,
#                            large CUs, sparse residuals; ~45 KB per picture = ~11 Mbit/s at 30 Hz, the bit rate of real 4K streams)
#   c3_4k_ra8_mid_65.hevc    the same at --calm 0.5
#   c3_4k_ra8_dense_33.hevc  dense random content (~460 KB per picture, ~110 Mbit/s: the parse-bound worst case)
#   c2_1080p_ra8_65.hevc     1920x1080 8-bit, GOP 8 (BASELINE.json config 2)
#   c3_4k_wpp_ra8_calm_33.hevc  3840x2160 Main10, GOP 8, WPP, lightly coded: the stream bench.py times the three threading modes on (-f 1 / 2 / 4)
#   c2_1080p_wpp_ra8_33.hevc 1920x1080 8-bit, GOP 8, entropy_coding_sync (WPP): the stream of the frame + slice thread (-f 4) tests
#   c1_832x480_i_16.hevc     832x480 8-bit all-intra (config 1)
#   c5_8k_422_wpp_tiles_9.hevc  7680x4320 4:2:2 Main10 (RExt), entropy_coding_sync + 4x2 tiles (one substream per CTB row of every
#                            tile, hevc.c:2834), GOP 8, 9 pictures, lightly coded (config 5)
set -e
cd "$(dirname "$0")/.."
D=oracle/_ref/streams
mkdir -p $D
gen() {   # name, then generator arguments
  local name=$1; shift
  if [ ! -f $D/$name.hevc ]; then
    python tools/hevc_stream_gen.py $D/$name.hevc "$@" > $D/$name.gen.txt
  fi
  if [ ! -f $D/$name.md5 ] && [ -x oracle/_ref/decode_ref ]; then
    oracle/_ref/decode_ref $D/$name.hevc 1 2>/dev/null | grep '^frame ' > $D/$name.md5
  fi
}
gen c3_4k_ra8_calm_65 --width 3840 --height 2160 --bit-depth 10 --frames 65 --pattern RA8 --calm 1.0 --seed 9 &
gen c3_4k_ra8_mid_65 --width 3840 --height 2160 --bit-depth 10 --frames 65 --pattern RA8 --calm 0.5 --seed 10 &
gen c2_1080p_ra8_65 --width 1920 --height 1080 --bit-depth 8 --frames 65 --pattern RA8 --calm 0.7 --seed 11 &
gen c3_4k_wpp_ra8_calm_33 --width 3840 --height 2160 --bit-depth 10 --frames 33 --pattern RA8 --wpp --calm 1.0 --seed 19 &
gen c2_1080p_wpp_ra8_33 --width 1920 --height 1080 --bit-depth 8 --frames 33 --pattern RA8 --wpp --calm 0.7 --seed 21 &
gen c1_832x480_i_16 --width 832 --height 480 --bit-depth 8 --frames 16 --pattern I --seed 12 &
gen c5_8k_422_wpp_tiles_9 --width 7680 --height 4320 --bit-depth 10 --cfi 2 --frames 9 --pattern RA8 --wpp --tiles 4x2 --calm 1.0 --seed 55 &
if [ -z "$SKIP_DENSE" ]; then gen c3_4k_ra8_dense_33 --width 3840 --height 2160 --bit-depth 10 --frames 33 --pattern RA8 --seed 33 & fi
wait
ls -la $D
