#!/bin/bash
# ThreadSanitizer build of the table-level shim (submission thread, ticket queue, read-back table) under frame threads and WPP
# workers: record-only (B200_SHIM_DUMP=-) and with the emulated device behind it.  Needs /root/reference (shim headers) and the
# test binaries of oracle/_ref.  Round 2: silent on all runs below.
set -e
cd "$(dirname "$0")/.."
gcc -O1 -g -std=gnu99 -fPIC -w -shared -fsanitize=thread -Ioracle/_ref/gen -I/root/reference -Iinclude -o /tmp/libb200hevc_shim_tsan.so \
    openhevc_b200/csrc/shim/hevcdsp_init_b200.c -Lopenhevc_b200 -lb200hevc -Wl,-rpath,$PWD/openhevc_b200
rt=$(gcc -print-file-name=libtsan.so)
cd oracle/_ref
S=../../tests/golden/streams
for args in "$S/b_416x240_10b_weighted.hevc 4" "$S/ra_416x240_8b.hevc 4" "$S/wpp_416x240_8b_lowdelay.hevc 4w" "$S/wpp_832x480_10b_weighted.hevc 2x" "streams/c2_1080p_wpp_ra8_33.hevc 2x" "streams/c2_1080p_ra8_65.hevc 8"; do
  B200_SHIM_DUMP=- TSAN_OPTIONS="halt_on_error=0 report_signal_unsafe=0" LD_PRELOAD="$rt /tmp/libb200hevc_shim_tsan.so" timeout 900 ./decode_b200 $args quiet > /dev/null 2> /tmp/tsan_err.txt || true
  echo "record-only  $args: $(grep -c 'WARNING: ThreadSanitizer' /tmp/tsan_err.txt) reports"
done
for args in "$S/b_416x240_10b_weighted.hevc 4" "$S/ra_416x240_8b.hevc 4" "$S/wpp_416x240_8b_lowdelay.hevc 2x"; do
  TSAN_OPTIONS="halt_on_error=0 report_signal_unsafe=0" LD_PRELOAD="$rt /tmp/libb200hevc_shim_tsan.so $PWD/libb200hevc_emul.so" timeout 900 ./decode_b200 $args > /tmp/tsan_out.txt 2> /tmp/tsan_err.txt || true
  echo "emulated dev $args: $(grep -c 'WARNING: ThreadSanitizer' /tmp/tsan_err.txt) reports, pictures $(grep '^frame ' /tmp/tsan_out.txt | diff -q - ${args%% *} > /dev/null 2>&1; grep '^frame ' /tmp/tsan_out.txt | diff -q - $(echo ${args%% *} | sed 's/.hevc$/.md5/') > /dev/null && echo identical || echo DIFFERENT)"
done
# two decoder instances in one process, fed alternately, one closed and re-opened (oracle/decode_two.c)
for args in "$S/b_416x240_10b_weighted.hevc $S/ra_416x240_8b.hevc 4" "$S/wpp_416x240_8b_lowdelay.hevc $S/tiles_832x480_8b_lowdelay.hevc 2x"; do
  TSAN_OPTIONS="halt_on_error=0 report_signal_unsafe=0" LD_PRELOAD="$rt /tmp/libb200hevc_shim_tsan.so $PWD/libb200hevc_emul.so" timeout 900 ./decode_two_b200 $args > /tmp/tsan_out.txt 2> /tmp/tsan_err.txt || true
  echo "two decoders $args: $(grep -c 'WARNING: ThreadSanitizer' /tmp/tsan_err.txt) reports, $(grep -c ' frame ' /tmp/tsan_out.txt) pictures"
done
