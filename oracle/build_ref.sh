#!/bin/bash
# TEST INFRASTRUCTURE — builds the UNMODIFIED reference (pure-C path) into oracle/_ref/.
#
# Compiles the reference's own sources *where they lie* under /root/reference
# (nothing is copied into the repo; outputs are git-ignored) following the
# recipe of SURVEY.md Appendix C, minus CMake: config.h is generated from
# platform/x86/config.h.in with every SIMD / asm switch forced to 0, so
# ff_hevc_dsp_init / ff_hevc_pred_init install the C templates
# (hevcdsp_template.c / hevcpred_template.c) only.
#
# Output: oracle/_ref/libohevc_ref.so  (exports ff_hevc_dsp_init, ff_hevc_pred_init,
#         ff_videodsp_init, libOpenHevc*), oracle/_ref/gen/config.h
set -euo pipefail
REF=${REF:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
OUT=$HERE/_ref
if [ ! -d "$REF/libavcodec" ]; then
  echo "build_ref: $REF not present (GPU box?) - keeping prebuilt $OUT" >&2
  exit 0
fi
mkdir -p "$OUT/gen" "$OUT/obj"
CFG=$OUT/gen/config.h
sed -e 's/@USE_[A-Z0-9_]*@/0/g' \
    -e 's/@PTHREADS_FOUND@/1/' -e 's/@GMTIME_R_FOUND@/1/' -e 's/@FCNTL_H_FOUND@/1/' \
    -e 's/@GETPROCESSAFFINITYMASK_FOUND@/0/' -e 's/@GETTIMEOFDAY_FOUND@/1/' \
    -e 's/@LOCALTIME_R_FOUND@/1/' -e 's/@SCHED_GETAFFINITY_FOUND@/1/' \
    -e 's/@STRERROR_R_FOUND@/1/' -e 's/@SYSCONF_FOUND@/1/' -e 's/@UNISTD_H_FOUND@/1/' \
    -e 's/@USLEEP_FOUND@/1/' -e 's/@WINDOWS_H_FOUND@/0/' \
    -e 's/^#define ARCH_X86 .*/#define ARCH_X86 0/' \
    -e 's/^#define ARCH_X86_32 .*/#define ARCH_X86_32 0/' \
    -e 's/^#define ARCH_X86_64 .*/#define ARCH_X86_64 0/' \
    -e 's/^#define HAVE_INLINE_ASM .*/#define HAVE_INLINE_ASM 0/' \
    "$REF/platform/x86/config.h.in" | tr -d '\r' > "$CFG.tmp"
if grep -q '@' "$CFG.tmp"; then echo "unsubstituted tokens in config.h" >&2; grep -n '@' "$CFG.tmp" >&2; exit 1; fi
cmp -s "$CFG.tmp" "$CFG" 2>/dev/null || mv "$CFG.tmp" "$CFG"; rm -f "$CFG.tmp"

# the library file list is the CMake `libfilenames` variable (CMakeLists.txt:166-292)
awk '/^set\(libfilenames/{f=1;next} f&&/^\)/{f=0} f{print $1}' "$REF/CMakeLists.txt" | tr -d '\r' | grep '\.c$' > "$OUT/files.txt"
CFLAGS=${REF_CFLAGS:--O3 -DNDEBUG -fno-tree-vectorize}
echo "$CFLAGS" > "$OUT/cflags.txt.new"
if ! cmp -s "$OUT/cflags.txt.new" "$OUT/cflags.txt" 2>/dev/null; then rm -f "$OUT"/obj/*.o; mv "$OUT/cflags.txt.new" "$OUT/cflags.txt"; fi
rm -f "$OUT/cflags.txt.new"
compile_one() {
  f=$1; o="$OUT/obj/$(echo "$f" | tr / _ | sed 's/\.c$/.o/')"
  if [ ! -f "$o" ] || [ "$REF/$f" -nt "$o" ] || [ "$CFG" -nt "$o" ]; then
    gcc $CFLAGS -fPIC -std=gnu99 -w -DPIC -I"$OUT/gen" -I"$REF" -I"$REF/gpac/modules/openhevc_dec" \
        -c "$REF/$f" -o "$o" || exit 255
  fi
}
export -f compile_one; export OUT REF CFG CFLAGS
xargs -P "$(nproc)" -I{} bash -c 'compile_one {}' < "$OUT/files.txt"
gcc -shared -o "$OUT/libohevc_ref.so" "$OUT"/obj/*.o -lm -lpthread
echo "built $OUT/libohevc_ref.so"
