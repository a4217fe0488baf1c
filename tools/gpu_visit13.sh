S=oracle/_ref/streams/c3_4k_ra8_calm_65.hevc
timeout 25 oracle/_ref/decode_b200 $S 32 md5 1 2>/dev/null | grep "^frame " | diff -q - oracle/_ref/streams/c3_4k_ra8_calm_65.md5 >/dev/null && echo "guard 32 threads: md5 equal" || echo "guard 32 threads: DIFFERENT"
echo "guard 32: $(timeout 25 oracle/_ref/decode_b200 $S 32 time 16 2>/dev/null | tail -1)"
echo "noguard 32: $(LD_LIBRARY_PATH=oracle/_ref/alt:openhevc_b200 timeout 25 oracle/_ref/decode_b200 $S 32 time 16 2>/dev/null | tail -1)"
echo "guard 16: $(timeout 25 oracle/_ref/decode_b200 $S 16 time 16 2>/dev/null | tail -1)"
echo "noguard 16: $(LD_LIBRARY_PATH=oracle/_ref/alt:openhevc_b200 timeout 25 oracle/_ref/decode_b200 $S 16 time 16 2>/dev/null | tail -1)"
