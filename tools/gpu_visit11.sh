#!/bin/bash
# Round 2, visit 11: the new boundary features on the device (-f 4, two decoders in one process), then a frame-thread sweep of
# both arms on the headline stream (is 2 x cores really the best oversubscription?) and the -f 4 mode's fps.
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_stream_dropin.py -m gpu -x -q -k "slice_threads_inside or two_decoders or wpp_threads" ) > gpurun_out/b11_pytest.log 2>&1
tail -4 gpurun_out/b11_pytest.log
S=oracle/_ref/streams/c3_4k_ra8_calm_65.hevc
W=oracle/_ref/streams/c2_1080p_wpp_ra8_33.hevc
{
for t in 16 24 32 48 64; do echo "hooked $t: $(oracle/_ref/decode_b200 $S $t time 12 2>/dev/null | tail -1)"; done
for t in 16 32 48; do echo "reference $t: $(oracle/_ref/decode_ref $S $t time 3 2>/dev/null | tail -1)"; done
for t in 8 2x 4x 4w; do echo "1080p wpp hooked $t: $(oracle/_ref/decode_b200 $W $t time 12 2>/dev/null | tail -1)"; done
for t in 8 2x 4x 4w; do echo "1080p wpp reference $t: $(oracle/_ref/decode_ref $W $t time 3 2>/dev/null | tail -1)"; done
} 2>&1 | tee gpurun_out/b11_sweep.txt
