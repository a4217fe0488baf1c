#!/bin/bash
tag=${1:-b3}
mkdir -p gpurun_out
python tools/pin_probe.py 2>&1 | tee gpurun_out/${tag}_pin.txt
S=oracle/_ref/streams
for st in c3_4k_ra8_calm_65 c3_4k_ra8_mid_65 c3_4k_ra8_dense_33; do
  for t in 1 16; do
    p=$(( t == 1 ? 2 : 6 ))
    echo "== $st threads $t" | tee -a gpurun_out/${tag}_sweep.txt
    ( timeout 300 oracle/_ref/decode_ref $S/$st.hevc $t time $p 2>/dev/null | tail -1 | sed 's/^/ref        /' ) | tee -a gpurun_out/${tag}_sweep.txt
    for m in 0 1; do
      ( B200_DBD=$m B200_SHIM_REPORT=1 timeout 300 oracle/_ref/decode_b200 $S/$st.hevc $t time $(( p * 2 )) 2>&1 | grep -E "^frames|rror" | sed "s/^/b200 dbd=$m /" ) | tee -a gpurun_out/${tag}_sweep.txt
    done
  done
  ( timeout 300 oracle/_ref/decode_b200 $S/$st.hevc 16 2>/dev/null | grep '^frame ' | diff -q - $S/$st.md5 && echo "md5 ok $st (16 threads)" ) | tee -a gpurun_out/${tag}_sweep.txt
done
( time timeout 900 python -m pytest tests -m gpu -x -q --durations=5 ) > gpurun_out/${tag}_pytest.log 2>&1
tail -8 gpurun_out/${tag}_pytest.log
timeout 900 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
cat gpurun_out/${tag}_bench.json
