#!/bin/bash
# GPU visit: the real decoder on the bench streams (thread sweep, both arms), then parity, then the bench line.
tag=${1:-s}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader > gpurun_out/${tag}_gpu.txt 2>&1
cat /sys/fs/cgroup/cpu.max >> gpurun_out/${tag}_gpu.txt 2>&1; nproc >> gpurun_out/${tag}_gpu.txt
S=oracle/_ref/streams
for st in ${STREAMS:-c3_4k_ra8_calm_65 c3_4k_ra8_mid_65 c3_4k_ra8_dense_33}; do
  [ -f $S/$st.hevc ] || continue
  for t in ${THREADS:-1 8 16 24}; do
    p=$(( t == 1 ? 2 : 6 ))
    echo "== $st threads $t" | tee -a gpurun_out/${tag}_sweep.txt
    ( timeout 300 oracle/_ref/decode_ref $S/$st.hevc $t time $p 2>/dev/null | tail -1 | sed 's/^/ref  /' ) | tee -a gpurun_out/${tag}_sweep.txt
    ( B200_SHIM_REPORT=1 timeout 300 oracle/_ref/decode_b200 $S/$st.hevc $t time $(( p * 2 )) 2>&1 | grep -E "^frames|b200 shim|rror" | sed 's/^/b200 /' ) | tee -a gpurun_out/${tag}_sweep.txt
  done
  ( timeout 300 oracle/_ref/decode_b200 $S/$st.hevc 16 2>/dev/null | grep '^frame ' | diff -q - $S/$st.md5 && echo "md5 ok $st (16 threads)" ) | tee -a gpurun_out/${tag}_sweep.txt
done
if [ -z "$SKIP_PYTEST" ]; then
  ( time timeout 900 python -m pytest tests -m gpu -x -q --durations=5 ${PYTEST_K:+-k "$PYTEST_K"} ) > gpurun_out/${tag}_pytest.log 2>&1
  tail -12 gpurun_out/${tag}_pytest.log
fi
if [ -z "$SKIP_BENCH" ]; then
  timeout 900 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
  cat gpurun_out/${tag}_bench.json
  timeout 600 python bench.py --impl reference > gpurun_out/${tag}_bench_ref.json 2>> gpurun_out/${tag}_bench.err
  cat gpurun_out/${tag}_bench_ref.json
fi
