#!/bin/bash
tag=${1:-b4}
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q --durations=5 ) > gpurun_out/${tag}_pytest.log 2>&1
tail -8 gpurun_out/${tag}_pytest.log
S=oracle/_ref/streams
st=c3_4k_ra8_calm_65
for t in 12 14 15 16 20 24; do
  echo "== $st threads $t" | tee -a gpurun_out/${tag}_sweep.txt
  ( { time B200_SHIM_REPORT=1 timeout 300 oracle/_ref/decode_b200 $S/$st.hevc $t time 12 2>&1 | grep -E "^frames|rror" ; } 2>&1 | tr '\n' ' ' ; echo ) | tee -a gpurun_out/${tag}_sweep.txt
done
( { time timeout 300 oracle/_ref/decode_ref $S/$st.hevc 16 time 6 2>/dev/null | tail -1 ; } 2>&1 | tr '\n' ' '; echo ) | tee -a gpurun_out/${tag}_sweep.txt
for w in c3_4k_main10_ra c1_832x480_main c3_4k_main10_intra; do
  B200_INTRA=2 timeout 900 python bench.py --workload $w --no-cpu-baseline $( [ $w = c3_4k_main10_ra ] || echo --no-stream ) > gpurun_out/${tag}_bench_$w.json 2>> gpurun_out/${tag}_bench.err
  cat gpurun_out/${tag}_bench_$w.json | cut -c1-400
done
B200_INTRA=1 timeout 900 python bench.py --no-cpu-baseline --no-stream > gpurun_out/${tag}_bench_intra1.json 2>> gpurun_out/${tag}_bench.err
B200_INTRA=1 timeout 900 python bench.py --no-cpu-baseline --no-stream --workload c3_4k_main10_intra > gpurun_out/${tag}_bench_4ki_intra1.json 2>> gpurun_out/${tag}_bench.err
python tools/run_pictures.py --help > /dev/null 2>&1
