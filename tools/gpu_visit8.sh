#!/bin/bash
tag=${1:-b8}
mkdir -p gpurun_out
S=oracle/_ref/streams
st=c3_4k_ra8_calm_65
echo "== record-and-drop (no device), 16 and 32 threads" | tee -a gpurun_out/${tag}_sweep.txt
for t in 16 32; do ( { time B200_SHIM_DUMP=- timeout 300 oracle/_ref/decode_b200 $S/$st.hevc $t time 12 2>&1 | grep -E "^frames" ; } 2>&1 | tr '\n' ' ' ; echo ) | tee -a gpurun_out/${tag}_sweep.txt; done
echo "== with the device, 16 and 32 threads" | tee -a gpurun_out/${tag}_sweep.txt
for t in 16 32; do ( { time timeout 300 oracle/_ref/decode_b200 $S/$st.hevc $t time 12 2>&1 | grep -E "^frames" ; } 2>&1 | tr '\n' ' ' ; echo ) | tee -a gpurun_out/${tag}_sweep.txt; done
echo "== two decoder processes x 8 threads" | tee -a gpurun_out/${tag}_sweep.txt
( timeout 300 oracle/_ref/decode_b200 $S/$st.hevc 8 time 12 2>&1 | grep "^frames" & timeout 300 oracle/_ref/decode_b200 $S/$st.hevc 8 time 12 2>&1 | grep "^frames"; wait ) | tee -a gpurun_out/${tag}_sweep.txt
for v in "B200_MC_PAD=0" "B200_MC_PAD=1"; do
  name=$(echo "$v" | tr ' =' '__')
  ( env $v timeout 300 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "sequence or weighted or far_out" 2>&1 | tail -1 )
  env $v timeout 400 python bench.py --no-cpu-baseline --no-stream > gpurun_out/${tag}_var_${name}.json 2>> gpurun_out/${tag}_bench.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/${tag}_var_${name}.json").read().strip().splitlines()[-1])
print("$v", "value %.0f" % d["value"], {k: round(x["ms"]*1000,1) for k,x in d["roofline"]["stages"].items()})
PY
done
