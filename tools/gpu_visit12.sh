#!/bin/bash
# Round 2, visit 12: fewer persistent blocks for the intra stage (visit 10 showed that MORE of them lower `value`: the polling warps
# take issue slots from the other pictures in flight).  Bench without the stream block, then the candidate against the stream MD5s.
tag=${1:-b12}
mkdir -p gpurun_out
for v in "B200_INTRA_CTAS=296" "B200_INTRA_CTAS=148" "B200_INTRA_CTAS=74"; do
  name=$(echo "$v" | tr ' =' '__')
  env $v timeout 300 python bench.py --steps 256 --no-cpu-baseline --no-stream > gpurun_out/${tag}_var_${name}.json 2>> gpurun_out/${tag}_bench.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/${tag}_var_${name}.json").read().strip().splitlines()[-1])
sp=d["roofline"]["stage_ms_by_picture"]
print("$v", "value %.0f" % d["value"], {k: round(x["ms"]*1000,1) for k,x in d["roofline"]["stages"].items()}, "I-picture intra us", round(sp["anchor_I#1"]["intra"]*1000), "b4 intra us", round(sp["b4#2"]["intra"]*1000,1))
PY
done
D=oracle/_ref/streams
for c in 148 74; do
  for a in "c1_832x480_i_16 1" "c3_4k_ra8_calm_65 16" "c2_1080p_ra8_65 8"; do
    set -- $a
    B200_INTRA_CTAS=$c oracle/_ref/decode_b200 $D/$1.hevc $2 md5 1 2>/dev/null | grep "^frame " | diff -q - $D/$1.md5 > /dev/null && echo "CTAS=$c $1 threads $2: md5 equal" || echo "CTAS=$c $1 threads $2: DIFFERENT"
  done
done
