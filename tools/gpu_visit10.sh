#!/bin/bash
tag=${1:-b10}
mkdir -p gpurun_out
for v in "B200_INTRA_CTAS=296" "B200_INTRA_CTAS=592" "B200_INTRA_CTAS=1184" "B200_INTRA_CTAS=592 B200_EDGES_SPARSE=1" "B200_INTRA_CTAS=1184 B200_EDGES_SPARSE=1"; do
  name=$(echo "$v" | tr ' =' '__')
  env $v timeout 400 python bench.py --steps 256 --no-cpu-baseline --no-stream > gpurun_out/${tag}_var_${name}.json 2>> gpurun_out/${tag}_bench.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/${tag}_var_${name}.json").read().strip().splitlines()[-1])
sp=d["roofline"]["stage_ms_by_picture"]
print("$v", "value %.0f" % d["value"], {k: round(x["ms"]*1000,1) for k,x in d["roofline"]["stages"].items()}, "I-picture intra us", round(sp["anchor_I#1"]["intra"]*1000), "b4 intra us", round(sp["b4#2"]["intra"]*1000,1))
PY
done
