#!/bin/bash
# What the driver does at round end, in one visit: GPU suite, smoke, both bench arms.
tag=${1:-f}
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q --durations=6 ) > gpurun_out/${tag}_pytest.log 2>&1
tail -10 gpurun_out/${tag}_pytest.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2 | tee gpurun_out/${tag}_smoke.txt
( time timeout 900 python bench.py --impl reference ) > gpurun_out/${tag}_bench_ref.json 2> gpurun_out/${tag}_bench_ref.err
tail -3 gpurun_out/${tag}_bench_ref.err
( time timeout 900 python bench.py ) > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
tail -3 gpurun_out/${tag}_bench.err
python - <<PY
import json
for f in ("gpurun_out/${tag}_bench_ref.json","gpurun_out/${tag}_bench.json"):
    try:
        d=json.loads([l for l in open(f).read().splitlines() if l.startswith("{")][-1])
    except Exception as e:
        print(f,"ERR",e); continue
    print(f, "value %.1f"%d["value"], "e2e", d["e2e"]["value"], d["e2e"].get("md5_equal_reference_decoder"), d.get("clocks"))
    for k,v in (d.get("stream_e2e") or {}).items():
        if isinstance(v,dict): print("  ",k,{a:v[a].get("best") for a in ("reference","b200") if a in v}, {a:b for a,b in v.items() if a.startswith("speedup")})
    if "roofline" in d: print("  roofline", d["roofline"]["kernel"], round(d["roofline"]["frac"],4), {k:round(x["ms"]*1000,1) for k,x in d["roofline"]["stages"].items()}, "launches", d.get("gpu_launches"))
PY
