#!/bin/bash
# 8-GPU visit: multi-GPU parity and the scaling bench line at N=8 (the driver runs N=1,2,4,8 itself at round end)
tag=${1:-b7}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $TR --nproc-per-node 8 --master-port 29521 tools/verify_multi_gpu.py --gops 8 2>&1 | tail -2 | tee gpurun_out/${tag}_verify8.txt
B200_TRACE=gpurun_out/${tag}_trace_gpu%d.csv timeout 900 $TR --nproc-per-node 8 --master-port 29522 bench.py --gpus 8 --no-cpu-baseline --no-stream > gpurun_out/${tag}_bench_n8.json 2> gpurun_out/${tag}_bench.err
cut -c1-400 gpurun_out/${tag}_bench_n8.json
for d in 0 7; do python tools/timeline.py gpurun_out/${tag}_trace_gpu$d.csv --from 64 --to 224 > gpurun_out/${tag}_timeline_gpu$d.txt 2>&1; tail -12 gpurun_out/${tag}_timeline_gpu$d.txt; done
timeout 900 $TR --nproc-per-node 8 --master-port 29523 bench.py --gpus 8 --steps 64 --no-cpu-baseline > gpurun_out/${tag}_bench_n8_stream.json 2>> gpurun_out/${tag}_bench.err
python -c "
import json
l=open('gpurun_out/${tag}_bench_n8_stream.json').read().strip().splitlines()
print('n8 e2e', json.loads(l[-1])['e2e'] if l else None)"
rm -f gpurun_out/${tag}_trace_gpu[1-6].csv
