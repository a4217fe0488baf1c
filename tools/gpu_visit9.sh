#!/bin/bash
tag=${1:-b9}
mkdir -p gpurun_out
for lib in libb200hevc.so libb200hevc_sao_r2_b2.so libb200hevc_sao_r2_b3.so libb200hevc_sao_r2_b4.so libb200hevc_sao_r4_b3.so; do
  export B200_LIB_PATH=$PWD/openhevc_b200/$lib
  echo "== $lib" | tee -a gpurun_out/${tag}_sao.txt
  ( timeout 300 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "sao or sequence or c3_4k" 2>&1 | tail -1 ) | tee -a gpurun_out/${tag}_sao.txt
  timeout 300 python tools/dbk_sao_sweep.py --lanes 1,4 --pictures 64 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('c5 sweep', [(r['lanes'], round(r.get('pictures_per_s',0)), round(r.get('frac_of_hbm_peak',0),3)) for r in d['rows']])" | tee -a gpurun_out/${tag}_sao.txt
  timeout 300 python tools/dbk_sao_sweep.py --workload c3_4k_main10_ra --lanes 1,4 --pictures 128 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('c3 sweep', [(r['lanes'], round(r.get('pictures_per_s',0)), round(r.get('frac_of_hbm_peak',0),3)) for r in d['rows']])" | tee -a gpurun_out/${tag}_sao.txt
done
