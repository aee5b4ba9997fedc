#!/bin/bash
# Forward-GEMM variants at the cfg2 shape, isolated back-to-back loops (tools/kernel_bench.py --sustained).
out=${1:-gpurun_out/fwd_sweep.log}
shift
: > $out
run() { echo "== $*" >> $out; env "$@" python tools/kernel_bench.py --sustained --only ${ONLY:-proj_fwd_2seg,torch_matmul_fwd} 2>&1 | grep kernel >> $out; }
if [ $# -eq 0 ]; then set -- "AVC_X=0" "AVC_GEMM_MT_TN=2"; fi
for v in "$@"; do run $v; done
cat $out
