#!/bin/bash
# Round-2 evidence on one B200: GPU tests, bench lines for every config, latency, probes, ncu launch list + full captures.
# Every ncu pass follows a plain run of the same command that exited 0; numbers printed under ncu are never bench values.
o=gpurun_out
mkdir -p $o
nvidia-smi --query-gpu=name,power.limit,clocks.max.sm --format=csv > $o/r02_smi.txt
timeout 400 python -m pytest tests -m gpu -q --timeout 240 2>&1 | tail -8 > $o/r02_tests_gpu.log; tail -2 $o/r02_tests_gpu.log
timeout 200 python bench.py --steps 20 --warmup 5 > $o/r02_bench_n1.json 2> $o/r02_bench_n1.err; echo bench20 rc=$?
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > $o/r02_bench_reference_arm.json 2>/dev/null; echo ref rc=$?
timeout 300 python bench.py > $o/r02_bench_n1_500steps.json 2> $o/r02_bench_n1_500steps.err; echo bench500 rc=$?
for c in cfg1 cfg2p cfg3 cfg3k4 cfg4; do timeout 240 python bench.py --config $c --steps 20 --warmup 5 > $o/r02_bench_$c.json 2> $o/r02_bench_$c.err; echo $c rc=$?; done
timeout 120 python tools/latency_bench.py > $o/r02_latency.jsonl 2>/dev/null; echo latency rc=$?
timeout 120 python tools/kernel_bench.py --iters 10 > $o/r02_kernel_bench.jsonl 2>/dev/null
timeout 120 python tools/kernel_bench.py --sustained --only proj_fwd,proj_fwd_2seg,proj_bwd_dw,proj_bwd_dw_db,torch_matmul_fwd,torch_matmul_dw > $o/r02_kernel_bench_loops.jsonl 2>/dev/null
PROBE_VARIANTS="AVC_GEMM_MT_TN=2" timeout 200 python tools/power_probe.py 1.5 2>/dev/null | grep kernel > $o/r02_power_probe.jsonl
timeout 100 python tools/gemm_profile.py 2>/dev/null | grep kernel > $o/r02_gemm_profile.jsonl
timeout 100 python tools/fused_ar_probe.py 2>&1 | tail -16 > $o/r02_fused_ar_probe.log
# ncu: launch list of a short bench run, then full captures of the step's kernels (fused and unfused step)
cmd="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-extras"
timeout 120 $cmd > /dev/null 2>&1 && timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $o/r02_launches_bench_steps3.csv $cmd > $o/r02_ncu_launches.log 2>&1; echo launches rc=$?
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'gemm_kernel|splice|pack_weight' --launch-skip 30 --launch-count 6 -o $o/r02_step_kernels -f $cmd > $o/r02_ncu_step.log 2>&1; echo ncu-step rc=$?
cmd4="python bench.py --config cfg4 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-extras"
timeout 120 $cmd4 > /dev/null 2>&1 && timeout 300 ncu --set full --clock-control none --import-source on -k regex:'gather_kernel|splice_kernel|gemm_kernel' --launch-skip 20 --launch-count 5 -o $o/r02_cfg4_kernels -f $cmd4 > $o/r02_ncu_cfg4.log 2>&1; echo ncu-cfg4 rc=$?
ls -la $o/r02_*.ncu-rep
