#!/bin/bash
# final-state evidence: launch list of the bench command, full-set capture of the GEMM kernels, bench lines (zk, lds, lxmert)
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 134 -c 140 --csv --log-file gpurun_out/r01k_launches_bench_zk_cfg2.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_under_ncu.log 2>&1
echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'gemm_pair16_kernel|gemm_ln_kernel' -c 8 -f -o gpurun_out/r01k_gemm python tools/ncu_gemm_shapes.py 2 > gpurun_out/ncu_gemm.log 2>&1
echo "ncu full rc=$?"
timeout 600 python bench.py --steps 100 --warmup 5 > gpurun_out/bench_zk.log 2>&1; echo "bench zk rc=$?"; tail -1 gpurun_out/bench_zk.log | cut -c1-200
for m in imagebert_lds lxmert; do
timeout 600 python bench.py --steps 50 --warmup 5 --model $m --no-cpu-baseline > gpurun_out/bench_$m.log 2>&1; echo "bench $m rc=$?"; tail -1 gpurun_out/bench_$m.log | cut -c1-200
done
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.log 2>&1; echo "reference arm rc=$?"; tail -1 gpurun_out/bench_reference.log | cut -c1-400
