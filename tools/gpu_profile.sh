#!/bin/bash
# ncu evidence for profiles/: launch list of bench steps + full-set captures of the GEMM and attention kernels
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 138 -c 140 --csv --log-file gpurun_out/launches_bench_zk_cfg2.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_under_ncu.log 2>&1; echo "launch list rc=$?"
python tools/summarize_launches.py gpurun_out/launches_bench_zk_cfg2.csv > gpurun_out/launches_bench_zk_cfg2.txt; cat gpurun_out/launches_bench_zk_cfg2.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'attention_tc2_kernel' -c 6 -f -o gpurun_out/attn_tc2 python tools/ncu_attention_shapes.py 2 > gpurun_out/ncu_attn.log 2>&1; echo "ncu attention rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'gemm_pair16_kernel|gemm_ln_kernel' -c 8 -f -o gpurun_out/gemm python tools/ncu_gemm_shapes.py 2 > gpurun_out/ncu_gemm.log 2>&1; echo "ncu gemm rc=$?"
ls -la gpurun_out/*.ncu-rep
