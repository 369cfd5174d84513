#!/bin/bash
# Round-end measurement on one GPU box: default bench (both arms), same-box A/B of this round's kernel switches,
# the ncu launch list of the bench command and full captures of the kernels named in DESIGN.md.
mkdir -p gpurun_out
python bench.py > gpurun_out/final_bench_n1.json 2> gpurun_out/final_bench_n1.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final_bench_ref.json 2> gpurun_out/final_bench_ref.err; echo "ref rc=$?"
sed -i "s/^timeout 900 python -m pytest.*$/true/" tools/gpu_ab.sh
tools/gpu_ab.sh finalab LMKD_LNG3=0 LMKD_TRX_DV_T=0 LMKD_LNG3=0,LMKD_TRX_DV_T=0
python tools/fusion_bench.py 1600 > gpurun_out/final_fusion.json 2> gpurun_out/final_fusion.err; cat gpurun_out/final_fusion.json
python tools/kernel_bench.py > gpurun_out/final_kernel_bench.json 2> gpurun_out/final_kernel_bench.err; echo "kernel_bench rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/final_launches.csv \
    python bench.py --steps 1 --warmup 3 --global-episodes 64 --no-cpu-baseline --no-extras --no-graph > gpurun_out/final_ncu_launches.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"ln_gather_bwd3|trx_attn_fwd|tuple_[kv]_fwd3" -c 8 -o gpurun_out/final_full_a -f \
    python tools/ncu_trx_step.py > gpurun_out/final_ncu_full_a.log 2>&1
echo "full a rc=$?"
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread \
    --clock-control none -c 200 --csv --log-file gpurun_out/final_kernel_metrics.csv python tools/ncu_trx_step.py > gpurun_out/final_ncu_metrics.log 2>&1
echo "metrics rc=$?"
