#!/bin/bash
# ncu captures for profiles/r2_*: run under gpurun on ONE GPU.  Every capture: --set full --clock-control none --import-source on.
set -u
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on -f"
cap() {  # name, kernel regex, skip, script arg
  timeout 400 $NCU -k regex:$2 -s $3 -c 1 -o gpurun_out/r2_$1 python tools/profile_paths.py $4 > gpurun_out/r2_$1.log 2>&1
  echo "$1 rc=$?"
  python tools/ncu_summary.py full gpurun_out/r2_$1.ncu-rep gpurun_out/r2_$1_full.txt > /dev/null 2>&1
  # the reports themselves only travel back when small (gpurun_out is capped at 64 MiB)
  if [ "$(stat -c %s gpurun_out/r2_$1.ncu-rep 2>/dev/null || echo 0)" -gt 9000000 ] && [ "$1" != "pq_tc" ]; then rm -f gpurun_out/r2_$1.ncu-rep; fi
}
cap c2_scan_half scan_half_kernel 4 scan
cap c2_scan scan_tma_kernel 4 scan32
cap c4_adc adc_fastscan_kernel 3 adc
cap c4_rank '^.*rank_kernel' 3 rerank
cap c2_b1024_gemm batch_gemm_kernel 2 batch
cap c2_b1024_select 'batch_select' 2 batch
cap c5_gemm_pair batch_gemm_pair_kernel 2 batch768
cap c5_select 'batch_select' 2 batch768
cap pq_tc pq_tc_assign_kernel 6 pq
cap knn knn_finalize_kernel 1 knn
# the exchange kernels (three ranks on this GPU): launch list only -- ncu serialises kernels, the host-wait shape tolerates that
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_exchange_launches.csv python tools/profile_paths.py exchange > gpurun_out/r2_exchange.log 2>&1
echo "exchange rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file gpurun_out/r2_c5_launches.csv python tools/profile_paths.py batch768 > gpurun_out/r2_c5.log 2>&1
echo "c5 launches rc=$?"
# launch list of the default bench
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r2_bench_launches.csv python bench.py --steps 20 --warmup 5 --no-cpu --no-check > gpurun_out/r2_bench_ncu.log 2>&1
echo "bench launches rc=$?"
ls -la gpurun_out/*.ncu-rep
