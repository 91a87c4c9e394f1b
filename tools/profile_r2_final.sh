set -u
NCU="ncu --set full --clock-control none --import-source on -f"
cap() {
  timeout 400 $NCU -k regex:$2 -s $3 -c 1 -o gpurun_out/r2_$1 python tools/profile_paths.py $4 > gpurun_out/r2_$1.log 2>&1
  echo "$1 rc=$?"
  python tools/ncu_summary.py full gpurun_out/r2_$1.ncu-rep gpurun_out/r2_$1_full.txt > /dev/null 2>&1
}
cap c2_scan_half scan_half_kernel 4 scan
cap c4_adc adc_fastscan_kernel 3 adc
rm -f gpurun_out/r2_c4_adc.ncu-rep
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r2_bench_launches.csv python bench.py --steps 20 --warmup 5 --no-cpu --no-check > gpurun_out/r2_bench_ncu.log 2>&1
echo "bench launches rc=$?"
ls -la gpurun_out/*.ncu-rep
