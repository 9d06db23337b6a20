set -x
python bench.py --steps 20 --warmup 5 > gpurun_out/r2b_bench_n1.json 2> gpurun_out/r2b_bench_n1.err
python bench.py --steps 10 --warmup 3 --workload config4 --no-decode --no-cpu-baseline --no-torch-cuda > gpurun_out/r2b_bench_config4_n1.json 2>/dev/null
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2b_step_metrics.csv python bench.py --profile-step --no-decode > /dev/null 2>&1
du -sh gpurun_out
