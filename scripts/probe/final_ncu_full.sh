set -x
mkdir -p /tmp/ncu
ncu --set full --clock-control none --profile-from-start off -k regex:"gemm_tf32_kernel|attn_fwd_pp_kernel|attn_bwd_dkdv|attn_bwd_dq|add_ln_bwd" -c 24 -o /tmp/ncu/r2b_full python bench.py --profile-step --no-decode > /dev/null 2>&1
python scripts/ncu_metrics.py /tmp/ncu/r2b_full.ncu-rep > gpurun_out/r2b_ncu_full_first24.csv
ls -la /tmp/ncu; du -sh gpurun_out
