for b in 64 32 16; do for e in 0 2; do
  echo "batch $b EPI=$e: $(PLANK_B200_GEMM_EPI=$e python bench.py --steps 20 --warmup 5 --batch $b --no-decode --no-cpu-baseline --no-torch-cuda 2>/dev/null | python -c 'import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["ms_per_step"], d["value"], d["clocks"])')"
done; done
