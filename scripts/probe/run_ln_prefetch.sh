for pf in 0 1 2 4; do for bl in 8 16; do
  echo "LN prefetch $pf blocks/SM $bl: $(PLANK_B200_LN_PREFETCH=$pf PLANK_B200_LN_BLOCKS_PER_SM=$bl python bench.py --steps 20 --warmup 5 --no-decode --no-cpu-baseline --no-torch-cuda 2>/dev/null | python -c 'import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["ms_per_step"])')"
done; done
