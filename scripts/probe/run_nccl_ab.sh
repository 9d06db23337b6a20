run() { echo "$1: $(env $2 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $3 bench.py --gpus 2 --steps 20 --warmup 5 --no-decode --no-cpu-baseline --no-torch-cuda 2>/dev/null | python -c 'import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["ms_per_step"], d["value"])')"; }
run default "X=1" 29521
run minch32 "NCCL_MIN_NCHANNELS=32" 29522
run minch32_ctas "NCCL_MIN_NCHANNELS=32 NCCL_MIN_CTAS=32" 29523
run nvls_off "NCCL_NVLS_ENABLE=0 NCCL_MIN_NCHANNELS=32" 29524
