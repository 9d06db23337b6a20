echo "== base"; python scripts/bench_ln.py | grep "bwd rows=32768"
echo "== no atomics"; PLANK_B200_LN_DEBUG=1 python scripts/bench_ln.py | grep "bwd rows=32768"
echo "== no dropout"; PDROP=0 python scripts/bench_ln.py | grep "bwd rows=32768"
echo "== no dropout no atomics"; PDROP=0 PLANK_B200_LN_DEBUG=1 python scripts/bench_ln.py | grep "bwd rows=32768"
