echo "== min blocks 4"; python scripts/bench_ln.py | grep "bwd"
for bl in 12 16; do echo "== min blocks 4, grid $bl/SM"; PLANK_B200_LN_BLOCKS_PER_SM=$bl python scripts/bench_ln.py | grep "bwd rows=32768"; done
cp plankassembly_b200/csrc/libplank_b200.so /tmp/keep.so; cp scripts/probe/lib_mb5.so plankassembly_b200/csrc/libplank_b200.so
for bl in 8 10 16; do echo "== min blocks 5, grid $bl/SM"; PLANK_B200_LN_BLOCKS_PER_SM=$bl python scripts/bench_ln.py | grep "bwd rows=32768"; done
cp /tmp/keep.so plankassembly_b200/csrc/libplank_b200.so
