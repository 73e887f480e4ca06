#!/bin/bash
# the round's last verification on one GPU: GPU tests, smoke, default bench, reference arm, racecheck of the new kernels
OUT=gpurun_out/${1:-r2final2}
mkdir -p $OUT
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 > $OUT/pytest_gpu.txt; cat $OUT/pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.txt 2>&1; tail -2 $OUT/smoke.txt
timeout 900 python bench.py > $OUT/bench_n1.json 2> $OUT/bench_n1.err; tail -c 600 $OUT/bench_n1.json; tail -3 $OUT/bench_n1.err
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_reference_arm.json 2> $OUT/bench_ref.err; tail -c 300 $OUT/bench_reference_arm.json
