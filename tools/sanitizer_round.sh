#!/bin/bash
# compute-sanitizer over the kernels that are new in round 2 (one gpurun call; output -> gpurun_out/$1/compute_sanitizer.txt)
OUT=gpurun_out/${1:-r2san}
mkdir -p $OUT
F=$OUT/compute_sanitizer.txt
echo "compute-sanitizer on B200, round 2 ($(git rev-parse --short HEAD 2>/dev/null))" > $F
echo "--- memcheck: tests/test_gpu_hector_gmapping.py (HectorSlamProcessor on the device: stream kernel exact / fast-cluster, batch kernels) + tests/test_deskew.py" >> $F
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_hector_gmapping.py tests/test_deskew.py -q -m gpu -k "slam or deskew" 2>&1 | tail -5 >> $F
echo "rc=$?" >> $F
echo "--- racecheck: the stream kernel (exact, fast with the cluster exchange) and the de-skew stage" >> $F
timeout 500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_hector_gmapping.py tests/test_deskew.py -q -m gpu -k "fast_mode or self_driven or deskew" 2>&1 | grep -v "^=========     \(Saved\|Host\)" | cut -c1-220 | sort | uniq -c | sort -rn | head -30 >> $F
echo "rc=$?" >> $F
echo "--- memcheck: tests/test_gpu_mapper.py (three robots into one mapper on the CUDA matcher)" >> $F
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_mapper.py -q -m gpu -k "multi_robot or golden" 2>&1 | tail -5 >> $F
echo "rc=$?" >> $F
cat $F
