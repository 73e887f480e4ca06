#!/bin/bash
# One gpurun call that produces the round's profiling evidence under gpurun_out/$1 (copied into profiles/ afterwards):
#   launches_bench_steps2.csv   every launch of `bench.py --steps 2 --warmup 1 --no-k2 --no-cpu-baseline` with its device time
#   k1_raw.csv / k1_source_*.csv ncu --set full of k_offsets_sorted / k_sweep_window / k_reduce at the bench batch (raw + source pages)
#   k2_raw.csv                  ncu --set full of the K2 / Hector kernels on bench-shaped inputs (tools/profile_k2.py)
# The .ncu-rep files are exported to CSV on the box and the large one is dropped (gpurun returns at most 64 MiB).
set -u
OUT=gpurun_out/${1:-r2z}
mkdir -p $OUT
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_bench_steps2.csv \
    python bench.py --steps 2 --warmup 1 --no-k2 --no-cpu-baseline > $OUT/launches_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_offsets_sorted|k_sweep_window|k_reduce" -s 3 -c 3 \
    -o $OUT/k1 python bench.py --steps 1 --warmup 1 --no-k2 --no-cpu-baseline > $OUT/ncu_k1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:"k_raytrace|k_gm_update|k_hs_mark|k_hs_apply|k_hs_stream|k_hs_match" -c 32 -o $OUT/k2 python tools/profile_k2.py 128 > $OUT/ncu_k2.log 2>&1
for n in k1 k2; do
  [ -f $OUT/$n.ncu-rep ] && ncu -i $OUT/$n.ncu-rep --page raw --csv > $OUT/${n}_raw.csv 2>/dev/null
done
[ -f $OUT/k1.ncu-rep ] && ncu -i $OUT/k1.ncu-rep --page source --print-source cuda,sass --csv --kernel-name regex:k_sweep_window > $OUT/k1_source_sweep.csv 2>/dev/null
[ -f $OUT/k2.ncu-rep ] && ncu -i $OUT/k2.ncu-rep --page source --print-source cuda,sass --csv --kernel-name regex:k_hs_apply --launch-skip 1 --launch-count 1 > $OUT/k2_source_apply.csv 2>/dev/null
[ -f $OUT/k2.ncu-rep ] && ncu -i $OUT/k2.ncu-rep --page source --print-source cuda,sass --csv --kernel-name regex:k_hs_mark --launch-skip 1 --launch-count 1 > $OUT/k2_source_mark.csv 2>/dev/null
rm -f $OUT/k2.ncu-rep
gzip -f $OUT/k1_source_sweep.csv $OUT/k2_source_apply.csv $OUT/k2_source_mark.csv 2>/dev/null
du -sh $OUT; ls -la $OUT
