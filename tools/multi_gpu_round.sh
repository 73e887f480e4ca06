#!/bin/bash
# usage: tools/multi_gpu_round.sh N tag   — the N-GPU measurements of one round under gpurun_out/<tag>/
set -u
N=$1; OUT=gpurun_out/${2:-r2m}; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port"
B2S_SPLIT_BIG=1 timeout 400 $TR 29511 tests/multi_gpu_checks.py > $OUT/multi_gpu_checks_n$N.json 2> $OUT/err_n$N.txt
[ "${4:-}" = "cfg2" ] && timeout 300 $TR 29512 bench.py --gpus $N --steps 20 --warmup 3 --min-seconds 2 --no-k2 --no-cpu-baseline > $OUT/bench_cfg2_n$N.json 2>> $OUT/err_n$N.txt
timeout 300 $TR 29513 bench.py --gpus $N --workload strong --steps 20 --min-seconds 2 > $OUT/bench_strong_n$N.json 2>> $OUT/err_n$N.txt
timeout 300 $TR 29514 bench.py --gpus $N --workload cfg4 --steps 10 --min-seconds 1 > $OUT/bench_cfg4_n$N.json 2>> $OUT/err_n$N.txt
timeout 240 $TR 29515 bench.py --gpus $N --workload cfg5 --steps 9 --nodes ${3:-600} > $OUT/bench_cfg5_n$N.json 2>> $OUT/err_n$N.txt
tail -c 600 $OUT/err_n$N.txt | grep -v OMP_NUM | grep -v "^\*" | tail -5
python - <<PY
import json
for w in ("cfg2","strong","cfg4","cfg5"):
    try:
        txt=open("$OUT/bench_%s_n$N.json" % w).read()
        d=json.loads([l for l in txt.splitlines() if l.startswith("{")][-1])
        print(w, "N=$N", round(d["value"]), "ms/step", round(d["ms_per_step"],3), "exact", (d.get("multi_gpu_exact") or {}).get("ok"), "timed_s", d.get("timed_seconds"))
    except Exception as e:
        print(w, "ERR", e)
PY
grep -h check $OUT/multi_gpu_checks_n$N.json | cut -c1-260
