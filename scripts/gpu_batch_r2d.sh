#!/bin/bash
# Round-2 batch D (last gpurun call): the committed defaults -- full GPU parity suite, the bench
# line of every workload, the reference arm, launch list + ncu capture of the final sort kernels.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T0=$(date +%s)
stamp() { echo "== $1 @ $(( $(date +%s) - T0 ))s"; }
timeout 600 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider 2>&1 | tail -40 > gpurun_out/r2d_pytest_gpu.txt
tail -6 gpurun_out/r2d_pytest_gpu.txt
stamp pytest
timeout 200 python bench.py 2>gpurun_out/err_d_room.txt | tail -1 > gpurun_out/r2d_bench_room.json
for wl in sortcheck gridworld arena room_render; do
  timeout 150 python bench.py --workload $wl --steps 100 --warmup 10 --no-cpu-baseline 2>gpurun_out/err_d_$wl.txt \
    | tail -1 > gpurun_out/r2d_bench_$wl.json
done
timeout 150 python bench.py --impl reference --steps 20 --warmup 5 2>gpurun_out/err_d_ref.txt | tail -1 > gpurun_out/r2d_reference_room.json
python - <<'PY'
import json
for wl in ("room", "arena", "sortcheck", "gridworld", "room_render"):
    try:
        d = json.loads(open(f"gpurun_out/r2d_bench_{wl}.json").read())
        r = d.get("roofline") or {}
        print("FINAL %-12s ms/step %.4f value %.4g e2e %.4g top %s frac %.3f cpu %s" % (wl, d["ms_per_step"], d["value"],
              d["e2e"]["value"], r.get("kernel"), r.get("frac") or 0, (d.get("cpu_baseline") or {}).get("value")))
    except Exception as e:
        print("FINAL", wl, "FAILED", e)
try:
    print("REFERENCE ARM", json.loads(open("gpurun_out/r2d_reference_room.json").read())["value"])
except Exception as e:
    print("REFERENCE ARM FAILED", e)
PY
stamp finals
timeout 150 ncu --set full --clock-control none -k regex:"sort" --launch-skip 14 --launch-count 7 -f \
  -o gpurun_out/r2d_sort python bench.py --workload sortcheck --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/ncu_sort.log 2>&1
stamp ncu_sort
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 300 -c 60 --csv \
  --log-file gpurun_out/r2d_launches_gridworld.csv python bench.py --workload gridworld --steps 12 --warmup 4 --no-cpu-baseline \
  > gpurun_out/ncu_launch_gridworld.log 2>&1
stamp launch_list
ls -la gpurun_out/r2d_*
