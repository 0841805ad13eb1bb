#!/bin/bash
# Round-2 evidence batch (one gpurun call): the whole GPU test suite, the bench lines of every
# workload, launch lists, and ncu --set full captures of one launch of every hot kernel.
# Everything lands in gpurun_out/; scripts/ncu_summary.py turns the .ncu-rep files into the
# tables committed under profiles/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T0=$(date +%s)
stamp() { echo "== $1 @ $(( $(date +%s) - T0 ))s"; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv,noheader
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 | tee gpurun_out/r2_pytest_gpu.txt
stamp pytest
timeout 300 python bench.py > gpurun_out/r2_bench_room.json 2> gpurun_out/bench_room.err
stamp bench_room
for wl in arena sortcheck gridworld room_render; do
  timeout 200 python bench.py --workload $wl --steps 100 --warmup 10 --no-cpu-baseline \
    > gpurun_out/r2_bench_$wl.json 2> gpurun_out/bench_$wl.err
done
stamp bench_others
for wl in room arena; do
  timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 700 -c 140 --csv \
    --log-file gpurun_out/r2_launches_$wl.csv python bench.py --workload $wl --steps 12 --warmup 4 --no-cpu-baseline \
    > gpurun_out/ncu_launch_$wl.log 2>&1
done
stamp launch_lists
timeout 300 ncu --set full --import-source on --clock-control none -k regex:"phys" --launch-skip 84 --launch-count 30 -f \
  -o gpurun_out/r2_final_phys python bench.py --workload room --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/ncu_phys.log 2>&1
stamp ncu_phys
timeout 200 ncu --set full --import-source on --clock-control none -k regex:"lidar" --launch-skip 4 --launch-count 1 -f \
  -o gpurun_out/r2_final_lidar python bench.py --workload room --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/ncu_lidar.log 2>&1
stamp ncu_lidar
timeout 200 ncu --set full --import-source on --clock-control none -k regex:"sort" --launch-skip 16 --launch-count 8 -f \
  -o gpurun_out/r2_final_sort python bench.py --workload sortcheck --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/ncu_sort.log 2>&1
stamp ncu_sort
timeout 300 ncu --set full --import-source on --clock-control none -k regex:"render" --launch-skip 8 --launch-count 6 -f \
  -o gpurun_out/r2_final_render python bench.py --workload room_render --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_render.log 2>&1
stamp ncu_render
for f in room arena sortcheck gridworld room_render; do
  python - "$f" <<'EOF'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/r2_bench_{sys.argv[1]}.json").read().strip().splitlines()[-1])
    r = d.get("roofline") or {}
    print(sys.argv[1], "ms/step %.4f" % d["ms_per_step"], "value %.3e" % d["value"], "e2e %.3e" % d["e2e"]["value"],
          "top", r.get("kernel"), "frac %.3f" % (r.get("frac") or 0), "cpu", (d.get("cpu_baseline") or {}).get("value"))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
EOF
done
ls -la gpurun_out/*.ncu-rep gpurun_out/r2_launches_*.csv
du -sh gpurun_out
