#!/bin/bash
# Round-2 evidence batch (one gpurun call): new GPU tests, launch lists, ncu --set full captures
# of one launch of every hot kernel, sort tile-size A/B.  Everything lands in gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_physics_assets.py tests/test_tgs.py -m gpu -q 2>&1 | tail -3
for wl in room arena; do
  timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 700 -c 140 --csv \
    --log-file gpurun_out/r2_launches_$wl.csv python bench.py --workload $wl --steps 12 --warmup 4 --no-cpu-baseline \
    > gpurun_out/ncu_launch_$wl.log 2>&1
done
timeout 400 ncu --set full --import-source on --clock-control none -k regex:"phys" --launch-skip 84 --launch-count 12 -f \
  -o gpurun_out/r2_final_phys python bench.py --workload room --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/ncu_phys.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:"lidar" --launch-skip 4 --launch-count 1 -f \
  -o gpurun_out/r2_final_lidar python bench.py --workload room --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/ncu_lidar.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:"sort" --launch-skip 16 --launch-count 8 -f \
  -o gpurun_out/r2_final_sort python bench.py --workload sortcheck --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/ncu_sort.log 2>&1
timeout 400 ncu --set full --import-source on --clock-control none -k regex:"render" --launch-skip 8 --launch-count 4 -f \
  -o gpurun_out/r2_final_render python bench.py --workload room_render --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_render.log 2>&1
L=madrona_b200
scripts/ab_bench.sh "$L/libmadrona_b200.so $L/libmadrona_b200_sort12.so $L/libmadrona_b200_sort16.so" --workload sortcheck --steps 100 --warmup 10
MADRONA_B200_SWEEP_BLOCKS_PER_SM=2 scripts/ab_bench.sh "$L/libmadrona_b200_sort16.so" --workload sortcheck --steps 100 --warmup 10
MADRONA_B200_SWEEP_BLOCKS_PER_SM=3 scripts/ab_bench.sh "$L/libmadrona_b200_sort12.so" --workload sortcheck --steps 100 --warmup 10
ls -la gpurun_out/*.ncu-rep gpurun_out/r2_launches_*.csv
du -sh gpurun_out
