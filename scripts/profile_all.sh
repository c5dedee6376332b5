#!/bin/bash
# ncu evidence for profiles/: launch list of the bench + full captures of every kernel of the search path.
# usage (under gpurun): bash scripts/profile_all.sh <tag>
TAG=${1:-r2}
set -x
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-subrecords > gpurun_out/plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -s 12 -c 24 --csv \
    --log-file gpurun_out/${TAG}_launches_B.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-subrecords > gpurun_out/ncu_launch.log 2>&1
for WL in B 10M E C; do
  python scripts/quick_gpu.py --workloads $WL --iters 1 > gpurun_out/plain_$WL.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:k_score_topk -s 2 -c 1 -o gpurun_out/${TAG}_score_$WL \
      python scripts/quick_gpu.py --workloads $WL --iters 1 > gpurun_out/ncu_$WL.log 2>&1
done
python scripts/quick_gpu.py --workloads 10M --iters 1 --compress > gpurun_out/plain_10M_bf16.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_score_topk -s 2 -c 1 -o gpurun_out/${TAG}_score_10M_bf16 \
    python scripts/quick_gpu.py --workloads 10M --iters 1 --compress > gpurun_out/ncu_10M_bf16.log 2>&1
python scripts/quick_gpu.py --workloads 10M --iters 1 > gpurun_out/plain_aux.log 2>&1 &&
ncu --set full --clock-control none -k "regex:k_segments|k_merge" -s 4 -c 2 -o gpurun_out/${TAG}_aux_10M \
    python scripts/quick_gpu.py --workloads 10M --iters 1 > gpurun_out/ncu_aux.log 2>&1
ls -la gpurun_out/${TAG}_*
