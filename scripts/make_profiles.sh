#!/bin/bash
# Turn the ncu reports of scripts/profile_all.sh (gpurun_out/<tag>_*.ncu-rep) into the committed text summaries.
TAG=${1:-r2}
for WL in B 10M E C 10M_bf16; do
  R=gpurun_out/${TAG}_score_$WL.ncu-rep
  [ -f $R ] || continue
  python scripts/ncu_summary.py $R > profiles/${TAG}_score_${WL}_summary.txt
  python scripts/ncu_src.py $R 40 > profiles/${TAG}_score_${WL}_source_hotspots.txt
  python scripts/ncu_smem.py $R 24 > profiles/${TAG}_score_${WL}_smem_wavefronts.txt
done
[ -f gpurun_out/${TAG}_aux_10M.ncu-rep ] && python scripts/ncu_summary.py gpurun_out/${TAG}_aux_10M.ncu-rep > profiles/${TAG}_aux_kernels_10M_summary.txt
[ -f gpurun_out/${TAG}_launches_B.csv ] && cp gpurun_out/${TAG}_launches_B.csv profiles/${TAG}_launches_B.csv
python scripts/sass_opcodes.py > profiles/${TAG}_sass_opcodes.txt
python scripts/ncu_traffic.py $TAG > /dev/null
