#!/bin/bash
# Developer A/B probe: time every library build under build/ab/*.so on the same box, interleaved.
# usage: scripts/ab_gpu.sh "<workloads>" [rounds] [extra quick_gpu args]
WL=${1:-B,10M}
ROUNDS=${2:-2}
shift; shift
for r in $(seq 1 $ROUNDS); do
  for so in build/ab/*.so; do
    echo "## $(basename $so) round $r"
    BM25_B200_LIB=$PWD/$so python scripts/quick_gpu.py --workloads $WL "$@" | grep -v '^#'
  done
done
