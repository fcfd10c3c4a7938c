#!/bin/bash
# A/B of the side-stream schedule of the dense predict path (one box, back to back)
for v in "MURAL_TC_SIDE_STREAMS=0 MURAL_POOL_GRID=4" "MURAL_TC_SIDE_STREAMS=1" "MURAL_TC_SIDE_STREAMS=0 MURAL_POOL_GRID=4" "MURAL_TC_SIDE_STREAMS=1"; do
  echo "== $v"
  env $v bash scratch/bench_short.sh "$@" | head -1
done
