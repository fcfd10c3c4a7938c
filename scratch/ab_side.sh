#!/bin/bash
# A/B variants of the dense predict path via environment switches (one box, back to back)
for v in "X=1" "MURAL_TC_SIDE_POOLS=0" "X=1" "MURAL_TC_SIDE_POOLS=0"; do
  echo "== $v"
  env $v bash scratch/bench_short.sh "$@" | head -${HEADN:-1}
done
