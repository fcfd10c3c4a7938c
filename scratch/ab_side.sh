#!/bin/bash
# A/B variants of the dense predict path via environment switches (one box, back to back)
for v in "X=1" "MURAL_AUTO_PRIO=1" "X=1" "MURAL_AUTO_PRIO=1"; do
  echo "== $v"
  env $v bash scratch/bench_short.sh "$@" | head -${HEADN:-1}
done
