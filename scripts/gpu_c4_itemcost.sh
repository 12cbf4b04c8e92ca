#!/bin/bash
cd "$(dirname "$0")/.."
for cost in 16 64 128 256; do
  echo "== item cost $cost"; CLP_ITEM_COST=$cost timeout 600 python scripts/c4_once.py 2>&1 | tail -1
done
for m in 65536 50000; do echo "== m=$m cost 16 / 128"; CLP_ITEM_COST=16 timeout 600 python scripts/c4_once.py $m 2>&1 | tail -1; CLP_ITEM_COST=128 timeout 600 python scripts/c4_once.py $m 2>&1 | tail -1; done
