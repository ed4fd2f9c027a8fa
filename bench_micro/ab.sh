#!/bin/bash
# A/B of advance-kernel variants: prints value, ms/step, achieved GB/s, frac, avg launch ms, window stats
for v in "$@"; do
  echo -n "variant $v: "
  ISKB_ADV_VARIANT=$v timeout 300 python bench.py --steps 24 --warmup 4 --no-e2e --no-cpu $ISKB_BENCH_ARGS 2>&1 | python bench_micro/pick.py
done
