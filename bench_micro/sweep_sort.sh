#!/bin/bash
run() { echo -n "$* : "; timeout 300 python bench.py --steps 64 --warmup 8 --no-e2e --no-cpu "$@" 2>&1 | python bench_micro/pick.py; }
run --sort-interval 8


run --sort-interval 4 --sort-miss 0.03 --sort-max 64
run --sort-interval 4 --sort-miss 0.10 --sort-max 64
run --sort-interval 4 --sort-miss 0.20 --sort-max 64
