#!/bin/bash
run() { echo -n "$* : "; timeout 300 python bench.py --steps 64 --warmup 8 --no-e2e --no-cpu "$@" 2>&1 | python bench_micro/pick.py; }
run --sort-full 64
run --sort-full 32
run --sort-full 16
run --sort-full 64 --sort-miss 0.02
run --sort-full 64 --sort-miss 0.05
run --sort-full 128
