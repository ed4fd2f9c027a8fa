import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def steps(default):
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=default)
    ap.add_argument("--every", type=int, default=0, help="print the status line every N iterations (0: 10 lines per run)")
    a = ap.parse_args()
    return a.steps, (a.every or max(1, a.steps // 10))
