"""The Python counterparts of the reference's problem scripts (problem/*.py) run end to end on the device:
BASELINE configs[0..2] (10_two_streams, 11_rf_discharge, 12_avalanche) and the N1 / N3 / N4 cases."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
@pytest.mark.parametrize("script,steps", [("p10_two_streams", 30), ("p11_rf_discharge", 20), ("p12_avalanche", 20),
                                          ("p07_boundaries", 20), ("p01_single_electron", 12), ("p04_mcc", 12), ("p06_circuit", 12), ("p13_seed", 12), ("p05_dsmc", 6), ("p_ts", 30)])
def test_problem_script_runs(script, steps):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "problem", script + ".py"), "--steps", str(steps)],
                       capture_output=True, text=True, timeout=120, cwd=os.path.join(ROOT, "problem"))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "Complete!" in r.stdout
    assert "('iteration', %d)" % steps in r.stdout
