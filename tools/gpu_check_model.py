"""Command-line front end of tests/gpu_cases_model.py (the engine-vs-oracle parity cases live under tests/ because they
import the oracle, which only tests/, smoke() and bench.py's CPU legs may do).
Run under gpurun:  python tools/gpu_check_model.py [case ...] > gpurun_out/model_check.log 2>&1"""
import os
import runpy
import sys

if __name__ == "__main__":
    here = os.path.dirname(os.path.abspath(__file__))
    target = os.path.join(os.path.dirname(here), "tests", "gpu_cases_model.py")
    sys.argv[0] = target
    runpy.run_path(target, run_name="__main__")
