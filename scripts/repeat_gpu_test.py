"""Runs one GPU test function many times in ONE process with torch imported (the conditions under which the round-1
memset / memcpy ordering bug showed): python scripts/repeat_gpu_test.py tests.test_schemes_gpu:test_water_management_in_kwt 100"""
import importlib
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402,F401

mod, fn = sys.argv[1].split(":")
n = int(sys.argv[2]) if len(sys.argv) > 2 else 100
f = getattr(importlib.import_module(mod), fn)
t0 = time.time()
bad = 0
for i in range(n):
    try:
        f()
    except Exception as e:  # noqa: BLE001
        bad += 1
        print("run %d FAILED: %r" % (i, e), flush=True)
print("%s: %d runs, %d failures, %.1f s" % (sys.argv[1], n, bad, time.time() - t0), flush=True)
sys.exit(1 if bad else 0)
