#!/usr/bin/env python3
"""The device source compiled for the host (tests/device_host) against the oracle on MANY random small worlds -- the
fuzz of tests/test_device_source_on_host.py::test_random_small_worlds beyond the 24 cases the suite runs (general, fast,
fused and single-step march; odd world sizes, out-of-world probes, magicSize dispatches, random fields and seeds).  No GPU.
usage: host_fuzz.py [first_case] [count]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
import test_device_source_on_host as T
import oracle as O

first = int(sys.argv[1]) if len(sys.argv) > 1 else 24
count = int(sys.argv[2]) if len(sys.argv) > 2 else 100
lib = T.build_host_lib(T.OUT)
t0 = time.time()
for case in range(first, first + count):
    T.test_random_small_worlds(lib, O, case)
print(f"cases {first}..{first + count - 1}: bit-exact in every march mode ({time.time() - t0:.0f} s)")
