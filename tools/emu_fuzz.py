#!/usr/bin/env python3
"""The kernels on the host SIMT emulator (tests/device_host/simt_emu.hpp) against the oracle on MANY random small worlds --
tests/test_kernels_on_host.py::test_fuzz_worlds_kernels beyond the six cases the suite runs: pooled and per-workgroup
kernel, odd world sizes, out-of-world probes, magicSize dispatches, a loaded brick 0 and ids outside the palette (the
general march through the host flag and through validate_world_kernel's device flag).  No GPU; ~1 s per case.
usage: emu_fuzz.py [first_case] [count]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
import test_kernels_on_host as T
import oracle as O

first = int(sys.argv[1]) if len(sys.argv) > 1 else 16
count = int(sys.argv[2]) if len(sys.argv) > 2 else 50
lib = T.build(T.OUT)
t0 = time.time()
for case in range(first, first + count):
    T.test_fuzz_worlds_kernels(lib, O, case)
print(f"cases {first}..{first + count - 1}: both kernels bit-exact on the emulator ({time.time() - t0:.0f} s)")
