#!/usr/bin/env python3
"""Where an end-to-end frame of bench.py goes: each piece of Runner.step_e2e timed alone (CUDA events on the stage's
stream, L2 flushed before each), then the pipelined loop.  usage: e2e_breakdown.py [workload] [steps]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench

wl = sys.argv[1] if len(sys.argv) > 1 else "synthetic-256x256x64"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
torch.cuda.set_device(0)
r = bench.Runner(bench.make_scene(wl), 0, 0, 1, torch, None)
st = r.st


def timed(name, fn, sync_streams=True):
    ms = []
    for k in range(steps + 2):
        r.flush_l2(k)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(r.stream)
        t0 = time.perf_counter()
        fn(k)
        host = (time.perf_counter() - t0) * 1e3
        b.record(r.stream)
        st.wait_idle(); torch.cuda.synchronize()
        if k >= 2:
            ms.append((a.elapsed_time(b), host))
    g, h = np.mean([m[0] for m in ms]), np.mean([m[1] for m in ms])
    print(f"{name:48s} stream {g:8.3f} ms   host enqueue {h:7.3f} ms")


timed("upload_world_async (stream waits for the copy)", lambda k: st.upload_world_async(r.world_host))
timed("build_requests(sync=False)", lambda k: st.build_requests(sync=False))
timed("build_requests(sync=True)", lambda k: st.build_requests())
timed("update (validated world)", lambda k: st.update(k))
def up_update(k):
    st.upload_world_async(r.world_host); st.build_requests(sync=False); st.update(k)
timed("upload + build + update", up_update)
def seq(*parts):
    def f(k):
        for q in parts:
            q(k)
    return f
U = lambda k: st.upload_world_async(r.world_host)
B0 = lambda k: st.build_requests(sync=False)
B1 = lambda k: st.build_requests()
X = lambda k: st.update(k)
timed("build(async) + update", seq(B0, X))
timed("build(sync) + update", seq(B1, X))
timed("upload + update", seq(U, X))
timed("upload + build(sync) + update", seq(U, B1, X))
st.set_profiling(True)
for name, fn in (("update alone", X), ("upload + build(async) + update", seq(U, B0, X)), ("upload + update", seq(U, X)), ("build(async) + update", seq(B0, X))):
    ks = []
    for k in range(6):
        r.flush_l2(k); fn(k); st.wait_idle(); ks.append(st.last_kernel_ms())
    print(f"trace kernel alone inside '{name}': {np.mean(ks[2:]):.3f} ms")
for flips in range(3):
    ks = []
    for k in range(6):
        r.flush_l2(k); X(k); st.wait_idle(); ks.append(st.last_kernel_ms())
    print(f"trace kernel, update alone, after {flips} more world-buffer flips: {np.mean(ks[2:]):.3f} ms")
    U(0); st.wait_idle()
st.set_profiling(False)
timed("download_field_async + fence wait", lambda k: (st.download_field_async(r.field_host2[0]), st.record_fence(0), st.wait_fence(0)))
for span in (False, True):
    tot, per = r.timed(r.step_e2e, steps, 3, span=span)
    print(f"pipelined step_e2e, span={span}: {tot / steps:.3f} ms per step; per-step events: " + " ".join(f"{p:.2f}" for p in per))
tot, per = r.timed(r.step_resident, steps, 3)
print(f"step_resident: {tot / steps:.3f} ms per step")
r.close()
