#!/usr/bin/env python3
"""bench.py -- radiance-probe rays/s and ms per full radiance-field update (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload ...]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...      (N > 1, one rank per GPU)

A step = one full radiance-field update: every probe selected by Lum's request rule x 64 rays.

Workloads (config.workload):
  synthetic-256x256x64  BASELINE.json configs[3] (SURVEY.md 8d "C4"): the configuration the metric
                        is quoted on at 1/2/4/8 GPUs; fits one GPU, so it is the workload at every N
                        (strong scaling: total work fixed, the dispatch's probes dealt to the ranks in
                        interleaved chunks, the new texels gathered by every rank each step).  DEFAULT.
  demo                  BASELINE.json configs[1]: Lum's demo world, full recompute (2,689 probes).
                        Also measured at N=1 in the default run and reported under "demo".

Timed regions (CUDA events on the stage's stream, L2 flushed between timed steps, max over ranks):
  value  inputs resident in HBM: update kernel(s) + commit; when N > 1 the commit IS the exchange (one flag
         barrier, then one kernel reads every texel from its owner's staging array over NVLink and stores
         it locally; --exchange nccl = z-slab update + NCCL all-gather of the field)
  e2e    through the C-ABI with HOST buffers: world upload (pinned staging) + device request-list
         build + update (+ all-gather) + field download
The oracle (oracle/) is used here only for the cpu_baseline leg and --impl reference.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "radiance_probe_rays_per_sec"
UNIT = "rays/s"
RAYS = 64


# ----------------------------------------------------------------------------- helpers ----
def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def load_profile_evidence(workload, n_gpus, rays):
    """What one `ncu --set full` capture of the trace kernel on this workload says (profiles/roofline_evidence.json,
    written by tools/ncu_evidence.py from the .ncu-rep; the file names the capture it came from): DRAM bytes per
    launch (roofline.traffic), issue-slot and lane utilisation, achieved L1 / L2 sector rates -- next to the
    ceilings of this access pattern measured by tools/gather_ceiling.cu (profiles/r02_gather_ceiling.json).
    Only single-GPU launches at 64 rays per probe were captured; {} otherwise."""
    if n_gpus != 1 or rays != 64:
        return {}
    try:
        ev = json.load(open(os.path.join(ROOT, "profiles", "roofline_evidence.json"))).get(workload)
        return ev or {}
    except Exception:
        return {}


def make_scene(workload):
    from lum_b200 import scenes
    if workload == "demo":
        d = scenes.demo_scene()
        return dict(world=d["world"], bricks=d["bricks"], materials=d["materials"])
    if workload.startswith("synthetic-"):
        wx, wy, wz = (int(v) for v in workload.split("-")[1].split("x"))
        return scenes.synthetic_scene(wx, wy, wz)
    raise SystemExit(f"unknown workload {workload}")


class ClockSampler:
    """nvidia-smi clocks line of B200_PROFILING.md, sampled during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.proc = None
        self.path = None

    def start(self):
        try:
            f = tempfile.NamedTemporaryFile(prefix="clocks_", suffix=".csv", delete=False)
            self.path = f.name
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if not self.proc:
            return out
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            for line in open(self.path):
                c = [x.strip() for x in line.split(",")]
                if len(c) < 9:
                    continue
                try:
                    sm.append(float(c[1])); mx.append(float(c[2]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), samples=len(sm))
        out["reasons"] = sorted(reasons)
        return out


# ----------------------------------------------------------------------------- ours -------
class Runner:
    """One rank's stage + the per-step sequences (value region and e2e region)."""

    def __init__(self, scene, device, rank, world_size, torch, dist, exchange="fused", rays=RAYS):
        import lum_b200
        from lum_b200 import distributed as D
        self.D = D
        self.torch, self.dist = torch, dist
        self.rank, self.n = rank, world_size
        self.scene = scene
        wz, wy, wx = scene["world"].shape
        if wz % world_size:
            raise SystemExit(f"world z={wz} not divisible by {world_size} ranks")
        self.z0, self.z1 = D.slab_bounds(wz, rank, world_size)
        self.rays = rays
        self.device = device
        self.st = lum_b200.RadianceStage((wx, wy, wz), scene["bricks"].shape[0], rays_per_probe=rays, device=device)
        self.st.upload_world(scene["world"])
        self.st.upload_bricks(0, scene["bricks"])
        self.st.upload_materials(scene["materials"])
        self.nreq = self.st.build_requests()
        self.cells = wx * wy * wz
        self.slab_cells = self.cells // world_size
        self.stream = torch.cuda.ExternalStream(self.st.stream(), device=device)
        self.dims = (wx, wy, wz)
        self.field_t = D.device_tensor_view(self.st.field_device_ptr(), self.cells, f"cuda:{device}")
        # e2e host buffers live in pinned memory (the contract's "from pinned host memory"): the world
        # goes up without a staging copy and the field comes down at full PCIe rate
        self._pin_w = torch.empty(self.cells, dtype=torch.int16, pin_memory=True)
        self._pin_f = torch.empty(2 * self.cells, dtype=torch.int32, pin_memory=True)     # one result buffer per frame in flight
        self.world_host = self._pin_w.numpy()
        self.world_host[:] = np.ascontiguousarray(scene["world"], np.int16).reshape(-1)
        self.field_host = self._pin_f.numpy().view(np.uint32)[:self.cells]
        self.field_host2 = [self._pin_f.numpy().view(np.uint32)[:self.cells], self._pin_f.numpy().view(np.uint32)[self.cells:]]
        self.zero_field = np.zeros(self.cells, np.uint32)
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device=f"cuda:{device}")   # > 126 MB L2
        self.fused = False
        if world_size > 1 and exchange == "fused":
            D.open_peer_exchange(self.st, rank, world_size, device=f"cuda:{device}")
            self.fused = True

    def close(self):
        """Tear down in an order torch can live with: the pinned buffers were used on the stage's stream (an
        ExternalStream), and torch's host allocator records an event on that stream when they are freed -- so they
        must go while the stream still exists, i.e. before lumcu_destroy."""
        self.torch.cuda.synchronize()
        if self.fused:
            self.st.peers_close()
            self.fused = False
        self.world_host = self.field_host = self.field_host2 = None
        self._pin_w = self._pin_f = None
        self.field_t = self.flush = None
        self.torch.cuda.synchronize()
        self.st.close()

    def flush_l2(self, k=0):
        with self.torch.cuda.stream(self.stream):
            self.flush.fill_(k & 0xFF)

    def reset(self):
        if self.n > 1:                      # nobody may still be storing into a peer's replica
            self.st.wait_idle()
            self.dist.barrier()
        self.st.upload_field(self.zero_field)
        self.st.wait_idle()
        if self.n > 1:
            self.dist.barrier()

    def exchange(self):
        if self.n > 1:
            with self.torch.cuda.stream(self.stream):
                self.D.all_gather_field(self.field_t, self.dims, self.rank, self.n)

    def step_resident(self, seed):
        if self.fused:
            self.st.update_shard_exchange(seed, 1, 0)          # interleaved shards; commit = exchange (peer stores)
        elif self.n > 1:
            self.st.update_slab(seed, 1, 0, self.z0, self.z1)
            self.exchange()                                                # NCCL all-gather
        else:
            self.st.update(seed, 1, 0)

    def step_e2e(self, seed):
        """One frame through the C-ABI with HOST buffers, queued without waiting for the GPU: world H2D from pinned
        memory (end_blockify/exec_copies), device-side request build (the count stays on the device), update
        (+ exchange), D2H of the result into pinned memory (each rank reads its slab at N > 1), a fence.  Like the
        reference (fif = 2, al.cpp:315-317) two frames are in flight: the host waits for the PREVIOUS frame's fence,
        so the copies of one frame overlap the tracing of its neighbours.  Everything is inside the timed region."""
        slot = seed & 1
        self.st.wait_fence(slot)                              # the frame that last used this result buffer is done
        self._mark()
        self.st.upload_world_async(self.world_host)           # H2D
        self._mark()
        self.st.build_requests(sync=False)
        self._mark()
        self.step_resident(seed)
        self._mark()
        lo, hi = (self.rank * self.slab_cells, (self.rank + 1) * self.slab_cells) if self.n > 1 else (0, self.cells)
        self.st.download_field_async(self.field_host2[slot][lo:hi], lo, hi - lo)      # D2H
        self.st.record_fence(slot)
        self._mark()
        return self.field_host2[slot][lo:hi]

    phase_events = None

    def _mark(self):
        if self.phase_events is not None:
            e = self.torch.cuda.Event(enable_timing=True)
            e.record(self.stream)
            self.phase_events.append(e)

    def phase_breakdown(self):
        """ms on the stage's stream between the marks of step_e2e (mean over the steps of the last diagnostic loop):
        [world upload seen by the stream, request build, update (+ exchange), after the download was queued]."""
        ev = self.phase_events or []
        n = len(ev) // 5
        if n < 2:
            return None
        out = np.zeros(4)
        for k in range(1, n):                      # skip the first step (nothing to overlap with)
            for j in range(4):
                out[j] += ev[5 * k + j].elapsed_time(ev[5 * k + j + 1])
        return [float(v / (n - 1)) for v in out]

    def drain(self):
        self.st.wait_fence(0); self.st.wait_fence(1)

    def parity_check(self, frames=3):
        """N > 1: `frames` frames from a fixed pseudo-random field, sharded + exchanged the way the timed steps are,
        against the same frames unsharded on rank 0 (a scratch single-GPU stage): every replica is compared with
        rank 0's unsharded result word for word."""
        torch, dist = self.torch, self.dist
        f0 = (np.random.default_rng(20261018).integers(0, 2 ** 32, self.cells, dtype=np.uint64).astype(np.uint32) | np.uint32(0xC0000000))
        self.st.wait_idle(); dist.barrier()
        self.st.upload_field(f0); self.st.wait_idle(); dist.barrier()
        for f in range(frames):
            self.step_resident(1000 + f)
        self.st.wait_idle(); dist.barrier()
        mine = torch.from_numpy(self.st.download_field().view(np.int32)).to(self.field_t.device)
        ref = torch.empty_like(mine)
        if self.rank == 0:
            import lum_b200
            wx, wy, wz = self.dims
            st1 = lum_b200.RadianceStage((wx, wy, wz), self.scene["bricks"].shape[0], rays_per_probe=self.rays, device=self.device)
            st1.upload_world(self.scene["world"]); st1.upload_bricks(0, self.scene["bricks"]); st1.upload_materials(self.scene["materials"])
            st1.build_requests(); st1.upload_field(f0)
            for f in range(frames):
                st1.update(1000 + f)
            ref = torch.from_numpy(st1.download_field().view(np.int32)).to(self.field_t.device)
            st1.close()
        dist.broadcast(ref, src=0)
        equal = torch.tensor([int(torch.equal(mine, ref))], dtype=torch.int32, device=self.field_t.device)
        differing = torch.tensor([int((mine != ref).sum().item())], dtype=torch.int64, device=self.field_t.device)
        dist.all_reduce(equal, op=dist.ReduceOp.MIN)
        dist.all_reduce(differing, op=dist.ReduceOp.MAX)
        return {"checked": True, "equal": bool(equal.item()), "frames": frames, "replicas_compared": self.n,
                "words_per_replica": self.cells, "max_differing_words": int(differing.item()),
                "against": "the same frames unsharded on rank 0 (lumcu_update, scratch stage), random initial field"}

    def timed(self, fn, steps, warmup, first_seed=0, span=False):
        """W untimed + K timed steps; per-step CUDA events on the stage stream, L2 flushed in
        between (outside the events); returns (total_ms_max_over_ranks, per_step_ms list).
        span=False: total = sum of the per-step event pairs (each step's work, queued back to back).
        span=True (the pipelined e2e steps, whose host side waits on fences): total = first step's start to the
        last step's end minus the L2 flushes (each bracketed by its own event pair), so that any time the GPU sat
        idle waiting for the host is counted."""
        torch, dist = self.torch, self.dist
        self.reset()
        for i in range(warmup):
            fn(first_seed + i)
        self.drain()
        with torch.cuda.stream(self.stream):
            ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
            torch.cuda.synchronize()
            if self.n > 1:
                dist.barrier()
            torch.cuda.synchronize()
            fl = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
            for k in range(steps):
                fl[k][0].record(self.stream)
                self.flush_l2(k)                              # L2 flush on the same stream, outside the events
                fl[k][1].record(self.stream)
                ev[k][0].record(self.stream)
                fn(first_seed + warmup + k)
                ev[k][1].record(self.stream)
            self.drain()
            torch.cuda.synchronize()
            if self.n > 1:
                dist.barrier()
            torch.cuda.synchronize()
        per = [a.elapsed_time(b) for a, b in ev]
        total = float(sum(per))
        if span:
            total = float(ev[0][0].elapsed_time(ev[-1][1]) - sum(a.elapsed_time(b) for a, b in fl[1:]))
        if self.n > 1:
            t = torch.tensor([total], dtype=torch.float64, device=self.field_t.device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            total = float(t.item())
        return total, per

    def algorithmic_bytes(self, steps, warmup, first_seed=0):
        """S,H,N of the timed steps from the instrumented kernel variant (untimed pass)."""
        self.reset()
        import lum_b200
        out = []
        for i in range(warmup + steps):
            self.st.set_debug(lum_b200.DEBUG_PRE_BLEND if i >= warmup else 0)
            if self.fused:
                self.st.update_shard_exchange(first_seed + i, 1, 0)       # the instrumented trace of this rank's shard
                nwg = 1
            elif self.n > 1:
                b, e = self.st.slab_range(1, 0, self.z0, self.z1)
                self.st.update_range(first_seed + i, 1, 0, b, e)
                nwg = e - b
                self.exchange()
            else:
                self.st.update(first_seed + i, 1, 0)
                nwg = self.nreq
            if i >= warmup:
                S, H, N = self.st.download_debug(nwg, pre_blend=False)["totals"]
                out.append(3 * S + 52 * H + 12 * N)
        self.st.set_debug(0)
        return out


def measure(runner, steps, warmup, peak):
    """value / e2e / roofline for one workload on this rank set."""
    st = runner.st
    total_rays = runner.nreq * runner.rays
    # device-resident
    l0 = st.launch_count()
    total_ms, per = runner.timed(runner.step_resident, steps, warmup)
    launches = st.launch_count() - l0 - 0
    launches_timed = launches * steps // (steps + warmup) if (steps + warmup) else 0
    ms_per_step = total_ms / steps
    value = total_rays / (ms_per_step * 1e-3)
    # per-kernel time (profiling events around the trace kernel) + algorithmic bytes
    st.set_profiling(True)
    ks = []
    runner.reset()
    for i in range(warmup + steps):
        runner.flush_l2(i)
        runner.step_resident(i)
        if i >= warmup:
            ks.append(st.last_kernel_ms())
    st.set_profiling(False)
    kernel_ms = float(np.mean(ks))
    kernel_ms_ranks = [kernel_ms]
    if runner.n > 1:                                          # how evenly the shards load the ranks
        t = runner.torch.tensor([kernel_ms], dtype=runner.torch.float64, device=runner.field_t.device)
        out = [runner.torch.empty_like(t) for _ in range(runner.n)]
        runner.dist.all_gather(out, t)
        kernel_ms_ranks = [float(o.item()) for o in out]
    ab = runner.algorithmic_bytes(steps, warmup)
    alg_bytes = float(np.mean(ab))
    kt = kernel_ms if kernel_ms else ms_per_step
    achieved = alg_bytes / (kt * 1e-3) / 1e9
    # end to end
    e2e_ms, _ = runner.timed(runner.step_e2e, steps, warmup, span=True)
    runner.phase_events = []
    span_ms, _ = runner.timed(runner.step_e2e, steps, 1, span=True)
    phases = runner.phase_breakdown()
    runner.phase_events = None
    e2e_value = total_rays / (e2e_ms / steps * 1e-3)
    world_bytes = runner.world_host.nbytes
    return dict(value=value, ms_per_step=ms_per_step, per_step_ms=per, kernel_ms=kernel_ms, kernel_ms_ranks=kernel_ms_ranks,
                alg_bytes=alg_bytes, achieved=achieved, frac=achieved / peak,
                e2e_phases=phases, e2e_value=e2e_value, e2e_ms=e2e_ms / steps, h2d=world_bytes * runner.n,      # every rank uploads the world
                d2h=runner.cells * 4 // runner.n * runner.n, launches=launches_timed,       # each rank reads its slab back
                probes=runner.nreq, rays=total_rays)


def measure_api_config3(frames=64, warmup=8):
    """BASELINE configs[2] through Lum's own C API (include/clum.h, liblum.so): the demo world with two
    mapped models moving every frame -- blockify, world upload, brick copies, map, request rebuild and a
    full radiance update per lum_submit_frame.  Wall clock over `frames` frames, host included."""
    import math
    from lum_b200 import clum, scenes
    d = scenes.demo_scene()
    r = clum.Renderer((48, 48, 16), 15, profile=False)
    r.load_world(d["world"])
    for i in range(1, 15):
        r.load_block(i, d["bricks"][i])
    r.set_material_palette(d["materials"])
    r.upload_block_palette(); r.upload_material_palette()
    r.set_radiance_controller(False, 1)                      # every probe every frame
    ha, hb = r.load_mesh_data(d["model_tank_body"]), r.load_mesh_data(d["model_tank_head"])
    hz = d["model_tank_body"].shape[0]

    def frame(f):
        a = 0.05 * f
        base = (13.1 * 16 + 0.75 * (f % 16), 14.1 * 16 + 0.5 * (f % 16), 3.1 * 16)
        r.start_frame(); r.draw_world()
        r.draw_model(ha, (math.cos(a / 2), 0, 0, math.sin(a / 2)), base)
        r.draw_model(hb, (math.cos(-a), 0, 0, math.sin(-a)), (base[0] + 2, base[1] + 3, base[2] + hz))
        r.prepare_frame(); r.submit_frame()

    for f in range(warmup):
        frame(f)
    r.wait_idle()
    t0 = time.perf_counter()
    for f in range(frames):
        frame(warmup + f)
    r.wait_idle()
    dt = time.perf_counter() - t0
    n = r.request_count()
    r.close()
    return {"workload": "demo world + tank_body/tank_head moving each frame, via lum_submit_frame (BASELINE configs[2])",
            "frames": frames, "ms_per_frame": dt / frames * 1e3, "probes_per_frame": int(n),
            "value": n * RAYS * frames / dt, "unit": UNIT, "timing": "host wall clock, includes the host side of the API"}


def measure_glossy(torch, device, steps=20, warmup=5):
    """Secondary line: the glossy pass (shaders/glossy.frag as lumcu_glossy) over a 1920x1080
    G-buffer of the demo world -- a mirror floor at z = 24 voxels seen by the demo's orthographic start
    camera (internal_types.hpp:54-66), every fragment covered, one mirror ray each.  value = device-resident
    (lumcu_glossy_device on GPU tensors, CUDA events); e2e = lumcu_glossy with host buffers."""
    import ctypes as C
    import lum_b200
    from lum_b200 import scenes
    d = scenes.demo_scene()
    w, h = 1920, 1080
    st = lum_b200.RadianceStage((48, 48, 16), d["bricks"].shape[0], device=device)
    st.upload_world(d["world"]); st.upload_bricks(0, d["bricks"]); st.upload_materials(d["materials"])
    st.build_requests()
    for i in range(8):                                   # a lit field to sample
        st.update(i)
    cd = np.array([0.61, 1.0, -0.8]); cd /= np.linalg.norm(cd)
    plane = np.array([cd[0], cd[1], 0.0]); plane /= np.linalg.norm(plane)
    hz = np.cross(plane, [0, 0, 1.0]); hz /= np.linalg.norm(hz)
    vt = np.cross(cd, hz); vt /= np.linalg.norm(vt)
    vs = np.array([w, h]) / 5.0                          # pixelsInVoxel = 5
    campos = np.array([60.0, 0.0, 194.0])
    view = dict(campos=campos, camdir=cd, horizline_scaled=hz * vs[0] / 2, vertiline_scaled=vt * vs[1] / 2,
                global_light_dir=np.array([0.5, 0.5, -0.9]) / np.linalg.norm([0.5, 0.5, -0.9]), frame_size=[w, h])
    cx = ((np.arange(w) + 0.5) / w * 2 - 1)[None, :]
    cy = ((np.arange(h) + 0.5) / h * 2 - 1)[:, None]
    z = campos[2] + view["horizline_scaled"][2] * cx + view["vertiline_scaled"][2] * cy
    depth = ((z - 24.0) / -cd[2] / 1000.0).astype(np.float32)
    glossy_ids = np.nonzero(d["materials"][:, 5] <= 0.5)[0]
    mat = int(glossy_ids[glossy_ids > 0][0])
    mn = np.zeros((h, w, 4), np.uint8); mn[..., 0] = mat; mn[..., 1:3] = 128; mn[..., 3] = 255
    dev = torch.device(f"cuda:{device}")
    t_mn, t_dp = torch.from_numpy(mn).to(dev), torch.from_numpy(depth).to(dev)
    t_fr = torch.zeros((h, w, 4), dtype=torch.float32, device=dev)
    stream = torch.cuda.ExternalStream(st.stream(), device=device)
    v = st.glossy_view(view)
    torch.cuda.synchronize()

    def resident():
        rc = st.lib.lumcu_glossy_device(st.ctx, C.byref(v), w, h, t_mn.data_ptr(), t_dp.data_ptr(), t_fr.data_ptr(), 0)
        assert rc == 0, st.lib.lumcu_last_error()

    fr_host = np.zeros((h, w, 4), np.float32)

    def e2e():
        rc = st.lib.lumcu_glossy(st.ctx, C.byref(v), w, h, mn.ctypes.data_as(C.c_void_p), depth.ctypes.data_as(C.c_void_p),
                                 fr_host.ctypes.data_as(C.c_void_p), None)
        assert rc == 0, st.lib.lumcu_last_error()

    out = {}
    for name, fn in (("resident", resident), ("e2e", e2e)):
        for _ in range(warmup):
            fn()
        with torch.cuda.stream(stream):
            ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
            torch.cuda.synchronize()
            for k in range(steps):
                ev[k][0].record(stream); fn(); ev[k][1].record(stream)
            torch.cuda.synchronize()
        out[name] = float(np.mean([a.elapsed_time(b) for a, b in ev]))
    covered = int(np.count_nonzero(np.any(fr_host != 0, axis=2)))
    st.close()
    return {"workload": "demo world, 1920x1080 G-buffer (mirror floor at z=24, orthographic demo camera), glossy.frag as lumcu_glossy",
            "fragments": w * h, "covered": covered, "ms_per_frame": out["resident"], "value": w * h / (out["resident"] * 1e-3),
            "unit": "mirror rays/s", "e2e": {"ms_per_frame": out["e2e"], "value": w * h / (out["e2e"] * 1e-3),
                                             "h2d_bytes_per_step": int(mn.nbytes + depth.nbytes + fr_host.nbytes),
                                             "d2h_bytes_per_step": int(fr_host.nbytes)}}


def host_threads():
    """All the host threads this process may use.  torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU arm
    must not inherit that, so the count is passed to the oracle explicitly (num_threads overrides the variable)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def cpu_baseline(scene, budget_s, threads=0, rays=RAYS):
    """The oracle port on the host cores, on a bounded strided sample of the request list."""
    import oracle as O
    threads = threads or host_threads()
    sc = O.Scene(scene["world"], scene["bricks"], scene["materials"], rays_per_probe=rays)
    req = O.build_requests(scene["world"])
    f0 = np.zeros(sc.cells, np.uint32)
    probe = req[:: max(1, len(req) // 2048)][:2048]
    t0 = time.perf_counter()
    r = O.update(sc, f0, probe, 1, nthreads=threads)
    dt = time.perf_counter() - t0
    rate = len(probe) / max(dt, 1e-6)
    n = int(min(len(req), max(len(probe), rate * budget_s)))
    sample = req[:: max(1, len(req) // n)][:n]
    t0 = time.perf_counter()
    r = O.update(sc, f0, sample, 1, nthreads=threads)
    dt = time.perf_counter() - t0
    return dict(value=len(sample) * rays / dt, unit=UNIT, cores=int(r["threads"]), kind="port",
                sample=f"{len(sample)} of {len(req)} probes (every {max(1, len(req) // n)}-th request, "
                       f"time_seed 1, zero field), {dt:.2f} s wall, OpenMP over probes"), sample, sc, f0, len(req)


def run_reference(args, rank):
    """--impl reference: the reference's CPU implementation of the path = the oracle port
    (the reference's own radiance.comp cannot execute here: no Vulkan)."""
    global RAYS
    if rank != 0:
        return
    import oracle as O
    scene = make_scene(args.workload)
    nsteps = args.steps + args.warmup
    budget = max(1.0, min(20.0, 150.0 / max(1, nsteps)))
    base, sample, sc, f0, nprobes = cpu_baseline(scene, budget, rays=args.rays_per_probe)
    RAYS = args.rays_per_probe
    for i in range(args.warmup):
        O.update(sc, f0, sample, i, nthreads=host_threads())
    t0 = time.perf_counter()
    for k in range(args.steps):
        O.update(sc, f0, sample, args.warmup + k, nthreads=host_threads())
    dt = time.perf_counter() - t0
    value = len(sample) * RAYS * args.steps / dt
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": workload_config(args, max(1, args.gpus), nprobes, RAYS),
            "sample_probes": int(len(sample)),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": base["cores"], "kind": "port",
                             "sample": base["sample"]},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    # Informational: the reference's OWN radiance.comp text executed on the CPU (oracle/_ref, built where the
    # reference tree exists).  It is NOT the timed arm: the GLSL shim runs every invocation twice (subgroupAdd
    # is resolved in two passes) and its requests are i8vec4, so only probes with coordinates <= 127 qualify.
    try:
        if O.ref_glsl_available() and RAYS == 64:
            sub = sample[(sample[:, :3].max(axis=1) <= 127)][:256]
            if len(sub):
                t0 = time.perf_counter()
                O.ref_glsl_update(sc, f0, sub, 1)
                dt2 = time.perf_counter() - t0
                line["reference_shader_on_cpu"] = {"value": len(sub) * RAYS / dt2, "unit": UNIT, "probes": int(len(sub)),
                                                   "what": "shaders/radiance.comp itself on oracle/glsl_shim.hpp, OpenMP over workgroups"}
    except Exception as e:
        line["reference_shader_on_cpu"] = {"error": repr(e)}
    print(json.dumps(line), flush=True)


def workload_config(args, world_size, probes, rays):
    """The `config` object of the JSON line -- the same for both arms (the reference arm times a bounded SAMPLE of
    this workload, described in its cpu_baseline.sample; the per-ray metric is what makes the two comparable)."""
    fused = args.exchange == "fused"
    par = "single GPU"
    if world_size > 1:
        par = f"probes sharded x{world_size}: " + (
            ("contiguous equal ranges of the dispatch" if args.shard_mode == "contiguous" else "interleaved chunks of 3072 workgroups")
            + "; one flag barrier per frame, then one kernel reads every texel from its owner's staging array over NVLink and commits it locally"
            if fused else "z-slabs + NCCL field all-gather")
    return {"workload": args.workload, "probes": int(probes), "rays_per_probe": int(rays), "rays_per_step": int(probes) * int(rays),
            "parallelism": par, "l2": "flushed between timed steps (256 MiB fill)", "time_seed": "step index", "initial_field": "zero"}


KERNEL_NAMES = {True: "lumcu::radiance_update_pool_kernel<{r}, {t}, 3, false>", False: "lumcu::radiance_update_wg_kernel<{r}, false>"}


def roofline_object(workload, world_size, rays, m, peak, peak_src, scene, runner):
    """The contract's roofline object for the dominant kernel (the trace kernel), plus what the counters say the
    kernel is really limited by.  achieved/frac: algorithmic bytes (3*S + 52*H + 12*N, S and H counted by the
    instrumented kernel in this run) over the kernel time measured in this run (CUDA events on the stage's stream),
    against the measured HBM peak.  Every other number comes from a file under profiles/ and says which."""
    pooled = m["rays"] // max(1, world_size) >= (1 << 23)
    ev = load_profile_evidence(workload, world_size, rays)
    ro = {"bound": "hbm", "achieved": m["achieved"], "peak": peak, "unit": "GB/s", "frac": m["frac"],
          "traffic": ev.get("dram_bytes_per_launch"), "peak_source": peak_src,
          "kernel": KERNEL_NAMES[pooled].format(r=rays, t=256 if rays == 256 else 128), "kernel_ms": m["kernel_ms"],
          "kernel_ms_per_rank": m["kernel_ms_ranks"],
          "algorithmic_bytes_per_launch": m["alg_bytes"],
          # SURVEY 8d's second figure: what one update must touch at least once
          "compulsory_unique_bytes": int(scene["world"].nbytes + scene["bricks"].nbytes + 6144 + 2 * runner.cells * 4)}
    if ev:
        # the formal bound of the contract is HBM, but the working set is L2-resident: the capture shows what limits it
        ro["limiter"] = ev.get("limiter")
        ro["issue_active"] = ev.get("issue_active")
        ro["lanes_active"] = ev.get("lanes_active")
        ro["l1"] = ev.get("l1")
        ro["l2"] = ev.get("l2")
        ro["evidence"] = ev.get("source")
    return ro


def run_ours(args, rank, local_rank, world_size):
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    if world_size > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    peak, peak_src = load_peaks()
    scene = make_scene(args.workload)
    rays = args.rays_per_probe
    runner = Runner(scene, local_rank, rank, world_size, torch, dist, exchange=args.exchange, rays=rays)
    import lum_b200
    runner.st.set_shard_mode(lum_b200.SHARD_INTERLEAVED if args.shard_mode == "interleaved" else lum_b200.SHARD_CONTIGUOUS)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    m = measure(runner, args.steps, args.warmup, peak)
    clocks = sampler.stop() if rank == 0 else None
    extra = {}
    if world_size > 1:
        # correctness of the sharded + exchanged frames, on the hardware and in the run that is being reported
        par = runner.parity_check()
        if rank == 0:
            par["exchange"] = "fused (NVLink peer stores + flag barrier)" if runner.fused else "nccl all-gather"
            extra["multi_gpu_parity"] = par
    if rank == 0 and world_size == 1 and args.workload != "demo" and not args.no_demo and rays == 64:
        dr = Runner(make_scene("demo"), local_rank, 0, 1, torch, dist)
        dm = measure(dr, max(args.steps, 20), max(args.warmup, 5), peak)
        dr.close()
        try:     # in a child process: the API mirror abort()s on errors (reference behaviour) and must not take the bench down
            p = subprocess.run([sys.executable, os.path.abspath(__file__), "--api-config3"], capture_output=True,
                               text=True, timeout=300, env=dict(os.environ, LUM_B200_DEVICE=str(local_rank)))
            extra["config3_api"] = json.loads(p.stdout.strip().splitlines()[-1])
        except Exception as e:
            extra["config3_api"] = {"error": repr(e)}
        try:
            extra["glossy"] = measure_glossy(torch, local_rank)
        except Exception as e:
            extra["glossy"] = {"error": repr(e)}
        extra["demo"] = {"workload": "demo (BASELINE configs[1]: 48x48x16 demo world, full recompute)",
                         "probes": dm["probes"], "rays": dm["rays"], "value": dm["value"], "unit": UNIT,
                         "ms_per_step": dm["ms_per_step"], "kernel_ms": dm["kernel_ms"],
                         "e2e": {"value": dm["e2e_value"], "unit": UNIT, "ms_per_step": dm["e2e_ms"],
                                 "h2d_bytes_per_step": dm["h2d"], "d2h_bytes_per_step": dm["d2h"]},
                         "roofline": {"bound": "hbm", "achieved": dm["achieved"], "peak": peak, "unit": "GB/s",
                                      "frac": dm["frac"], "traffic": load_profile_evidence("demo", 1, 64).get("dram_bytes_per_launch")}}
    if rank == 0:
        cb = None
        if world_size == 1 and not args.no_cpu:
            cb = cpu_baseline(scene, 12.0, rays=rays)[0]
        line = {"metric": METRIC, "value": m["value"], "unit": UNIT, "n_gpus": world_size,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": m["ms_per_step"],
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic",
                "config": workload_config(args, world_size, m["probes"], rays),
                "e2e": {"value": m["e2e_value"], "unit": UNIT, "ms_per_step": m["e2e_ms"],
                        "h2d_bytes_per_step": m["h2d"], "d2h_bytes_per_step": m["d2h"],
                        "frames_in_flight": 2, "stream_ms_upload_build_update_download": m["e2e_phases"]},
                "gpu_launches": m["launches"],
                "roofline": roofline_object(args.workload, world_size, rays, m, peak, peak_src, scene, runner),
                "clocks": clocks}
        if cb:
            line["cpu_baseline"] = cb
        line.update(extra)
        print(json.dumps(line), flush=True)
    if world_size > 1:
        dist.barrier()                  # nobody unmaps a peer's field while another rank may still store into it
    runner.close()
    if world_size > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="synthetic-256x256x64")
    ap.add_argument("--exchange", default="fused", choices=["fused", "nccl"],
                    help="N>1: how the ranks exchange their slabs of the field each step")
    ap.add_argument("--rays-per-probe", type=int, default=64, choices=[16, 32, 64, 128, 256],
                    help="BASELINE configs[4] sweeps 16-256 (the reference only defines 64); the demo / config-3 / glossy lines need 64")
    ap.add_argument("--shard-mode", default="interleaved", choices=["contiguous", "interleaved"],
                    help="N>1, fused exchange: how the dispatch's workgroups are dealt to the ranks (lumcu_set_shard_mode)")
    ap.add_argument("--no-demo", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--api-config3", action="store_true", help=argparse.SUPPRESS)
    args = ap.parse_args()
    if args.api_config3:
        print(json.dumps(measure_api_config3()), flush=True)
        return
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world_size != args.gpus:
        if args.gpus > 1 and world_size == 1:
            raise SystemExit("for --gpus N>1 launch with torch.distributed.run --nproc-per-node N")
    run_ours(args, rank, local_rank, world_size)


if __name__ == "__main__":
    main()
