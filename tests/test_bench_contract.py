"""bench.py's JSON contract, the parts that need no GPU: both arms describe the workload with the SAME `config` object
(the reference arm times a bounded sample of it), the reference arm's line carries what the tier asks for, and the
roofline evidence bench.py quotes is a file under profiles/ that names the capture it came from."""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def _args(**kw):
    d = dict(workload="synthetic-256x256x64", exchange="fused", shard_mode="interleaved", gpus=1, rays_per_probe=64)
    d.update(kw)
    return argparse.Namespace(**d)


def test_config_object_is_shared_by_both_arms():
    a = bench.workload_config(_args(), 1, 4191908, 64)
    assert a == {"workload": "synthetic-256x256x64", "probes": 4191908, "rays_per_probe": 64, "rays_per_step": 4191908 * 64,
                 "parallelism": "single GPU", "l2": "flushed between timed steps (256 MiB fill)", "time_seed": "step index",
                 "initial_field": "zero"}
    b = bench.workload_config(_args(gpus=8), 8, 4191908, 64)
    assert b["parallelism"].startswith("probes sharded x8: interleaved chunks")
    assert bench.workload_config(_args(exchange="nccl"), 2, 10, 64)["parallelism"].endswith("NCCL field all-gather")


def test_reference_arm_line(tmp_path):
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--workload", "demo"], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    line = json.loads(p.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == bench.METRIC and line["unit"] == bench.UNIT
    assert line["higher_is_better"] is True and line["gpu_launches"] == 0
    assert line["config"] == bench.workload_config(_args(workload="demo"), 1, 2689, 64)
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and "2689 of 2689 probes" in cb["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": bench.UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_under_torchrun_prints_one_line():
    """The driver launches the reference arm like ours (torchrun for N > 1): rank 0 alone works and prints the line, the
    other ranks exit 0 without work."""
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29591", os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "0", "--workload", "demo"], capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [json.loads(l) for l in p.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    line = lines[0]
    assert line["impl"] == "reference" and line["n_gpus"] == 2 and line["gpu_launches"] == 0
    assert line["config"] == bench.workload_config(_args(workload="demo", gpus=2), 2, 2689, 64)
    assert line["cpu_baseline"]["cores"] == bench.host_threads()     # not torchrun's OMP_NUM_THREADS=1: the count is passed explicitly


def test_roofline_evidence_file():
    ev = json.load(open(os.path.join(ROOT, "profiles", "roofline_evidence.json")))
    for wl in ("synthetic-256x256x64", "synthetic-512x512x64", "demo"):
        e = ev[wl]
        assert e["source"].startswith("profiles/r02_ncu_") and "commit" in e["source"]
        assert os.path.exists(os.path.join(ROOT, e["source"].split(":")[0]))
        assert 0 < e["issue_active"] <= 1 and 0 < e["lanes_active"] <= 32 and e["dram_bytes_per_launch"] > 0
        assert 0 < e["l1"]["frac"] < 1 and 0 < e["l2"]["frac"] < 1
    assert bench.load_profile_evidence("synthetic-256x256x64", 1, 64)["kernel"].startswith("void radiance_update_pool_kernel<64, 128, 3, 0>")
    assert bench.load_profile_evidence("synthetic-256x256x64", 8, 64) == {}          # only single-GPU launches were captured
