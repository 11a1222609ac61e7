#!/usr/bin/env python3
"""Times every tuning build of tools/final_variants.txt (lum_b200/exp_<name>.so, built by tools/final_sweep.sh) and a few
run-time knobs on the C4 world in ONE process (a build = its own dlopen), picks the fastest, and -- if it beats the
shipped build by more than the noise between builds (1.5 %) -- runs the whole GPU test-suite on it.
usage (B200): python tools/final_sweep.py [tag]   -> gpurun_out/<tag>/sweep.json, sweep.md, parity.log"""
import ctypes as C, json, os, subprocess, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lum_b200
from lum_b200 import scenes

args = [a for a in sys.argv[1:] if not a.startswith("--")]
tag = args[0] if args else "sweep"
NO_ENV = "--no-env" in sys.argv       # builds only (the run-time knobs are in profiles/r02_final_sweep.md already)
out = os.path.join(ROOT, "gpurun_out", tag)
os.makedirs(out, exist_ok=True)
STEPS = 6


def bind(path):
    lib = C.CDLL(path)
    for name, (res, args) in lum_b200.SYMBOLS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return lib


cands = [("shipped", os.path.join(ROOT, "lum_b200", "liblumcu.so"), {})]
for line in open(os.path.join(ROOT, "tools", "final_variants.txt")):
    if line.startswith("#") or ":" not in line:
        continue
    n, f = line.rstrip("\n").split(":", 1)
    p = os.path.join(ROOT, "lum_b200", f"exp_{n}.so")
    if os.path.exists(p):
        cands.append((n, p, {"flags": f}))
base_lib = cands[0][1]
for env in () if NO_ENV else ({"LUMCU_CARVEOUT": "50"}, {"LUMCU_CARVEOUT": "65"}, {"LUMCU_CARVEOUT": "100"}, {"LUMCU_POOL_THREADS": "256"},
            {"LUMCU_KERNEL_VARIANT": "0"}):
    cands.append(("env " + " ".join(f"{k}={v}" for k, v in env.items()), base_lib, env))

sc = scenes.synthetic_scene(256, 256, 64)
wz, wy, wx = sc["world"].shape
rows = []
for name, path, extra in cands:
    env = {k: v for k, v in extra.items() if k.startswith("LUMCU_")}
    for k, v in env.items():
        os.environ[k] = v
    try:
        lum_b200._lib = bind(path)
        st = lum_b200.RadianceStage((wx, wy, wz), sc["bricks"].shape[0])
        st.upload_world(sc["world"]); st.upload_bricks(0, sc["bricks"]); st.upload_materials(sc["materials"])
        n = st.build_requests()
        ms = []
        for i in range(STEPS + 2):
            st.update(i)
            if i >= 2:
                ms.append(st.last_update_ms())
        word = int(np.bitwise_xor.reduce(st.download_field().ravel()))      # the same frames: the same field
        st.close()
        rows.append(dict(name=name, lib=os.path.basename(path), flags=extra.get("flags", ""), env=env, probes=n,
                         mean_ms=float(np.mean(ms)), min_ms=float(np.min(ms)), field_xor=word))
    except Exception as e:                                                      # a build that does not run is a result too
        rows.append(dict(name=name, lib=os.path.basename(path), flags=extra.get("flags", ""), env=env, error=str(e)))
    for k in env:
        del os.environ[k]
    print(rows[-1], flush=True)

ok = [r for r in rows if "mean_ms" in r]
base = next(r for r in ok if r["name"] == "shipped")
best = min(ok, key=lambda r: r["mean_ms"])
res = dict(commit=os.environ.get("LUM_COMMIT", "unknown"), workload="synthetic-256x256x64", steps=STEPS, rows=rows,
           shipped_ms=base["mean_ms"], best=best["name"], best_ms=best["mean_ms"], gain=base["mean_ms"] / best["mean_ms"] - 1.0)
json.dump(res, open(os.path.join(out, "sweep.json"), "w"), indent=1)
with open(os.path.join(out, "sweep.md"), "w") as f:
    f.write(f"| build / knob | flags | C4 ms mean | min | vs shipped | field xor |\n|---|---|---|---|---|---|\n")
    for r in rows:
        if "mean_ms" in r:
            f.write(f"| {r['name']} | `{r['flags'] or ' '.join(f'{k}={v}' for k, v in r['env'].items()) or '(default)'}` | {r['mean_ms']:.3f} | {r['min_ms']:.3f} | "
                    f"{(r['mean_ms'] / base['mean_ms'] - 1) * 100:+.1f} % | {'same' if r['field_xor'] == base['field_xor'] else 'DIFFERENT'} |\n")
        else:
            f.write(f"| {r['name']} | `{r['flags']}` | failed: {r['error'][:80]} | | | |\n")
    f.write(f"\nfastest: {best['name']} ({res['gain'] * 100:+.1f} % against the shipped build)\n")
print(open(os.path.join(out, "sweep.md")).read())
if best["name"] != "shipped" and res["gain"] > 0.015 and best["field_xor"] == base["field_xor"]:
    env = dict(os.environ, **best["env"])
    if best["lib"] != "liblumcu.so":
        env["LUMCU_LIB"] = os.path.join(ROOT, "lum_b200", best["lib"])
    # the parity tests proper first (without the whole-C4 test), then the whole-C4 test, then the rest: a call that runs out
    # of time still leaves the most important verdicts behind
    for sel in (["tests/test_parity_gpu.py", "-k", "not full_size"], ["tests/test_parity_gpu.py", "-k", "full_size"],
                ["tests", "--deselect", "tests/test_parity_gpu.py"]):
        t = time.time()
        p = subprocess.run([sys.executable, "-m", "pytest", *sel, "-m", "gpu", "-q", "-x"], cwd=ROOT, env=env, capture_output=True, text=True)
        tail = "\n".join(p.stdout.strip().splitlines()[-4:])
        with open(os.path.join(out, "parity.log"), "a") as f:
            f.write(f"# {best['name']}: pytest {' '.join(sel)} -m gpu (rc {p.returncode}, {time.time() - t:.0f} s)\n{tail}\n")
        print(tail, flush=True)
