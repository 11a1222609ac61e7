#!/usr/bin/env python3
"""The store side of a full update, A/B in one process: element-wise commit_kernel (LUMCU_VEC_COMMIT=0) against
commit_vec4_kernel (default).  Reports update - trace kernel = commit kernel + one launch gap (CUDA events of
lumcu_set_profiling), and checks that both stages end with the same field.
usage: commit_ab.py [workload] [steps]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lum_b200
from lum_b200 import scenes

wl = sys.argv[1] if len(sys.argv) > 1 else "synthetic-256x256x64"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 8
if wl == "demo":
    sc = scenes.demo_scene()
else:
    wx, wy, wz = (int(v) for v in wl.split("-")[1].split("x"))
    sc = scenes.synthetic_scene(wx, wy, wz)
wz, wy, wx = sc["world"].shape
fields = {}
for vec in (0, 1):
    os.environ["LUMCU_VEC_COMMIT"] = str(vec)           # read by lumcu_create
    st = lum_b200.RadianceStage((wx, wy, wz), sc["bricks"].shape[0])
    st.upload_world(sc["world"]); st.upload_bricks(0, sc["bricks"]); st.upload_materials(sc["materials"])
    n = st.build_requests()
    st.set_profiling(True)
    rest = []
    for i in range(steps + 2):
        st.update(i)
        if i >= 2:
            rest.append(st.last_update_ms() - st.last_kernel_ms())
    fields[vec] = st.download_field()
    print(f"{wl} LUMCU_VEC_COMMIT={vec} probes={n} update - trace kernel: mean {np.mean(rest) * 1e3:.1f} us  min {np.min(rest) * 1e3:.1f} us")
    st.close()
print("fields equal:", bool(np.array_equal(fields[0], fields[1])))
