import os, sys, numpy as np
sys.path.insert(0, os.getcwd())
import lum_b200
from lum_b200 import scenes
for wl, dims in (("demo", None), ("s32", (32, 32, 16)), ("s64", (64, 64, 16)), ("s96", (96, 96, 16)), ("s128", (128, 128, 32))):
    sc = scenes.demo_scene() if dims is None else scenes.synthetic_scene(*dims)
    wz, wy, wx = sc["world"].shape
    for R in (16, 64, 256):
        row = []
        for variant in (0, 1):
            st = lum_b200.RadianceStage((wx, wy, wz), sc["bricks"].shape[0], rays_per_probe=R)
            st.upload_world(sc["world"]); st.upload_bricks(0, sc["bricks"]); st.upload_materials(sc["materials"])
            n = st.build_requests(); st.set_kernel_variant(variant)
            ms = []
            for i in range(12):
                st.update(i)
                if i >= 4: ms.append(st.last_update_ms())
            row.append(float(np.median(ms))); st.close()
        print(f"{wl} {wx}x{wy}x{wz} R={R} rays={n*R} wg={row[0]:.4f} pool={row[1]:.4f} ms", flush=True)
