"""Deterministic inputs of the glossy-pass tests (shared by the CPU tests, the GPU tests and the golden
generator): scene, radiance field, camera view and a fabricated G-buffer per case."""
import numpy as np

from lum_b200 import scenes

CASES = ("demo", "demo_all", "synth")


def make_case(O, tag):
    rng = np.random.default_rng({"demo": 1, "demo_all": 2, "synth": 3}[tag])
    if tag.startswith("demo"):
        d = scenes.demo_scene()
        mats = d["materials"].copy()
        if tag == "demo_all":
            mats[:, 5] *= 0.5
        scn = dict(world=d["world"], bricks=d["bricks"], materials=mats)
        w, h = 96, 54
        # the demo's start camera (internal_types.hpp:55-56), zoomed out so the whole world is in view
        view = O.camera_view((60, 0, 194), (0.61, 1.0, -0.8), w, h, pixels_in_voxel=0.25)
    else:
        s = scenes.synthetic_scene(24, 24, 8, palette=64)
        mats = s["materials"].copy()
        mats[:, 5] *= 0.7
        scn = dict(world=s["world"], bricks=s["bricks"], materials=mats)
        w, h = 64, 64
        view = O.camera_view((-40, -60, 330), (0.5, 0.7, -0.75), w, h, pixels_in_voxel=0.16)
    sc = O.Scene(scn["world"], scn["bricks"], scn["materials"])
    field = np.zeros(sc.cells, np.uint32) if tag == "demo" else \
        rng.integers(0, 2 ** 32, sc.cells, dtype=np.uint64).astype(np.uint32)
    mat_norm, depth = O.gbuffer(sc, view, w, h)
    return dict(scene=sc, scn=scn, field=field, view=view, mat_norm=mat_norm, depth=depth, width=w, height=h)
