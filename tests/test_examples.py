"""The headless demos are the reference's own example programs (examples/demo.cpp, examples/cdemo.c)
retargeted at this repository's API mirror: they must compile as C++17 / C99 against include/ and,
on a GPU box, run."""
import os
import shutil
import subprocess

import numpy as np
import pytest

from voxutil import write_scene, write_vox

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "lum_b200")
BLOCKS = ["dirt", "grass", "grassNdirt", "stone_dirt", "bush", "leaves", "iron", "lamp", "stone_brick",
          "stone_brick_cracked", "stone_pack", "bark", "wood", "planks"]


def build(tmp, src, cc, std):
    out = os.path.join(tmp, os.path.basename(src).split(".")[0])
    cmd = [cc, f"-std={std}", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", src),
           "-L", LIBDIR, "-llum", f"-Wl,-rpath,{LIBDIR}", "-o", out]
    subprocess.run(cmd, check=True, capture_output=True, text=True)
    return out


def test_examples_compile(tmp_path):
    from lum_b200 import clum
    clum.load_library()                      # liblum.so is built
    assert os.path.exists(build(str(tmp_path), "demo_headless.cpp", shutil.which("g++") or "g++", "c++17"))
    assert os.path.exists(build(str(tmp_path), "cdemo_headless.c", shutil.which("gcc") or "gcc", "c99"))


@pytest.mark.gpu
def test_headless_demo_runs(tmp_path, demo):
    assets = tmp_path / "assets"
    assets.mkdir()
    write_scene(str(assets / "scene"), demo["world"])
    for i, n in enumerate(BLOCKS, start=1):
        write_vox(str(assets / f"{n}.vox"), demo["bricks"][i].transpose(2, 1, 0))
    pal = np.full((256, 4), 200, np.uint8)
    write_vox(str(assets / "tank_body.vox"), demo["model_tank_body"], palette_rgba=pal,
              matl={19: {"_type": "_emit", "_emit": 1, "_flux": 4}})
    write_vox(str(assets / "tank_head.vox"), demo["model_tank_head"], palette_rgba=pal)
    exe = build(str(tmp_path), "demo_headless.cpp", shutil.which("g++") or "g++", "c++17")
    p = subprocess.run([exe, str(assets), "12"], capture_output=True, text=True, timeout=120)
    assert p.returncode == 0, p.stderr
    assert "update_radiance" in p.stdout and "frames 12" in p.stdout
    last = p.stdout.strip().splitlines()[-1]
    assert int(last.split("field checksum")[1]) > 0
    assert "requests last frame" in last and int(last.split("requests last frame")[1].split(",")[0]) > 2689
    cexe = build(str(tmp_path), "cdemo_headless.c", shutil.which("gcc") or "gcc", "c99")
    p = subprocess.run([cexe, str(assets / "scene")], capture_output=True, text=True, timeout=120)
    assert p.returncode == 0 and "radiance dispatch" in p.stdout, p.stderr
