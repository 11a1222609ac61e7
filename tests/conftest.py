import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def demo():
    from lum_b200 import scenes
    return scenes.demo_scene()


@pytest.fixture(scope="session")
def oracle_mod():
    import oracle
    oracle.build()
    return oracle


@pytest.fixture(scope="session")
def small_synth():
    from lum_b200 import scenes
    return scenes.synthetic_scene(24, 20, 12, palette=48, seed=1234)


def make_stage(scene, rays=64, reactiveness=0.01, max_requests=0):
    """RadianceStage with a scene dict uploaded (GPU tests only)."""
    import lum_b200
    wz, wy, wx = scene["world"].shape
    st = lum_b200.RadianceStage((wx, wy, wz), scene["bricks"].shape[0], rays_per_probe=rays,
                                reactiveness=reactiveness, max_requests=max_requests)
    st.upload_world(scene["world"])
    st.upload_bricks(0, scene["bricks"])
    st.upload_materials(scene["materials"])
    return st


def oracle_scene(oracle, scene, rays=64, reactiveness=0.01):
    return oracle.Scene(scene["world"], scene["bricks"], scene["materials"], rays_per_probe=rays,
                        reactiveness=reactiveness)
