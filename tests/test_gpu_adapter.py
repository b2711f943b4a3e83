"""The C++ host adapter (adapter/B200Render : oka::Render) driven like a Strelka client, on the GPU."""
import os
import struct
import subprocess

import numpy as np
import pytest

from conftest import rel_rmse
from oracle import pyoracle
from strelka_b200 import _abi
from strelka_b200.scenes import make_cornell

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _write_vec(f, arr):
    f.write(struct.pack("<Q", len(arr)))
    f.write(np.ascontiguousarray(arr).tobytes())


def test_cpp_adapter_renders_like_oracle(tmp_path):
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "adapter"), "-s"])
    w, h, spp, depth = 64, 48, 6, 4
    scene, settings, _ = make_cornell(w, h, spp, depth=depth)
    a = scene.arrays()
    cam = scene.getCamera(0)
    mats = np.zeros(len(a["materials"]), dtype=[("mx", "<u4"), ("color", "<f4", (3,)), ("rough", "<f4"), ("metal", "<f4")])
    mats["mx"] = a["materials"]["model"] == _abi.SB_MATERIAL_USD_PREVIEW_SURFACE
    mats["color"] = a["materials"]["base_color"]
    mats["rough"] = a["materials"]["roughness"]
    mats["metal"] = a["materials"]["metallic"]
    path = tmp_path / "scene.bin"
    with open(path, "wb") as f:
        f.write(struct.pack("<4I", w, h, spp, depth))
        f.write(struct.pack("<8f", *cam.position, *cam.orientation, cam.fov))
        for k in ("vertices", "indices", "meshes", "instances", "lights"):
            _write_vec(f, a[k])
        _write_vec(f, mats)
    out = tmp_path / "out.raw"
    r = subprocess.run([os.path.join(ROOT, "adapter", "_build", "adapter_test"), str(path), str(out)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    img = np.fromfile(out, dtype=np.float32).reshape(h, w, 4)
    ref, _, _, _ = pyoracle.OracleScene(scene).render(settings, w, h, spp)
    assert rel_rmse(img, ref) <= 1e-3


def test_adapter_built_against_the_reference_headers_renders_like_oracle(tmp_path):
    """oracle/_ref/adapter_ref_test = B200Render.cpp compiled against the reference's OWN render / scene / settings
    headers with the reference's scene.cpp and camera.cpp (built where the reference tree exists, travels to the GPU
    box): the scene goes through the real oka::Scene API, the image must be the oracle's."""
    from util import random_scene

    tool = os.path.join(ROOT, "oracle", "_ref", "adapter_ref_test")
    if not os.path.exists(tool):
        pytest.skip("oracle/_ref/adapter_ref_test has not been built (needs the reference tree)")
    w, h, spp, depth = 64, 48, 6, 4
    for scene, settings in (make_cornell(w, h, spp, depth=depth)[:2], random_scene(seed=3)):
        settings.setAs("render/pt/sppTotal", spp)
        settings.setAs("render/pt/depth", depth)
        j, out = tmp_path / "journal.bin", tmp_path / "out.raw"
        scene.write_journal(j, w, h, spp, depth)
        ref, _, _, _ = pyoracle.OracleScene(scene).render(settings, w, h, spp)
        assert ref[..., :3].mean() > 0
        for mode in ("render", "sharded"):  # B200Render::render, and joinGroup + renderSharded with a group of one
            r = subprocess.run([tool, mode, str(j), str(out)], capture_output=True, text=True)
            assert r.returncode == 0, mode + ": " + r.stdout + r.stderr
            img = np.fromfile(out, dtype=np.float32).reshape(h, w, 4)
            assert rel_rmse(img, ref) <= 1e-3, mode
