"""The viewer shim (reference src/main.cu:148-407): the window's controls as a command session of the headless
host, the scene resident between frames, and the display path - the frame written into a device buffer the caller owns, as
Render::run_view writes into the mapped pixel buffer object (include/Render.cuh:446-469)."""
import json
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu(crt):
    if crt.device_count() < 1:
        pytest.fail("no CUDA device: the product has no CPU fallback and these tests need the B200")
    return crt


def test_frame_into_a_device_buffer_equals_the_host_frame(gpu, scene_files):
    import torch
    files = scene_files["cornell-box"]
    cfg = gpu.load_config(files["cfg_path"])
    S = gpu.Scene().add_obj(files["obj"], files["dir"])
    S.set_BVH(cfg.bvh_thresh_n, builder=3, device=0)
    M = gpu.inverse_view_matrix(cfg.eye_pos, cfg.lookat, cfg.up)
    W, H = 200, 150
    R = gpu.Render(S, W, H, 3, cfg.P_RR, cfg.light_sample_n)
    R.run_view(cfg.eye_pos, M, cfg.fovy_rad)
    host = R.get_frame_buffer()
    pbo = torch.zeros(W * H * 3, dtype=torch.uint8, device="cuda:0")       # stands in for the mapped pixel buffer object
    R.frame_to_device(pbo.data_ptr())
    assert np.array_equal(pbo.cpu().numpy().reshape(H, W, 3), host)
    with pytest.raises(gpu.CrtError):
        R.frame_to_device(host.ctypes.data)                                 # host memory is refused, not written through


def test_session_renders_what_separate_runs_render(gpu, scene_files, tmp_path):
    exe = os.path.join(ROOT, "cudaraytracing_b200", "crt")
    cfg_path = scene_files["cornell-box"]["cfg_path"]
    base = [exe, "--config", cfg_path, "--width", "160", "--height", "120"]
    a, b, c = str(tmp_path / "a.png"), str(tmp_path / "b.png"), str(tmp_path / "c.png")
    cfg = gpu.load_config(cfg_path)
    eye2 = [cfg.eye_pos[0] + 0.25, cfg.eye_pos[1], cfg.eye_pos[2] - 0.5]
    script = "\n".join([
        "save %s" % a,                                     # nothing rendered yet: an error line, no file
        "spp 3", "render", "save %s" % a,
        "eye %r %r %r" % tuple(float(v) for v in eye2), "spp 5", "light_sample_n 2", "p_rr 0.7", "render", "save %s" % b,
        "bogus", "quit"]) + "\n"
    r = subprocess.run(base + ["--session"], input=script, capture_output=True, text=True, cwd=str(tmp_path))
    assert r.returncode == 0, r.stderr
    lines = [json.loads(l) for l in r.stdout.strip().splitlines()]
    assert "error" in lines[0] and "error" in lines[-1]
    renders = [l for l in lines if "render_cost_s" in l]
    assert len(renders) == 2 and renders[1]["spp"] == 5
    assert [l["saved"] for l in lines if "saved" in l] == [a, b]
    # the first frame is the one-shot run's; the second differs (camera and parameters changed) and is reproducible
    r1 = subprocess.run(base + ["--spp", "3", "--out", c], capture_output=True, text=True)
    assert r1.returncode == 0, r1.stderr
    assert open(a, "rb").read() == open(c, "rb").read()
    assert open(a, "rb").read() != open(b, "rb").read()
    # the same second frame through the library
    S = gpu.Scene().add_obj(scene_files["cornell-box"]["obj"], scene_files["cornell-box"]["dir"])
    S.set_BVH(cfg.bvh_thresh_n, builder=3, device=0)
    R = gpu.Render(S, 160, 120, 5, 0.7, 2)
    R.run_view(np.asarray(eye2, np.float32), gpu.inverse_view_matrix(np.asarray(eye2, np.float32), cfg.lookat, cfg.up), cfg.fovy_rad)
    ref = str(tmp_path / "ref.png")
    R.save_frame_buffer(ref)
    assert open(b, "rb").read() == open(ref, "rb").read()
    # the Save button's default name: .tmp/<time>.png under the working directory
    r2 = subprocess.run(base + ["--session"], input="spp 1\nrender\nsave\n", capture_output=True, text=True, cwd=str(tmp_path))
    assert r2.returncode == 0, r2.stderr
    saved = json.loads(r2.stdout.strip().splitlines()[-1])["saved"]
    assert saved.startswith(".tmp/") and os.path.exists(os.path.join(str(tmp_path), saved))
