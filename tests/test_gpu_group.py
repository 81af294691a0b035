"""Several GPUs behind one handle (crt_group, include/crt.h): the reference renders on device 0 only (src/main.cu:99-100).

The reduced fixed-point buffer must be bit-identical to the single-GPU render's (and therefore to the oracle's) for every
number of GPUs, for spp >= N (whole samples per GPU) and spp < N (pixel ranges of a sample), for both estimators; the PNG of
`crt --gpus N` must be the file `crt` writes. The N > 1 cases need a box with that many GPUs (gpurun --gpus N); on a one-GPU
box they are skipped and the N = 1 group (same code path: begin / step / finish state machine, no collective) still runs."""
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


def _scene(gpu, files, builder=3):
    cfg = gpu.load_config(files["cfg_path"])
    S = gpu.Scene().add_obj(files["obj"], files["dir"])
    S.set_BVH(cfg.bvh_thresh_n, builder=builder, device=0)
    return cfg, S, gpu.inverse_view_matrix(cfg.eye_pos, cfg.lookat, cfg.up)


@pytest.mark.parametrize("n_gpus", [1, 2, 4, 8])
@pytest.mark.parametrize("name,est,spp", [("veach-mis", 0, 5), ("veach-mis", 1, 3), ("cornell-box", 0, 1)])
def test_group_buffer_equals_the_single_gpu_buffer(gpu, scene_files, name, est, spp, n_gpus):
    if gpu.device_count() < n_gpus:
        pytest.skip("needs %d GPUs" % n_gpus)
    cfg, S, M = _scene(gpu, scene_files[name])
    W, H = 200, 150
    R = gpu.Render(S, W, H, spp, cfg.P_RR, cfg.light_sample_n)
    R.set_estimator(est)
    R.run_view(cfg.eye_pos, M, cfg.fovy_rad)
    want, frame = R.get_accum_i64(), R.get_frame_buffer()
    st1 = R.stats()
    G = gpu.RenderGroup(S, W, H, list(range(n_gpus)), spp, cfg.P_RR, cfg.light_sample_n, 0, est)
    for _ in range(2):                                            # a second frame through the same handles
        G.run_view(cfg.eye_pos, M, cfg.fovy_rad)
        assert np.array_equal(G.get_accum_i64(), want)
        assert np.array_equal(G.get_frame_buffer(), frame)
    per_gpu = [G.stats(k) for k in range(n_gpus)]
    assert sum(s["samples"] for s in per_gpu) == W * H * spp
    assert sum(s["extend_rays"] for s in per_gpu) == st1["extend_rays"] and sum(s["shadow_rays"] for s in per_gpu) == st1["shadow_rays"]
    if n_gpus > 1:
        assert all(s["samples"] > 0 for s in per_gpu) and per_gpu[0]["reduce_ms"] > 0
    G.close()


def test_group_rejects_bad_arguments(gpu, scene_files):
    cfg, S, M = _scene(gpu, scene_files["veach-mis"])
    with pytest.raises(gpu.CrtError):
        gpu.RenderGroup(S, 64, 64, [])
    with pytest.raises(gpu.CrtError):
        gpu.RenderGroup(S, 64, 64, [0, 0])
    with pytest.raises(gpu.CrtError):
        gpu.RenderGroup(S, 64, 64, [gpu.device_count()])
    unbuilt = gpu.Scene().add_obj(scene_files["veach-mis"]["obj"], scene_files["veach-mis"]["dir"])
    with pytest.raises(gpu.CrtError):
        gpu.RenderGroup(unbuilt, 64, 64, [0])


@pytest.mark.parametrize("n_gpus", [2, 8])
def test_cli_gpus_writes_the_same_png(gpu, scene_files, tmp_path, n_gpus):
    if gpu.device_count() < n_gpus:
        pytest.skip("needs %d GPUs" % n_gpus)
    exe = os.path.join(ROOT, "cudaraytracing_b200", "crt")
    base = [exe, "--config", scene_files["cornell-box"]["cfg_path"], "--width", "320", "--height", "240", "--spp", "6"]
    one, many = str(tmp_path / "one.png"), str(tmp_path / "many.png")
    r = subprocess.run(base + ["--out", one], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run(base + ["--out", many, "--gpus", str(n_gpus)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    info = json.loads(r.stdout.strip().splitlines()[-1])
    assert info["gpus"] == n_gpus and info["reduce_ms"] > 0
    assert open(one, "rb").read() == open(many, "rb").read()
