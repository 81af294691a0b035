#!/usr/bin/env python
"""bench.py — headline benchmark: Msamples/s of the path-tracing hot path on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c1|c2|c3] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One step = one run_view of the whole frame (all spp). With N > 1 the samples are sharded over the
ranks (strong scaling: total work fixed), each rank renders its shard against its own scene replica,
and the int64 accumulation buffers are summed onto rank 0 with one NCCL reduce, then resolved.
Prints ONE JSON line on rank 0. See DESIGN.md "Measurement" for the definition of every key.
"""
import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (source, width, height, spp, description)   None -> value from the shipped config
    "c1": ("cornell-box", None, None, None, "C1 cornell-box 800x600 spp=2 light_sample_n=2 bvh_thresh_n=2 P_RR=0.6 (shipped config.json)"),
    "c2": ("veach-mis", None, None, None, "C2 veach-mis 800x600 spp=4 light_sample_n=1 (shipped config.json)"),
    "c3": ("cornell-box", 3840, 2160, 1024, "C3 cornell-box 3840x2160 spp=1024, samples sharded across the GPUs"),
    "c4": ("synthetic", 1920, 1080, 64, "C4 synthetic 10M-triangle heightfield scene 1920x1080 spp=64 (GPU BVH build + deep traversal)"),
    "c5": ("synthetic", 0, 0, 0, "C5 incoherent rays: 100M random rays vs the 10M-triangle BVH, closest-hit and any-hit"),
}


BUILDERS = {"lbvh": 0, "lbvh8": 1, "ploc": 2, "ploc8": 3}     # include/crt.h crt_builder


class Workload:
    """Camera + render settings + how to build the scene (product and oracle side)."""

    def __init__(self, name, small=False):
        import numpy as np
        import cudaraytracing_b200 as crt
        self.name = name
        self.builder = BUILDERS[os.environ.get("CRT_BUILDER", "ploc8")]
        source, W, H, spp, self.desc = WORKLOADS[name]
        self.source = source
        self.tmp = tempfile.mkdtemp(prefix="crt_bench_")
        if source == "synthetic":
            from tools import synthetic as sy
            c = sy.C4_CAMERA
            self.eye, self.lookat, self.up = (np.asarray(c[k], np.float32) for k in ("eye", "lookat", "up"))
            self.fov_y = c["fov_y"]
            self.width, self.height, self.spp = W or c["width"], H or c["height"], spp or c["spp"]
            self.light_sample_n, self.P_RR, self.bvh_thresh_n = c["light_sample_n"], c["P_RR"], c["bvh_thresh_n"]
            self.grid_n = int(os.environ.get("CRT_C4_GRID", "257" if small else str(sy.C4_FULL_N)))
            self.obj = None
        else:
            from tools import scene_fixture as sf
            cfg = crt.load_config(sf.unpack(sf.fixture(source), self.tmp))
            self.eye, self.lookat, self.up, self.fov_y = cfg.eye_pos, cfg.lookat, cfg.up, cfg.fov_y
            self.width, self.height, self.spp = W or cfg.width, H or cfg.height, spp or cfg.spp
            self.light_sample_n, self.P_RR, self.bvh_thresh_n = cfg.light_sample_n, cfg.P_RR, cfg.bvh_thresh_n
            self.obj = os.path.join(self.tmp, cfg.OBJ_paths[0][0])
        self.fovy_rad = np.float32(np.float32(self.fov_y) * np.float32(math.pi) / np.float32(180.0))
        self._arrays = None

    def arrays(self):
        if self._arrays is None:
            from tools import synthetic as sy
            self._arrays = sy.c4_scene(self.grid_n)
        return self._arrays

    def build_scene(self, crt, device):
        if self.obj:
            scene = crt.Scene().add_obj(self.obj, self.tmp)
        else:
            scene = crt.Scene().add_triangles(*self.arrays())
        return scene, scene.set_BVH(self.bvh_thresh_n, builder=self.builder, device=device)

    def build_oracle(self, orc):
        import numpy as np
        if self.obj:
            S = orc.Scene().add_obj(self.obj, self.tmp)
        else:
            v, m, o, mats = self.arrays()
            S = orc.Scene().add_arrays(v, m.astype(np.int32), o.astype(np.int32), mats)
        S.build_new_bvh(self.bvh_thresh_n, self.builder & 2)          # same topology as the product
        if self.builder & 1:
            S.build_wide8(self.bvh_thresh_n, self.builder)
        return S


DATA_NOTE = {
    "fixture": "synthetic camera paths (Philox seed 0) over the reference's cornell-box / veach-mis geometry "
               "(tests/golden/scenes fixtures, regenerated to OBJ/MTL at run time)",
    "synthetic": "procedural scene generated from a seed at run time (tools/synthetic.py), Philox seed 0",
}

S_NODE, S_NODE_WIDE, S_TRI, S_RAY_IO_CLOSEST, S_RAY_IO_ANY = 64, 80, 48, 32 + 8, 48 + 0   # bytes, DESIGN.md "Algorithmic bytes"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1])); smax.append(float(p[2]))
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if p[5 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


# ------------------------------------------------------------------------------------------------
# CPU baseline (oracle port) — bounded sample of the same workload; also yields the per-ray
# node / triangle visit counts the roofline uses (counted by the oracle on the same BVH and rule).
# ------------------------------------------------------------------------------------------------
def cpu_baseline_leg(cfg, budget_samples=float(os.environ.get("CRT_CPU_BUDGET", "4.0e7"))):
    from oracle import orc
    import numpy as np
    S = cfg.build_oracle(orc)
    M = orc.inverse_view_matrix(cfg.eye, cfg.lookat, cfg.up)
    # bounded sample: the full frame at reduced resolution (same camera, same per-ray statistics), 1..spp samples
    scale = 1
    while (cfg.width // scale) * (cfg.height // scale) > budget_samples:
        scale *= 2
    w, h = cfg.width // scale, cfg.height // scale
    spp = max(1, min(cfg.spp, int(budget_samples // (w * h))))
    threads = os.cpu_count() or orc.max_threads()          # torchrun exports OMP_NUM_THREADS=1; use every host core
    t0 = time.time()
    _, st = S.render(cfg.eye, M, float(cfg.fovy_rad), w, h, 0, spp, cfg.P_RR, cfg.light_sample_n, threads=threads,
                      wide=bool(cfg.builder & 1))
    dt = time.time() - t0
    per_ray = {
        "closest_inner": st["closest_inner"] / max(st["closest_rays"], 1), "closest_tris": st["closest_tris"] / max(st["closest_rays"], 1),
        "any_inner": st["any_inner"] / max(st["any_rays"], 1), "any_tris": st["any_tris"] / max(st["any_rays"], 1),
        "rays_per_sample": (st["extend_rays"] + st["shadow_rays"] + st["probe_rays"]) / st["samples"],
    }
    base = {"value": round(w * h * spp / dt / 1e6, 4), "unit": "Msamples/s", "cores": threads, "kind": "port",
            "sample": "%dx%d spp=%d of the workload's camera (oracle/liborc.so, OpenMP), %.1f s" % (w, h, spp, dt)}
    return base, per_ray


def ours(args):
    import numpy as np
    import torch
    import cudaraytracing_b200 as crt
    from cudaraytracing_b200 import distributed as cd

    rank, world, local = dist_env()
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("bench.py --gpus %d must be launched with torch.distributed.run --nproc-per-node %d" % (args.gpus, args.gpus))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    cfg = Workload(args.workload)
    desc = cfg.desc
    npix = cfg.width * cfg.height
    scene, build_ms = cfg.build_scene(crt, local)
    M = crt.inverse_view_matrix(cfg.eye, cfg.lookat, cfg.up)
    render = crt.Render(scene, cfg.width, cfg.height, cfg.spp, cfg.P_RR, cfg.light_sample_n)
    render.set_seed(0)
    stream = torch.cuda.current_stream()
    render.set_stream(stream.cuda_stream)
    w0, w1 = cd.shard_work(npix, cfg.spp, rank, world)
    render.set_work_range(w0, w1)
    accum_t = cd.accum_as_tensor(render, dev)
    frame_host = torch.empty((cfg.height, cfg.width, 3), dtype=torch.uint8).pin_memory()
    frame_np = frame_host.numpy()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    launches = [0]

    def step_device():
        render.run_view(cfg.eye, M, cfg.fovy_rad)
        launches[0] += render.stats()["kernel_launches"]
        cd.reduce_accum(accum_t, 0)

    def step_e2e():
        # host buffers in, host buffer out: camera (13 floats) goes in with the call, the RGB8 frame comes back
        render.run_view(cfg.eye, M, cfg.fovy_rad)
        cd.reduce_accum(accum_t, 0)
        if rank == 0:
            render.get_frame_buffer(frame_np)

    def timed(fn, steps):
        flush.zero_()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.time()
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        barrier()
        wall = (time.time() - t0) * 1e3
        ms = e0.elapsed_time(e1)
        t = torch.tensor([ms, wall], dtype=torch.float64, device=dev)
        if world > 1:
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        return float(t[0]), float(t[1])

    for _ in range(args.warmup):
        step_device()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches[0] = 0
    ms_dev, _ = timed(step_device, args.steps)
    n_launches = launches[0]
    clocks = sampler.stop() if rank == 0 else None
    step_e2e()
    ms_e2e, wall_e2e = timed(step_e2e, args.steps)

    total_samples = npix * cfg.spp
    value = total_samples * args.steps / (ms_dev * 1e3)                 # Msamples/s, whole job
    e2e_value = total_samples * args.steps / (max(ms_e2e, wall_e2e) * 1e3)

    # per-kernel roofline pass (rank 0 only, stage timing on, bounded spp): CUDA events around every stage
    roof, cpu_base, extra = None, None, {}
    if rank == 0:
        peaks, peak_kind = load_peaks()
        cpu_base, per_ray = cpu_baseline_leg(cfg)
        render.clear_range()
        spp_probe = min(cfg.spp, 16)
        render.set_spp(spp_probe)
        render.set_stage_timing(True)
        render.run_view(cfg.eye, M, cfg.fovy_rad)
        st = render.stats()
        render.set_stage_timing(False)
        render.set_spp(cfg.spp)
        s_node = S_NODE_WIDE if cfg.builder & 1 else S_NODE
        b_closest = per_ray["closest_inner"] * s_node + per_ray["closest_tris"] * S_TRI + S_RAY_IO_CLOSEST
        b_any = per_ray["any_inner"] * s_node + per_ray["any_tris"] * S_TRI + S_RAY_IO_ANY
        kernels = {
            "k_extend": {"ms": st["ms_extend"], "rays": st["extend_rays"] + st["probe_rays"], "bytes_per_ray": b_closest},
            "k_shadow": {"ms": st["ms_shadow"], "rays": st["shadow_rays"], "bytes_per_ray": b_any},
        }
        for k in kernels.values():
            k["achieved_gbs"] = k["rays"] * k["bytes_per_ray"] / (k["ms"] * 1e6) if k["ms"] > 0 else 0.0
            k["mrays_s"] = k["rays"] / (k["ms"] * 1e3) if k["ms"] > 0 else 0.0
        dom = max(kernels, key=lambda n: kernels[n]["ms"])
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            with open(tp) as f:
                traffic = json.load(f).get(args.workload, {}).get(dom)
        roof = {"bound": "hbm", "kernel": dom, "achieved": round(kernels[dom]["achieved_gbs"], 1), "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": round(kernels[dom]["achieved_gbs"] / peaks["hbm_gbs"], 4), "traffic": traffic, "peak_kind": peak_kind,
                "traffic_note": ("ncu dram bytes of ONE steady-state launch of this kernel in the committed capture (profiles/r01_s29.md, "
                                 "2^24 paths in flight); launches of a frame this size now carry up to 2^25 paths" if args.workload == "c3" else
                                 "ncu dram bytes of one launch of this kernel in the committed capture of this workload (profiles/), if any"),
                "note": "algorithmic bytes = oracle-counted node (%d B) and triangle (48 B) visits + ray I/O per ray x rays per launch; " % s_node +
                        ("the BVH and triangles (%.0f MB) exceed the 126 MB L2 on this workload: node and triangle fetches are HBM sector traffic" %
                         ((scene.counts()["n_nodes"] * s_node + scene.counts()["n_tris"] * 64) / 1e6) if cfg.source == "synthetic" else
                         "the BVH is L2-resident on this workload, so this is L1/L2-bound work measured against the HBM copy peak"),
                "launches": int(st["iterations"]), "avg_launch_ms": round(kernels[dom]["ms"] / max(st["iterations"], 1), 4),
                "stage_ms": {"generate": round(st["ms_generate"], 3), "extend": round(st["ms_extend"], 3), "shade": round(st["ms_shade"], 3),
                             "shadow": round(st["ms_shadow"], 3), "spp": spp_probe},
                "kernels": {n: {"mrays_s": round(k["mrays_s"], 1), "achieved_gbs": round(k["achieved_gbs"], 1),
                                "bytes_per_ray": round(k["bytes_per_ray"], 1)} for n, k in kernels.items()},
                "per_ray": {k: round(v, 3) for k, v in per_ray.items()}}
        extra = {"bvh_build_gpu_ms": round(build_ms, 3), "triangles": scene.counts()["n_tris"], "nodes": scene.counts()["n_nodes"]}

    if rank == 0:
        out = {
            "metric": "Msamples/s", "value": round(value, 2), "unit": "Msamples/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(ms_dev / args.steps, 3), "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": DATA_NOTE[cfg.source if cfg.source == "synthetic" else "fixture"],
            "config": {"workload": desc, "width": cfg.width, "height": cfg.height, "spp": cfg.spp, "light_sample_n": cfg.light_sample_n,
                       "P_RR": round(float(cfg.P_RR), 4), "bvh_thresh_n": cfg.bvh_thresh_n, "estimator": "compat",
                       "builder": [k for k, v in BUILDERS.items() if v == cfg.builder][0],
                       "parallelism": "samples sharded over %d GPU(s), one int64 NCCL reduce" % world,
                       "l2": "256 MiB buffer written before each timed region (L2 flush); per-step path/shadow queues exceed L2"},
            "e2e": {"value": round(e2e_value, 2), "unit": "Msamples/s", "h2d_bytes_per_step": 13 * 4, "d2h_bytes_per_step": npix * 3,
                    "ms_per_step": round(max(ms_e2e, wall_e2e) / args.steps, 3)},
            "gpu_launches": int(n_launches), "clocks": clocks, "roofline": roof, "cpu_baseline": cpu_base,
        }
        out.update(extra)
        print(json.dumps(out), flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


# ------------------------------------------------------------------------------------------------
# reference arm: the reference's own code (oracle/_ref/libref.so = its headers compiled headless for
# sm_100a): host OBJ load + BVH build on the CPU, then its view_render_kernel on one B200.
# ------------------------------------------------------------------------------------------------
def reference(args):
    import ctypes as C
    import numpy as np
    rank, world, local = dist_env()
    if rank != 0:
        return
    so = os.path.join(ROOT, "oracle", "_ref", "libref.so")
    cfg = Workload(args.workload)
    obj, tmp, desc = cfg.obj, cfg.tmp, cfg.desc
    if obj is None:
        print(json.dumps({"impl": "reference", "unavailable": "C4/C5: the reference's host BVH build (std::sort of 152-byte triangles at every "
                          "level, BVH.h:64-76) over 10M triangles and its 140-byte device triangles are not run in this round"}))
        return
    if not os.path.exists(so):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libref.so not built (needs /root/reference at build time)"}))
        return
    R = C.CDLL(so)
    R.ref_host_load.restype = C.c_void_p
    R.ref_host_load.argtypes = [C.c_char_p, C.c_char_p, C.c_uint, C.c_uint, C.c_uint]
    R.ref_host_times.argtypes = [C.c_void_p, C.c_void_p]
    R.ref_device_init.argtypes = [C.c_void_p]
    R.ref_render.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_uint, C.c_float, C.c_int, C.c_void_p,
                             C.POINTER(C.c_float), C.POINTER(C.c_double)]
    R.ref_inverse_view.argtypes = [C.c_void_p] * 4
    os.environ.setdefault("CUDA_VISIBLE_DEVICES", str(local))
    devnull = os.open(os.devnull, os.O_WRONLY)
    saved = os.dup(1)
    os.dup2(devnull, 1)                      # the reference printf()s from its loaders
    try:
        h = R.ref_host_load(obj.encode(), (tmp + "/").encode(), cfg.width, cfg.height, cfg.bvh_thresh_n)
        host_ms = np.zeros(3)
        R.ref_host_times(h, host_ms.ctypes.data_as(C.c_void_p))
        rc = R.ref_device_init(h)
    finally:
        os.dup2(saved, 1)
    if rc != 0:
        print(json.dumps({"impl": "reference", "unavailable": "reference device init failed rc=%d (per-pixel stacks need %.1f GB)" %
                          (rc, 8712.0 * cfg.width * cfg.height / 1e9)}))
        return
    eye = np.asarray(cfg.eye, np.float32)
    M = np.zeros(9, np.float32)
    R.ref_inverse_view(eye.ctypes.data_as(C.c_void_p), np.asarray(cfg.lookat, np.float32).ctypes.data_as(C.c_void_p),
                       np.asarray(cfg.up, np.float32).ctypes.data_as(C.c_void_p), M.ctypes.data_as(C.c_void_p))
    # bounded sample: the same frame with fewer samples per pixel (the kernel's cost is linear in spp)
    spp = cfg.spp if args.ref_spp <= 0 else min(cfg.spp, args.ref_spp)
    frame = np.zeros(cfg.width * cfg.height * 3, np.uint8)
    kms, wms = C.c_float(), C.c_double()

    def step():
        r = R.ref_render(h, eye.ctypes.data_as(C.c_void_p), M.ctypes.data_as(C.c_void_p), float(cfg.fovy_rad), spp, float(cfg.P_RR),
                         int(cfg.light_sample_n), frame.ctypes.data_as(C.c_void_p), C.byref(kms), C.byref(wms))
        if r != 0:
            raise RuntimeError("reference render failed")
        return kms.value, wms.value

    for _ in range(args.warmup):
        step()
    sampler = ClockSampler(local)
    sampler.start()
    ks, ws = [], []
    for _ in range(args.steps):
        k, w = step()
        ks.append(k); ws.append(w)
    clocks = sampler.stop()
    samples = cfg.width * cfg.height * spp
    value = samples * len(ws) / (sum(ws) * 1e3)
    kernel_only = samples * len(ks) / (sum(ks) * 1e3)
    sample_txt = "full frame %dx%d at spp=%d of %d (cost is linear in spp)" % (cfg.width, cfg.height, spp, cfg.spp)
    print(json.dumps({
        "impl": "reference", "metric": "Msamples/s", "value": round(value, 3), "unit": "Msamples/s", "n_gpus": 1, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(sum(ws) / len(ws), 3), "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "same fixtures as the ours arm, written back to OBJ/MTL and read by the reference's own loader",
        "config": {"workload": desc, "width": cfg.width, "height": cfg.height, "spp": cfg.spp, "light_sample_n": cfg.light_sample_n,
                   "P_RR": round(float(cfg.P_RR), 4), "bvh_thresh_n": cfg.bvh_thresh_n,
                   "what": "view_render_kernel (Render.cuh:330) rebuilt headless for sm_100a, 16x16 blocks, timed like main.cu:370-376 "
                           "(kernel + sync + D2H of the RGB8 frame); single GPU (the reference has no multi-GPU path)"},
        "e2e": {"value": round(value, 3), "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "kernel_only_msamples_s": round(kernel_only, 3),
        "cpu_baseline": {"value": round(value, 3), "unit": "Msamples/s", "cores": 1, "kind": "reference", "sample": sample_txt,
                         "host_ms": {"obj_parse": round(host_ms[0], 1), "load_object": round(host_ms[1], 1), "bvh_build": round(host_ms[2], 1),
                                     "threads": 1, "cpu": cpu_model()}},
        "clocks": clocks}), flush=True)


def ours_c5(args):
    """C5: 100M incoherent rays against the 10M-triangle BVH; value = closest-hit Mrays/s (any-hit reported beside it)."""
    import numpy as np
    import torch
    import cudaraytracing_b200 as crt
    rank, world, local = dist_env()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=dev)
    cfg = Workload("c5")
    scene, build_ms = cfg.build_scene(crt, local)
    n_total = int(os.environ.get("CRT_C5_RAYS", "100000000"))
    n0, n1 = rank * n_total // world, (rank + 1) * n_total // world      # independent rays: shard, no collective
    n = n1 - n0
    stream = torch.cuda.current_stream()
    rays = torch.empty((n, 8), dtype=torch.float32, device=dev)
    t_out = torch.empty(n, dtype=torch.float32, device=dev)
    f_out = torch.empty(n, dtype=torch.int32, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    res = {}
    for mode, name in ((crt.RAY_CLOSEST, "closest"), (crt.RAY_ANY, "any")):
        scene.random_rays_device(rays.data_ptr(), n, start=n0, key=0xC5, any_hit=(mode == crt.RAY_ANY), stream=stream.cuda_stream)
        for _ in range(args.warmup):
            scene.trace_rays_device(rays.data_ptr(), n, mode, t_out.data_ptr(), f_out.data_ptr(), stream.cuda_stream)
        flush.zero_()
        barrier()
        sampler = ClockSampler(local)
        if rank == 0 and mode == crt.RAY_CLOSEST:
            sampler.start()
        kms = 0.0
        for _ in range(args.steps):
            kms += scene.trace_rays_device(rays.data_ptr(), n, mode, t_out.data_ptr(), f_out.data_ptr(), stream.cuda_stream)
        barrier()
        if rank == 0 and mode == crt.RAY_CLOSEST:
            res["clocks"] = sampler.stop()
        t = torch.tensor([kms], dtype=torch.float64, device=dev)
        if world > 1:
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        res[name] = dict(ms=float(t[0]) / args.steps, mrays=n_total * args.steps / (float(t[0]) * 1e3), hit_frac=float((f_out >= 0).float().mean()))
    # e2e through the host-buffer entry point on a bounded batch (rays H2D, results D2H inside the timed call)
    nb = min(n, 10_000_000)
    host_rays = rays[:nb].cpu().numpy()
    t0 = time.time()
    scene.trace_rays(host_rays, crt.RAY_CLOSEST)
    e2e_s = time.time() - t0
    if rank == 0:
        peaks, peak_kind = load_peaks()
        # algorithmic bytes from an oracle count on a bounded sample of the same rays / same BVH
        from oracle import orc
        O = cfg.build_oracle(orc)
        sample = rays[:200000].cpu().numpy()
        which = 4 if cfg.builder & 1 else 0
        _, _, st = O.trace(sample, which=which, mode=0, want_stats=True, threads=os.cpu_count())
        t0 = time.time()
        O.trace(sample, which=which, mode=0, threads=os.cpu_count())
        cpu_dt = time.time() - t0
        bpr = st["inner"] / st["rays"] * (S_NODE_WIDE if cfg.builder & 1 else S_NODE) + st["tris"] / st["rays"] * S_TRI + S_RAY_IO_CLOSEST
        ach = res["closest"]["mrays"] * 1e6 * bpr / 1e9
        traffic = None                     # ncu dram bytes of one captured launch, scaled to this launch's ray count
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            with open(tp) as f:
                t5 = json.load(f).get("c5", {})
            if t5.get("k_trace_batch") and t5.get("rays"):
                traffic = int(t5["k_trace_batch"] * (n / t5["rays"]))
        print(json.dumps({
            "metric": "Mrays/s (closest-hit)", "value": round(res["closest"]["mrays"], 1), "unit": "Mrays/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(res["closest"]["ms"], 3), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": DATA_NOTE["synthetic"],
            "config": {"workload": cfg.desc, "rays": n_total, "triangles": scene.counts()["n_tris"], "nodes": scene.counts()["n_nodes"],
                       "l2": "BVH + triangles (%.0f MB) and the ray buffers exceed L2; 256 MiB flush before the timed region" %
                             ((scene.counts()["n_nodes"] * 64 + scene.counts()["n_tris"] * 64) / 1e6)},
            "any_hit": {"mrays_s": round(res["any"]["mrays"], 1), "ms_per_step": round(res["any"]["ms"], 3), "blocked_frac": round(res["any"]["hit_frac"], 4)},
            "closest_hit_frac": round(res["closest"]["hit_frac"], 4),
            "e2e": {"value": round(nb / e2e_s / 1e6, 1), "unit": "Mrays/s", "h2d_bytes_per_step": nb * 32, "d2h_bytes_per_step": nb * 8,
                    "note": "crt_trace_rays with host buffers on a %d-ray batch" % nb},
            "gpu_launches": 2 * args.steps, "clocks": res.get("clocks"),
            "roofline": {"bound": "hbm", "kernel": "k_trace_batch<closest>", "achieved": round(ach, 1), "peak": peaks["hbm_gbs"], "unit": "GB/s",
                         "frac": round(ach / peaks["hbm_gbs"], 4), "traffic": traffic, "peak_kind": peak_kind,
                         "per_ray": {"inner": round(st["inner"] / st["rays"], 2), "tris": round(st["tris"] / st["rays"], 2), "bytes": round(bpr, 1)}},
            "cpu_baseline": {"value": round(len(sample) / cpu_dt / 1e6, 3), "unit": "Mrays/s", "cores": os.cpu_count(), "kind": "port",
                             "sample": "%d of the same rays, oracle traversal, %.1f s" % (len(sample), cpu_dt)},
            "bvh_build_gpu_ms": round(build_ms, 3)}), flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


def cpu_model():
    try:
        with open("/proc/cpuinfo") as f:
            for ln in f:
                if ln.startswith("model name"):
                    return ln.split(":", 1)[1].strip() + " x%d" % os.cpu_count()
    except OSError:
        pass
    return "unknown x%d" % (os.cpu_count() or 0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--builder", default=None, choices=sorted(BUILDERS), help="BVH topology + node layout (default: CRT_BUILDER or ploc8)")
    ap.add_argument("--ref-spp", type=int, default=4, help="reference arm: samples per pixel of the bounded sample (<=0: full)")
    args = ap.parse_args()
    if args.builder:
        os.environ["CRT_BUILDER"] = args.builder
    if args.impl == "reference":
        reference(args)
    elif args.workload == "c5":
        ours_c5(args)
    else:
        ours(args)


if __name__ == "__main__":
    main()
