"""Development aid: crt_trace_rays with page-locked and pageable host buffers against the C5 scene for several chunk sizes of the
library's copy / trace / copy pipeline (CRT_BATCH_CHUNK is read when a scene's first host batch is traced: one scene per size)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench

def main():
    import cudaraytracing_b200 as crt
    n = int(os.environ.get("EP_RAYS", "20000000"))
    cfg = bench.Workload("c5")
    dev = torch.device("cuda:0")
    st = torch.cuda.current_stream().cuda_stream
    pin_rays = torch.empty((n, 8), dtype=torch.float32).pin_memory()
    pin_t, pin_f = torch.empty(n, dtype=torch.float32).pin_memory(), torch.empty(n, dtype=torch.int32).pin_memory()
    pageable = None
    for chunk in [int(x) for x in os.environ.get("EP_CHUNKS", "1048576,2097152,3145728,4194304,8388608").split(",")]:
        os.environ["CRT_BATCH_CHUNK"] = str(chunk)
        scene, _ = cfg.build_scene(crt, 0)
        if pageable is None:
            rays = torch.empty((n, 8), dtype=torch.float32, device=dev)
            scene.random_rays_device(rays.data_ptr(), n, start=0, key=0xC5, any_hit=False, stream=st)
            torch.cuda.synchronize()
            pin_rays.copy_(rays)
            del rays
            pageable = np.array(pin_rays.numpy()[: min(n, 8000000)])
        host = pin_rays.numpy()
        scene.trace_rays(host[:1000000], crt.RAY_CLOSEST, out=(pin_t.numpy(), pin_f.numpy()))
        ts = []
        for _ in range(6):
            t0 = time.time(); scene.trace_rays(host, crt.RAY_CLOSEST, out=(pin_t.numpy(), pin_f.numpy())); ts.append(time.time() - t0)
        pt, pf = np.zeros(len(pageable), np.float32), np.zeros(len(pageable), np.int32)
        scene.trace_rays(pageable[:1000000], crt.RAY_CLOSEST, out=(pt, pf))
        tp = []
        for _ in range(4):
            t0 = time.time(); scene.trace_rays(pageable, crt.RAY_CLOSEST, out=(pt, pf)); tp.append(time.time() - t0)
        print("chunk %8d: page-locked %d rays best %.1f median %.1f Mrays/s (H2D %.1f GB/s at best); pageable %d rays best %.1f Mrays/s" % (
            chunk, n, n / min(ts) / 1e6, n / sorted(ts)[len(ts) // 2] / 1e6, 32 * n / min(ts) / 1e9, len(pageable), len(pageable) / min(tp) / 1e6), flush=True)
        del scene
    # what the link gives: one page-locked buffer host to device, nothing else running
    d = torch.empty((n, 8), dtype=torch.float32, device=dev)
    for _ in range(2):
        torch.cuda.synchronize(); t0 = time.time(); d.copy_(pin_rays, non_blocking=True); torch.cuda.synchronize(); dt = time.time() - t0
    print("plain H2D copy of the same buffer: %.1f GB/s" % (32 * n / dt / 1e9))

if __name__ == "__main__":
    main()
