"""Development aid (round 2, session 26): upper bound of what ordering the rays of a C5 batch buys. The rays of the batch are
ordered with torch (cell of the origin on a 2^b grid per axis in Morton order, then direction octant) - library code, outside the
product and outside the timed call - and the product's trace kernel is timed on the unordered and on the ordered buffer."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

def part1by2(x):
    x = x & 0x3ff
    x = (x | (x << 16)) & 0x30000ff
    x = (x | (x << 8)) & 0x300f00f
    x = (x | (x << 4)) & 0x30c30c3
    x = (x | (x << 2)) & 0x9249249
    return x

def main():
    import cudaraytracing_b200 as crt
    n = int(os.environ.get("SP_RAYS", "20000000"))
    cfg = bench.Workload("c5")
    scene, build_ms = cfg.build_scene(crt, 0)
    dev = torch.device("cuda:0")
    st = torch.cuda.current_stream().cuda_stream
    rays = torch.empty((n, 8), dtype=torch.float32, device=dev)
    t_out = torch.empty(n, dtype=torch.float32, device=dev)
    f_out = torch.empty(n, dtype=torch.int32, device=dev)
    for mode, name in ((crt.RAY_CLOSEST, "closest"), (crt.RAY_ANY, "any")):
        scene.random_rays_device(rays.data_ptr(), n, start=0, key=0xC5, any_hit=(mode == crt.RAY_ANY), stream=st)
        torch.cuda.synchronize()
        def run(buf):
            best = 1e30
            for _ in range(4):
                best = min(best, scene.trace_rays_device(buf.data_ptr(), n, mode, t_out.data_ptr(), f_out.data_ptr(), st))
            return best
        ms = run(rays)
        print("%s unordered: %.3f ms = %.1f Mrays/s" % (name, ms, n / ms / 1e3), flush=True)
        if hasattr(crt, "RAY_SORTED"):                # the product's own ordering (key kernel + radix sort + gather + scatter inside the timed call)
            best = min(scene.trace_rays_device(rays.data_ptr(), n, mode | crt.RAY_SORTED, t_out.data_ptr(), f_out.data_ptr(), st) for _ in range(4))
            print("%s CRT_RAY_SORTED (sort inside the call): %.3f ms = %.1f Mrays/s" % (name, best, n / best / 1e3), flush=True)
        if os.environ.get("SP_ONLY_UNORDERED"):
            continue
        lo = rays[:, 0:3].min(0).values
        hi = rays[:, 0:3].max(0).values
        for bits in (3, 5, 7, 10):
            q = ((rays[:, 0:3] - lo) / (hi - lo) * (2 ** bits - 1e-3)).to(torch.int64).clamp_(0, 2 ** bits - 1)
            key = part1by2(q[:, 0]) | (part1by2(q[:, 1]) << 1) | (part1by2(q[:, 2]) << 2)
            for with_oct in (False, True):
                k = key
                if with_oct:
                    o = (rays[:, 4] < 0).to(torch.int64) | ((rays[:, 5] < 0).to(torch.int64) << 1) | ((rays[:, 6] < 0).to(torch.int64) << 2)
                    k = (key << 3) | o
                perm = torch.argsort(k)
                srt = rays[perm].contiguous()
                torch.cuda.synchronize()
                ms = run(srt)
                print("%s cells 2^%d per axis%s: %.3f ms = %.1f Mrays/s" % (name, bits, " + octant" if with_oct else "", ms, n / ms / 1e3), flush=True)
                del srt, perm

if __name__ == "__main__":
    main()
