"""Run under torch.distributed.run on N GPUs: every rank renders its shard of the sample-major work index space of a
shipped scene, the int64 accumulation buffers are summed onto rank 0 with ONE NCCL reduce, and rank 0 checks that the
reduced buffer equals its own single-GPU render of the whole frame bit for bit (both estimators, spp >= N and spp < N)."""
import os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import cudaraytracing_b200 as crt
from cudaraytracing_b200 import distributed as cd
from tools import scene_fixture as sf


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    tmp = tempfile.mkdtemp()
    ok = True
    for name in ("cornell-box", "veach-mis"):
        cfg_path = sf.unpack(sf.fixture(name), os.path.join(tmp, name))
        cfg = crt.load_config(cfg_path)
        d = os.path.dirname(cfg_path)
        S = crt.Scene().add_obj(os.path.join(d, cfg.OBJ_paths[0][0]), d)
        S.set_BVH(cfg.bvh_thresh_n, builder=crt.BUILDER_PLOC8, device=local)
        M = crt.inverse_view_matrix(cfg.eye_pos, cfg.lookat, cfg.up)
        for est in (0, 1):
            for spp in (max(world, 4), 1):          # whole-sample shards, and pixel-range shards (spp < world)
                W, H = 400, 300
                R = crt.Render(S, W, H, spp, cfg.P_RR, cfg.light_sample_n)
                R.set_estimator(est)
                R.set_stream(torch.cuda.current_stream().cuda_stream)
                w0, w1 = cd.shard_work(W * H, spp, rank, world)
                R.set_work_range(w0, w1)
                R.run_view(cfg.eye_pos, M, cfg.fovy_rad)
                t = cd.accum_as_tensor(R, dev)
                cd.reduce_accum(t, 0)
                torch.cuda.synchronize()
                if rank == 0:
                    got = R.get_accum_i64().copy()
                    R.clear_range()
                    R.run_view(cfg.eye_pos, M, cfg.fovy_rad)
                    want = R.get_accum_i64()
                    same = np.array_equal(got, want)
                    ok &= same
                    print("%s est %d spp %d on %d GPUs: reduced buffer %s the single-GPU buffer (%d non-zero values)" % (
                        name, est, spp, world, "==" if same else "!=", int((want != 0).sum())), flush=True)
                dist.barrier()
                del R
    dist.destroy_process_group()
    if rank == 0 and not ok:
        raise SystemExit(1)


if __name__ == "__main__":
    main()
