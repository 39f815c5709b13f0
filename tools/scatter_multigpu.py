#!/usr/bin/env python
"""BASELINE config 5 check: MC scatter with photons sharded over the GPUs of one box, tallies summed with NCCL.

    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/scatter_multigpu.py [n_photons] [--small]

Verifies that the N-rank tally equals the single-rank tally bit for bit and prints photons/s.
"""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from deepdrr_b200 import Projector, phantoms, scatter  # noqa: E402
from deepdrr_b200.parallel import shard_range  # noqa: E402


class Dev:
    def __init__(self, carm, pose):
        self.source_to_detector_distance = carm.source_to_detector_distance
        self.camera_intrinsics = carm.camera_intrinsics
        self.detector_height, self.detector_width = carm.detector_height, carm.detector_width
        self._pose = pose

    def get_camera_projection(self):
        return self._pose


def main():
    n = int(float(sys.argv[1])) if len(sys.argv) > 1 and not sys.argv[1].startswith("-") else 100_000_000
    small = "--small" in sys.argv
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    vol = phantoms.thorax_volume((128, 128, 100), (3.2, 3.2, 4.0)) if small else phantoms.thorax_volume()
    carm = phantoms.MobileCArmGeometry(sensor_width=384, sensor_height=384, pixel_size=0.776)
    pose = phantoms.c2_poses(1, seed=1, carm=carm)[0]
    with Projector(vol, device=Dev(carm, pose), spectrum="120KV_AL43", neglog=False, scatter_num=n, cuda_device_id=local) as p:
        a, b = shard_range(n, rank, world)
        scatter.simulate(p, pose, 1000, seed=1)  # warm-up
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        mine, counters = scatter.simulate(p, pose, b - a, seed=11, photon_offset=a)
        kernel_ms = p.last_timing_ms()["march"]
        total = scatter.reduce_over_ranks(mine)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if rank == 0:
            img = scatter.scatter_image(total, n, pose)
            p.max_ray_length = carm.max_ray_length
            prim = p._project_batch([pose], want="intensity", raw=True)[0]
            line = f"[scatter] {world} GPU(s), {n:.3g} photons: {dt:.3f} s wall ({n / dt:.3e} photons/s), kernel {kernel_ms:.1f} ms/rank; " \
                   f"scatter/primary at centre {float(img[176:208, 176:208].mean() / prim[176:208, 176:208].mean()):.3f}"
            if world > 1 and n <= 20_000_000:
                whole, _ = scatter.simulate(p, pose, n, seed=11)
                line += f"; N-rank tally == 1-rank tally: {bool(np.array_equal(whole, total))}"
            print(line, flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
