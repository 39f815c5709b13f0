#!/usr/bin/env python
"""Benchmark of the projection hot path (BASELINE.json metric: DRRs/s, 512x512x400 CT -> 1536^2 detector).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--views-per-step B]

One "step" = one batch of B views of BASELINE config 2 (synthetic 512x512x400 thorax CT, 3 materials,
120 kV spectrum, random MobileCArm poses, 1536^2 detector) per rank.  N > 1: one process per GPU
(torchrun), every rank holds a replica of the volume and projects its own B views per step -- no
data-path collective; weak scaling.  Prints ONE JSON line on rank 0.

value  = views / s with volumes resident in HBM and images left in device memory (CUDA events on the
         stream the kernels run on, max over ranks).
e2e    = same through the public API call a user makes, ``images = projector.project(*poses)`` on host pose
         objects -> host images: pose math, pose upload, kernels, images D2H (into the projector's page-locked
         result pool), all inside the timed region.
--impl reference times the reference's own unmodified CUDA kernel (oracle/_ref, compiled from
/root/reference by oracle/Makefile) driven the way the reference's Python drives it: per view five
small H2D uploads, one launch, two blocking D2H copies, two host transposes, host neglog
(projector.py:679-702, 786-831).  The reference has no CPU implementation of this path; if oracle/_ref
is absent the CPU oracle port is timed instead.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SHAPE, SPACING = (512, 512, 400), (0.8, 0.8, 1.0)
SPECTRUM = "120KV_AL43"
STEP_MM = 0.1
WORKLOAD = "C2: synthetic 512x512x400 CT, 3 materials, 120KV_AL43, random MobileCArm poses, 1536x1536 detector, step 0.1 mm"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "250"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()
                self.proc.wait()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


_PROBE = ("import ctypes,sys\nl=ctypes.CDLL('libcuda.so.1')\nr=l.cuInit(0)\nn=ctypes.c_int(0)\nr2=l.cuDeviceGetCount(ctypes.byref(n)) if r==0 else -1\n"
          "print(r,r2,n.value)\nsys.exit(0 if (r==0 and r2==0 and n.value>0) else 3)")


def _wait_for_cuda(max_wait_s=150.0):
    """Bounded wait until a fresh process can initialise CUDA.  The driver starts this script seconds after another GPU
    process (the reference arm) has exited; a GPU that is still tearing down / re-initialising then fails cuInit once and
    PyTorch caches that failure for the life of the process -- so the probe runs in a child and is retried."""
    t0, last = time.time(), ""
    while True:
        try:
            r = subprocess.run([sys.executable, "-c", _PROBE], capture_output=True, text=True, timeout=60)
            last = (r.stdout + r.stderr).strip()[-300:]
            if r.returncode == 0:
                return True, last, time.time() - t0
        except Exception as e:  # probe hung or could not start
            last = repr(e)[:300]
        if time.time() - t0 > max_wait_s:
            return False, last, time.time() - t0
        time.sleep(3.0)


def _cuda_or_die(args, impl):
    ok, last, waited = _wait_for_cuda()
    if ok:
        return waited
    smi = ""
    try:
        smi = subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True, timeout=30).stdout.strip()[:300]
    except Exception as e:
        smi = repr(e)[:200]
    print(json.dumps({"impl": impl, "error": "CUDA did not initialise within the bounded wait", "waited_s": round(waited, 1),
                      "probe": last, "nvidia_smi_L": smi, "n_gpus": args.gpus}), flush=True)
    sys.exit(1)


def _dist_setup(n_gpus):
    import torch

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist

        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    elif n_gpus > 1:
        raise SystemExit("launch with torchrun for --gpus > 1")
    return rank, world, local


def _max_over_ranks(x, world):
    if world == 1:
        return x
    import torch
    import torch.distributed as dist

    t = torch.tensor([x], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def _sum_over_ranks(x, world):
    if world == 1:
        return x
    import torch
    import torch.distributed as dist

    t = torch.tensor([x], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def _barrier(world):
    import torch

    if world > 1:
        import torch.distributed as dist

        dist.barrier()
    torch.cuda.synchronize()


def _peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


def _ncu_capture():
    """The newest committed ncu summary of this command (profiles/rNN_bench_ncu.json, written by tools/ncu_summary.py)."""
    for name in ("r02_bench_ncu.json", "r01_bench_ncu.json"):
        try:
            cap = json.load(open(os.path.join(ROOT, "profiles", name)))
            cap["_file"] = "profiles/" + name
            return cap
        except Exception:
            continue
    return {}


def cpu_baseline_port(volume, st, carm, pose, crop=256):
    """The CPU oracle (oracle/drr_oracle.c, OpenMP over all host cores) on a centred crop of one C2 view."""
    from deepdrr_b200 import geo
    from oracle import cpu_oracle

    cores = os.cpu_count() or 1
    W = H = 1536
    u0 = (W - crop) // 2
    w2i, src, ijk = geo.pose_arrays(pose, [volume])
    t0 = time.perf_counter()
    r = cpu_oracle.project([volume.data], st.labels, st.M, W, H, STEP_MM, w2i, src, ijk, carm.max_ray_length, st.energies, st.pdf, st.mu,
                           u0=u0, v0=u0, sub=1, Ws=crop, Hs=crop, want_area=False, want_steps=True, nthreads=cores)
    dt = time.perf_counter() - t0
    rays_per_s = crop * crop / dt
    return {"value": rays_per_s / (W * H), "unit": "DRRs/s", "cores": cores, "kind": "port",
            "sample": f"{crop}x{crop} centred pixel crop of one C2 view ({int(r.steps.sum()):d} ray steps, {dt:.1f} s), extrapolated per ray to 1536x1536",
            "rays_per_s": rays_per_s}


def secondary_configs(ct, device, sampler, n_views=16):
    """BASELINE configs 3 and 4 (not the headline metric; parity is covered by tests/): march and whole-projection time per
    384x384 view for CT + two K-wire volumes and for CT + a 50k-triangle screw mesh, device timers of the library."""
    from deepdrr_b200 import Projector, phantoms
    from deepdrr_b200.vol import Mesh

    poses, sdd = phantoms.cone_poses(n_views)
    out = {}
    sv, sf = phantoms.screw_mesh()
    screw = Mesh(sv, sf, material="titanium")
    phantoms.place_kwire(screw, (-20.0, -60.0, 10.0), (0.2, 1.0, 0.1))
    carve = Mesh(sv, sf, material="titanium", subtractive=True, layer=1)
    phantoms.place_kwire(carve, (-20.0, -60.0, 10.0), (0.2, 1.0, 0.1))
    scenes = {"C3 CT + 2 K-wire volumes": (phantoms.c3_scene(ct=ct), n_views),
              "C4 CT + additive screw mesh": ([ct, screw], 4), "C4 CT + subtractive screw mesh": ([ct, carve], 4)}
    for name, (objs, n) in scenes.items():
        with Projector(objs, spectrum=SPECTRUM, step=STEP_MM, neglog=True, camera_intrinsics=poses[0].intrinsic,
                       source_to_detector_distance=sdd, cuda_device_id=device, sampler=sampler) as p:
            p.project(*poses[:n])                                   # warm-up at the same batch size (buffers are sized on first use)
            l0 = p.launch_count()
            p.project(*poses[:n])
            tm = p.last_timing_ms()
            out[name] = {"views": n, "sensor": [384, 384], "march_ms_per_view": tm["march"] / n, "total_ms_per_view": tm["total"] / n,
                         "gpu_launches": p.launch_count() - l0}
    return out


def run_ours(args):
    waited = _cuda_or_die(args, "ours")
    import torch

    from deepdrr_b200 import Projector, geo, phantoms
    from deepdrr_b200.scene import SceneTables

    rank, world, local = _dist_setup(args.gpus)
    torch.cuda.set_device(local)
    B = args.views_per_step
    carm = phantoms.MobileCArmGeometry()
    W, H = carm.sensor_width, carm.sensor_height
    volume = phantoms.thorax_volume(SHAPE, SPACING)
    n_pose = 1000
    poses = phantoms.c2_poses(n_pose, seed=1, carm=carm)
    p = Projector(volume, spectrum=SPECTRUM, step=STEP_MM, neglog=True, camera_intrinsics=carm.camera_intrinsics,
                  source_to_detector_distance=carm.source_to_detector_distance, cuda_device_id=local, sampler=args.sampler)
    p.initialize()
    stream = torch.cuda.current_stream()
    p.set_stream(stream.cuda_stream)
    mrl = carm.max_ray_length

    def step_poses(s):
        base = (s * world + rank) * B
        return [poses[(base + i) % n_pose] for i in range(B)]

    # ---- device-resident throughput ("value"): kernel-level pose arrays in, images left in HBM ---------
    out_dev = torch.empty((B, H, W), dtype=torch.float32, device="cuda")
    arrays = [geo.pose_arrays_batch(step_poses(s), [volume]) for s in range(args.warmup + args.steps)]
    for s in range(args.warmup):
        p.project_arrays(*arrays[s], (W, H), mrl, want="intensity", out=out_dev)
    launches0 = p.launch_count()
    clocks = ClockSampler(local)
    _barrier(world)
    if rank == 0:
        clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    march_ms, steps_total, window_total = [], 0, 0
    e0.record(stream)
    for s in range(args.warmup, args.warmup + args.steps):
        p.project_arrays(*arrays[s], (W, H), mrl, want="intensity", out=out_dev)
        march_ms.append(p.last_timing_ms()["march"])
        steps_total += p.last_sample_count()
        window_total += p.last_window_samples()
    e1.record(stream)
    _barrier(world)
    dev_ms = _max_over_ranks(e0.elapsed_time(e1), world)
    launches = p.launch_count() - launches0
    total_views = B * args.steps * world
    value = total_views / (dev_ms * 1e-3)

    # ---- end to end through the public API ("e2e"): images = projector.project(*poses) -------------------
    for s in range(min(args.warmup, 3)):
        img = p.project(*step_poses(s), max_ray_length=mrl)
    _barrier(world)
    t0 = time.perf_counter()
    checksum = None
    for s in range(args.warmup, args.warmup + args.steps):
        img = p.project(*step_poses(s), max_ray_length=mrl)      # host poses in, host [B, H, W] images out
        if checksum is None:
            checksum = float(np.mean(img[0], dtype=np.float64))   # same definition on the reference arm: its first timed view
    _barrier(world)
    e2e_s = _max_over_ranks(time.perf_counter() - t0, world)
    e2e_value = total_views / e2e_s
    clk = clocks.stop() if rank == 0 else None
    h2d = B * (9 + 3 + 12 + 9) * 4                               # one ViewDev record per view (poses + inverse matrix)
    d2h = B * H * W * 4

    # ---- roofline of the dominant kernel (ray march) ------------------------------------------------
    peaks, peak_src = _peaks()
    bytes_view = SHAPE[0] * SHAPE[1] * SHAPE[2] * 5 + W * H * 8          # 543 162 368 B
    avg_march_s = float(np.mean(march_ms)) * 1e-3
    achieved = bytes_view * B / avg_march_s / 1e9
    march_s_total = sum(march_ms) * 1e-3
    window_per_s = _sum_over_ranks(window_total / march_s_total, world)
    cap = _ncu_capture()
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"],
                "traffic": cap.get("march_dram_bytes_per_launch"), "traffic_source": cap.get("_file"),
                "kernel": "march_warp_kernel", "peak_source": peak_src,
                "note": "the march is a cache-resident gather (TEX / issue bound), so the HBM fraction is small by construction (SURVEY.md 8(d)); "
                        "binding-resource figures are in 'binding'"}
    binding = {"march_ms_per_view": float(np.mean(march_ms)) / B,
               "march_steps_per_view": steps_total / (B * args.steps),            # all steps the reference takes (K.cu:334), incl. the empty ones
               "window_samples_per_view": window_total / (B * args.steps),        # steps inside the volume: the samples that fetch density
               "window_samples_per_s": window_per_s,
               "gather_GBps_at_40B_per_sample": window_per_s * 40 / 1e9}
    if cap.get("tex_lane_fetches_per_launch"):
        # texture-unit load: fetches per view counted by ncu (SASS TEX instructions x active lanes) x views/s, against the unit's
        # measured peak of 1.98 trilinear fp32 fetches / clk / SM (tools/tex_rate.cu on a B200: 5.5e11 /s)
        per_view = cap["tex_lane_fetches_per_launch"] / cap.get("views_per_launch", B)
        rate = per_view * B / avg_march_s
        binding["tex_fetches_per_s"] = rate
        binding["tex_peak_fetches_per_s"] = 5.5e11
        binding["tex_frac_of_peak"] = rate / 5.5e11
    if cap:  # what binds, from the committed ncu capture of this command (stale if the kernel changed since: see its "head")
        binding["from_committed_capture"] = {k: cap.get(k) for k in ("_file", "head", "kernel", "issue_slot_utilisation", "tex_request_cycles_pct", "pipe_fma_pct",
                                                                     "pipe_alu_pct", "shared_pipe_wavefronts_pct", "avg_active_lanes", "l1tex_hit_pct",
                                                                     "l2_hit_pct", "warps_active_pct", "registers_per_thread", "l1tex_throughput_pct",
                                                                     "l1tex_tex_data_pipe_wavefronts_pct", "l1tex_filter_wavefronts_pct",
                                                                     "l1tex_lsu_data_pipe_wavefronts_pct", "l2_GBps", "dram_GBps")}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:  # the CPU baseline is reported at N = 1 only
        st = SceneTables([volume], SPECTRUM)
        cpu = cpu_baseline_port(volume, st, carm, poses[0], crop=args.cpu_crop)
    del img
    p.free()
    secondary = None
    if rank == 0 and world == 1 and args.secondary:
        secondary = secondary_configs(volume, local, args.sampler)
    if rank == 0:
        line = {"metric": "DRRs/s", "value": value, "unit": "DRRs/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "rays_per_s": value * W * H,
                "config": {"workload": WORKLOAD, "views_per_step_per_gpu": B, "views_total": total_views, "sensor": [W, H], "volume": list(SHAPE),
                           "step_mm": STEP_MM, "sampler": args.sampler,
                           "cache": "inputs larger than L2 (4.2 GB of cell records + 0.5 GB volume per GPU)",
                           "parallelism": f"views sharded over {world} GPU(s), volume replicated, no collective"},
                "e2e": {"value": e2e_value, "unit": "DRRs/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "call": "Projector.project(*poses) -> host ndarray"},
                "checksum": checksum, "gpu_launches": int(launches), "roofline": roofline, "binding": binding, "cpu_baseline": cpu, "clocks": clk,
                "cuda_init_wait_s": round(waited, 1)}
        if secondary is not None:
            line["secondary"] = secondary
        print(json.dumps(line), flush=True)
    if world > 1:
        import torch.distributed as dist

        dist.destroy_process_group()


def run_c5(args):
    """BASELINE config 5: Monte Carlo scatter, 1e8 photons per view sharded over the ranks, tallies summed with one NCCL all-reduce.
    One step = one view; strong scaling (the photons of a view are split).  Rank 0 also checks, outside the timed region, that the
    N-rank tally of the first timed view equals the tally of the same photons simulated on one GPU alone, bit for bit."""
    waited = _cuda_or_die(args, "ours")
    import torch

    from deepdrr_b200 import Projector, phantoms, scatter

    rank, world, local = _dist_setup(args.gpus)
    torch.cuda.set_device(local)
    n_photons = int(args.photons)
    volume = phantoms.thorax_volume(SHAPE, SPACING)
    poses, sdd = phantoms.cone_poses(args.warmup + args.steps, seed=6)

    class Dev:
        source_to_detector_distance = sdd
        camera_intrinsics = poses[0].intrinsic
        detector_height = detector_width = 384 * 0.3

        def get_camera_projection(self):
            return poses[0]

    p = Projector(volume, device=Dev(), spectrum=SPECTRUM, step=STEP_MM, neglog=False, scatter_num=n_photons, cuda_device_id=local,
                  coefficient_records=False)                       # the scatter kernel reads the raw density / label arrays only
    p.initialize()
    for s in range(args.warmup):
        scatter.simulate_sharded(p, poses[s], n_photons, seed=s)
    clocks = ClockSampler(local)
    _barrier(world)
    if rank == 0:
        clocks.start()
    t0 = time.perf_counter()
    kernel_ms, first = [], None
    for s in range(args.warmup, args.warmup + args.steps):
        tally, _ = scatter.simulate_sharded(p, poses[s], n_photons, seed=s)
        kernel_ms.append(p.last_timing_ms()["march"])
        if first is None:
            first = tally
    _barrier(world)
    dt = _max_over_ranks(time.perf_counter() - t0, world)
    clk = clocks.stop() if rank == 0 else None
    kms = _max_over_ranks(float(np.mean(kernel_ms)), world)
    same = None
    if rank == 0:
        alone, _ = scatter.simulate(p, poses[args.warmup], n_photons, seed=args.warmup)   # the same photon ids on this GPU alone
        same = bool(np.array_equal(alone, first))
    p.free()
    if rank == 0:
        value = n_photons * args.steps / dt
        print(json.dumps({"metric": "scatter photons/s", "value": value, "unit": "photons/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                          "ms_per_step": dt * 1e3 / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
                          "data": "synthetic",
                          "config": {"workload": "C5: Monte Carlo scatter through the 512x512x400 CT, 120KV_AL43, 384x384 detector", "photons_per_view": n_photons,
                                     "views_total": args.steps, "parallelism": f"photons of a view sharded over {world} GPU(s), tallies all-reduced (NCCL)"},
                          "scatter_kernel_ms_per_view_per_rank": kms, "n_rank_tally_equals_1_rank_tally": same, "gpu_launches": int(args.steps),
                          "clocks": clk, "cuda_init_wait_s": round(waited, 1)}), flush=True)
        if same is False:
            sys.exit("the sharded tally differs from the single-GPU tally")
    if world > 1:
        import torch.distributed as dist

        dist.destroy_process_group()


def run_c3(args):
    """BASELINE config 3: CT + two overlapping K-wire volumes, 384x384 sensor, 10 000 views over the GPUs (weak scaling per step:
    every rank projects its own B views per step; `views_total` = B x steps x ranks).  Timed through the public call."""
    waited = _cuda_or_die(args, "ours")
    import torch

    from deepdrr_b200 import Projector, phantoms

    rank, world, local = _dist_setup(args.gpus)
    torch.cuda.set_device(local)
    B = args.views_per_step if args.views_per_step != 8 else 100
    volumes = phantoms.c3_scene()
    poses, sdd = phantoms.cone_poses(2000, seed=7)
    p = Projector(volumes, spectrum=SPECTRUM, step=STEP_MM, neglog=True, camera_intrinsics=poses[0].intrinsic, source_to_detector_distance=sdd,
                  cuda_device_id=local, sampler=args.sampler)
    p.initialize()

    def step_poses(s):
        base = (s * world + rank) * B
        return [poses[(base + i) % len(poses)] for i in range(B)]

    for s in range(args.warmup):
        p.project(*step_poses(s))
    launches0 = p.launch_count()
    clocks = ClockSampler(local)
    _barrier(world)
    if rank == 0:
        clocks.start()
    t0 = time.perf_counter()
    march_ms, checksum = [], None
    for s in range(args.warmup, args.warmup + args.steps):
        img = p.project(*step_poses(s))
        march_ms.append(p.last_timing_ms()["march"])
        if checksum is None:
            checksum = float(np.mean(img[0], dtype=np.float64))
    _barrier(world)
    dt = _max_over_ranks(time.perf_counter() - t0, world)
    clk = clocks.stop() if rank == 0 else None
    launches = p.launch_count() - launches0
    del img
    p.free()
    if rank == 0:
        total = B * args.steps * world
        value = total / dt
        print(json.dumps({"metric": "DRRs/s", "value": value, "unit": "DRRs/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                          "ms_per_step": dt * 1e3 / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                          "data": "synthetic",
                          "config": {"workload": "C3: 512x512x400 CT + two 21x21x2000 K-wire volumes (0.1 mm), 120KV_AL43, 384x384 sensor, step 0.1 mm",
                                     "views_per_step_per_gpu": B, "views_total": total, "sensor": [384, 384],
                                     "parallelism": f"views sharded over {world} GPU(s), volumes replicated, no collective"},
                          "e2e": {"value": value, "unit": "DRRs/s", "h2d_bytes_per_step": B * 3 * (9 + 3 + 12 + 9) * 4, "d2h_bytes_per_step": B * 384 * 384 * 4,
                                  "call": "Projector.project(*poses) -> host ndarray"},
                          "march_ms_per_view": float(np.mean(march_ms)) / B, "checksum": checksum, "gpu_launches": int(launches), "clocks": clk,
                          "cuda_init_wait_s": round(waited, 1)}), flush=True)
    if world > 1:
        import torch.distributed as dist

        dist.destroy_process_group()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return  # rank 0 alone runs the reference arm
    from oracle import ref_gpu

    have_gpu_ref = ref_gpu.available()
    if have_gpu_ref:
        ok, _, _ = _wait_for_cuda(60.0)
        have_gpu_ref = ok
    from deepdrr_b200 import geo, phantoms
    from deepdrr_b200.scene import SceneTables

    B = args.views_per_step
    carm = phantoms.MobileCArmGeometry()
    W, H = carm.sensor_width, carm.sensor_height
    volume = phantoms.thorax_volume(SHAPE, SPACING)
    st = SceneTables([volume], SPECTRUM)
    poses = phantoms.c2_poses(1000, seed=1, carm=carm)
    config = {"workload": WORKLOAD, "views_per_step_per_gpu": B, "views_total": B * args.steps, "sensor": [W, H], "volume": list(SHAPE),
              "step_mm": STEP_MM}
    if have_gpu_ref:
        ref = ref_gpu.RefProjector([volume.data], st.labels, st.M, [volume.spacing])
        ref.set_spectrum(st.energies, st.pdf, st.mu)
        clocks = ClockSampler(0)

        def one_step(s):
            imgs, pps = [], []
            for i in range(B):
                pose = poses[(s * B + i) % 1000]
                w2i, src, ijk = geo.pose_arrays(pose, [volume])                       # projector.py:802-831
                inten, pp, ms = ref.project(W, H, STEP_MM, w2i, src, ijk, carm.max_ray_length, threads=8, transpose=True)  # :709-800
                imgs.append(inten); pps.append(pp)
                kernel_ms.append(ms)
            images = np.stack(imgs)                                                   # projector.py:687-688
            _ = np.stack(pps)
            # utils.neglog on the host (utils/image_utils.py:18-59), NumPy float32 as in the reference
            images = images + (images.min(axis=(1, 2), keepdims=True) + np.float32(0.01))
            images = -np.log(images)
            lo, hi = images.min(axis=(1, 2), keepdims=True), images.max(axis=(1, 2), keepdims=True)
            return (images - lo) / (hi - lo)

        kernel_ms = []
        for s in range(args.warmup):
            one_step(s)
        kernel_ms.clear()
        clocks.start()
        t0 = time.perf_counter()
        checksum = None
        for s in range(args.warmup, args.warmup + args.steps):
            out = one_step(s)
            if checksum is None:
                checksum = float(np.mean(out[0], dtype=np.float64))                   # first timed view, as on the other arm
        dt = time.perf_counter() - t0
        clk = clocks.stop()
        value = B * args.steps / dt
        line = {"impl": "reference", "metric": "DRRs/s", "value": value, "unit": "DRRs/s", "n_gpus": 1, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": dt * 1e3 / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": value, "unit": "DRRs/s", "cores": 1, "kind": "reference",
                                 "sample": f"{B * args.steps} full C2 views; reference CUDA kernel (unmodified, oracle/_ref) on the B200 with the "
                                           "reference's per-view host flow on one host thread -- the reference has no CPU implementation of this path"},
                "e2e": {"value": value, "unit": "DRRs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "reference_kernel_ms_per_view": float(np.mean(kernel_ms)), "reference_kernel_only_DRRs_per_s": 1e3 / float(np.mean(kernel_ms)),
                "clocks": clk, "checksum": checksum}
        ref.close()
        if args.secondary:
            try:  # the reference kernel on config 3 (CT + two K-wire volumes, 384x384); config 4 needs its OpenGL renderer
                vols = phantoms.c3_scene(ct=volume)
                st3 = SceneTables(vols, SPECTRUM)
                poses3, sdd3 = phantoms.cone_poses(4)
                ref3 = ref_gpu.RefProjector([v.data for v in vols], st3.labels, st3.M)
                ref3.set_spectrum(st3.energies, st3.pdf, st3.mu)
                ms3 = []
                for i, pose in enumerate(poses3):
                    w2i, src, ijk = geo.pose_arrays(pose, vols)
                    _, _, ms = ref3.project(384, 384, STEP_MM, w2i, src, ijk, 4 * sdd3, priority=st3.priorities)
                    if i > 0:
                        ms3.append(ms)
                ref3.close()
                line["secondary"] = {"C3 CT + 2 K-wire volumes": {"views": len(ms3), "sensor": [384, 384], "reference_kernel_ms_per_view": float(np.mean(ms3))}}
            except Exception as e:  # a missing (V, M) cubin must not break the headline line
                line["secondary"] = {"error": str(e)[:200]}
        ref_gpu.shutdown()                                                            # everything freed and the device idle before the next arm starts
        print(json.dumps(line), flush=True)
        return
    # no reference cubin (or no GPU): time the CPU oracle port on a bounded sample per step
    vals = []
    for s in range(args.warmup + args.steps):
        c = cpu_baseline_port(volume, st, carm, poses[s % 1000], crop=192)
        if s >= args.warmup:
            vals.append(c)
    value = float(np.mean([c["value"] for c in vals]))
    line = {"impl": "reference", "metric": "DRRs/s", "value": value, "unit": "DRRs/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 / value, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config, "cpu_baseline": {"value": value, "unit": "DRRs/s", "cores": vals[0]["cores"], "kind": "port", "sample": vals[0]["sample"]},
            "e2e": {"value": value, "unit": "DRRs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--views-per-step", type=int, default=8)
    ap.add_argument("--sampler", default="hybrid", choices=["hybrid", "alu", "tex"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--secondary", action="store_true", help="append C3 / C4 timings under 'secondary' (extra projectors after the headline run)")
    ap.add_argument("--no-secondary", action="store_true", help="accepted for compatibility (the default now)")
    ap.add_argument("--cpu-crop", type=int, default=768, help="side of the centred pixel crop the CPU oracle marches for cpu_baseline")
    ap.add_argument("--config", default="c2", choices=["c2", "c3", "c5"],
                    help="c2: the headline projection metric; c3: CT + two K-wire volumes, 384x384; c5: Monte Carlo scatter photons/s")
    ap.add_argument("--photons", type=float, default=1e8, help="--config c5: photons per view")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.config == "c5":
        run_c5(args)
    elif args.config == "c3":
        run_c3(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
