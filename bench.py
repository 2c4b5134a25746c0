#!/usr/bin/env python
"""bench.py — the path-tracing hot path on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload c3|c2|c1|c5]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

Workload (default) = BASELINE.json config 3, the one the metric is quoted on: "Dragon" at 1920x1080, 64 spp, 8 bounces,
HDR skysphere with env-map importance sampling + BRDF importance sampling + MIS. The reference ships neither the Dragon
OBJ nor the HDR (.MISSING_LARGE_BLOBS), so the Dragon-class procedural stand-in of SURVEY.md Appendix A is used
(1 000 002 triangles, gold metal roughness 0.25, 2048x1024 sun+sky) — `data` says so.

A step = one full frame: every rank renders its interleaved 16x16 tiles (wavefront integrator by default: wf_shade /
wf_trace kernel pairs over 4 tile groups on 4 streams), the tile buffers are gathered to rank 0 (NCCL) and un-tiled there. value = INTERSECT_SCENE-equivalent rays of the whole frame / step time
(CUDA events on the launching stream, max over ranks). Total work is fixed as N grows => "scaling": "strong".
One JSON line on stdout (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (width, height, spp, bounces, description)
    "c1": (512, 512, 16, 4, "cornell_pbr.obj 512x512 16spp 4 bounces, constant 1e-20 env (BASELINE config 1)"),
    "c2": (1920, 1080, 1, 1, "1M-triangle displaced sphere, 1920x1080 un-jittered primary rays (BASELINE config 2)"),
    "c3": (1920, 1080, 64, 8, "Dragon-class stand-in (1,000,002 tris, gold metal r=0.25), 1920x1080 64spp 8 bounces, "
                              "2048x1024 sun+sky env-map IS + BRDF IS + MIS (BASELINE config 3)"),
    "c5": (3840, 2160, 1024, 8, "20M-triangle procedural scene 3840x2160 1024spp 8 bounces (BASELINE config 5)"),
}


def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=3)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="b200", choices=["b200", "reference"])
    p.add_argument("--workload", default="c3", choices=list(WORKLOADS))
    p.add_argument("--spp", type=int, default=0, help="override samples per pixel (0 = the workload's)")
    p.add_argument("--width", type=int, default=0)
    p.add_argument("--height", type=int, default=0)
    p.add_argument("--integrator", type=int, default=1, help="0 = megakernel, 1 = wavefront (default)")
    p.add_argument("--flags", type=int, default=0)
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--cpu-sample", default="", help="WxHxSPP of the bounded CPU sample (default per workload)")
    p.add_argument("--verify", action="store_true", help="N>1: rank 0 also renders the frame alone and asserts the gathered frame is bit-identical")
    return p.parse_args()


def load_workload(name, args):
    from sycl_ray_tracing_b200 import scenes
    import sycl_ray_tracing_b200 as rt
    w, h, spp, bounces, desc = WORKLOADS[name]
    if name == "c1":
        g = np.load(os.path.join(ROOT, "tests", "golden", "scenes.npz"))
        s = dict(tri9=g["cornell_tri9"], mat_idx=g["cornell_mat_idx"], mats10=g["cornell_mats10"], emissive=g["cornell_emissive"],
                 env=rt.constant_env(1.0e-20), camera=rt.Camera.CORNELL_BOX_CAMERA)
    elif name == "c2":
        s = scenes.c2_scene()
        s["env"] = rt.constant_env(1.0)
    elif name == "c3":
        s = scenes.c3_scene()
    else:
        s = scenes.c5_scene()
    w = args.width or w
    h = args.height or h
    spp = args.spp or spp
    return s, w, h, spp, bounces, desc


class ClockSampler(threading.Thread):
    """Samples SM clocks and throttle reasons of one GPU during the timed region (pynvml)."""

    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x10: "sync_boost",
               0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz, self.ok = index, [], set(), False, None, False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        while not self.stop_flag:
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                mask = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, name in self.REASONS.items():
                    if mask & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.05)

    def result(self):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def profile_traffic(workload):
    """dram bytes per launch of the dominant kernel from the committed ncu capture, if one exists for this workload."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(path):
        try:
            return json.load(open(path)).get(workload)
        except Exception:
            return None
    return None


def cpu_sample_shape(name, args, w, h, spp):
    if args.cpu_sample:
        a = [int(v) for v in args.cpu_sample.lower().split("x")]
        return a[0], a[1], a[2]
    if name == "c3":
        return 768, 432, 4          # 4/25 of the pixels, 1/16 of the spp: ~10-30 s of reference CPU work on 16 cores
    if name == "c5":
        return 192, 108, 1
    return w, h, spp                # c1 / c2 run in full on the CPU


def host_threads():
    """All host cores (torchrun exports OMP_NUM_THREADS=1 to its workers, so the count is passed explicitly)."""
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def run_cpu_sample(oracle_scene, cam17, name, sw, sh, sspp, bounces):
    """One pass of the reference CPU implementation over the bounded sample on all host cores; returns seconds."""
    if name == "c2":
        _, _, sec = oracle_scene.primary(cam17, sw, sh, mode=0, nthreads=host_threads())
        return sec
    _, sec = oracle_scene.render(cam17, sw, sh, sspp, bounces, nthreads=host_threads())
    return sec


def make_oracle_scene(s, need_env=True):
    from oracle.oracle import best_oracle
    o = best_oracle()
    sc = o.scene_from_arrays(s["tri9"], s["mat_idx"], s["mats10"], s["emissive"])
    if need_env:
        sc.set_env(s["env"])
    return o, sc


def port_ray_count(s, cam17, name, sw, sh, sspp, bounces):
    """INTERSECT_SCENE-equivalent query count of the sample (the reference cannot count; the port restatement can)."""
    if name == "c2":
        return sw * sh
    from oracle.oracle import PortOracle
    po = PortOracle()
    ps = po.scene_from_arrays(s["tri9"], s["mat_idx"], s["mats10"], s["emissive"])
    ps.set_env(s["env"])
    _, _, rays = po.render_counted(ps, cam17, sw, sh, sspp, bounces, nthreads=host_threads())
    return rays


def reference_arm(args):
    """--impl reference: the reference's own CPU implementation (oracle/_ref when it was compiled, else the port) on all
    host cores, same workload config / metric / unit; each step is a bounded sample of the frame. Rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    s, w, h, spp, bounces, desc = load_workload(args.workload, args)
    sw, sh, sspp = cpu_sample_shape(args.workload, args, w, h, spp)
    cam17 = s["camera"].as_array17()
    o, sc = make_oracle_scene(s)
    rays = port_ray_count(s, cam17, args.workload, sw, sh, sspp, bounces)
    for _ in range(args.warmup):
        run_cpu_sample(sc, cam17, args.workload, min(sw, 96), min(sh, 54), 1, bounces)      # warm caches / thread pool only
    secs = [run_cpu_sample(sc, cam17, args.workload, sw, sh, sspp, bounces) for _ in range(args.steps)]
    sec = sum(secs) / len(secs)
    value = rays / sec / 1e6
    sample = f"{sw}x{sh} px x {sspp} spp of the {w}x{h} x {spp} spp frame per step ({rays} rays)"
    line = {
        "impl": "reference", "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "spp_per_s": sw * sh * sspp / sec,
        "config": {"workload": desc, "width": w, "height": h, "spp": spp, "max_bounces": bounces, "sample": sample},
        "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": host_threads(), "kind": o.kind, "sample": sample},
        "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    args = parse_args()
    if args.impl == "reference":
        reference_arm(args)
        return

    import torch
    import torch.distributed as dist
    import sycl_ray_tracing_b200 as rt
    from sycl_ray_tracing_b200 import distributed as D
    from sycl_ray_tracing_b200 import scenes

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"          # NCCL would print its version banner on stdout: stdout carries exactly one JSON line
        dist.init_process_group("nccl", device_id=device)

    s, w, h, spp, bounces, desc = load_workload(args.workload, args)
    t_build = time.perf_counter()
    scene = rt.Scene(s["tri9"], s["mat_idx"], s["mats10"], s["emissive"], skysphere=s.get("env"), device=local_rank)
    build_s = time.perf_counter() - t_build
    cam = s["camera"]
    n_tri = len(s["tri9"])
    bytes_per_ray = scenes.algorithmic_bytes_per_ray(n_tri)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=device)      # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    primary_only = args.workload == "c2"
    if primary_only:
        prim = torch.empty((h, w), dtype=torch.int32, device=device)
        tbuf = torch.empty((h, w), dtype=torch.float32, device=device)

        def step():
            st = torch.cuda.current_stream(device).cuda_stream
            scene.trace_primary_device(cam, w, h, prim.data_ptr(), tbuf.data_ptr(), stream_ptr=st, rank=rank, world=world, flags=args.flags)
        rays_per_frame = w * h
        kernel_name = "k_primary"
        launches_per_step = 1
    else:
        g = D.make_cuda_gatherer(scene, cam, w, h, spp, bounces, rank, world, device, integrator=args.integrator, flags=args.flags)
        step = g.frame
        # rays of the whole frame (deterministic): counted once on an untimed pass
        tiles_probe = torch.empty_like(g.tiles)
        st = scene.render_tiles_device(cam, w, h, spp, bounces, tiles_probe.data_ptr(), stream_ptr=torch.cuda.current_stream(device).cuda_stream,
                                       integrator=args.integrator, flags=args.flags, rank=rank, world=world, want_stats=True)
        rays_t = torch.tensor([st["rays"]], dtype=torch.int64, device=device)
        if world > 1:
            dist.all_reduce(rays_t)
        rays_per_frame = int(rays_t.item())
        del tiles_probe
        kernel_name = "k_pathtrace_mega" if args.integrator == 0 else "wf_trace_coop + wf_shade, all launches of the frame"
        launches_per_step = int(st["gpu_launches"]) + (1 if rank == 0 else 0)

    for _ in range(args.warmup):
        step()
    barrier()
    verified = None
    if args.verify and not primary_only:
        img = step()
        if rank == 0:
            solo = D.make_cuda_gatherer(scene, cam, w, h, spp, bounces, 0, 1, device, integrator=args.integrator, flags=args.flags).frame()
            verified = bool(torch.equal(img.view(torch.int32), solo.view(torch.int32)))
            if not verified:
                raise SystemExit(f"bench.py --verify: the {world}-rank frame differs from the 1-rank frame")
        barrier()

    sampler = ClockSampler(local_rank)
    sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    for i in range(args.steps):
        flush.fill_(i & 0xff)                 # L2 flush between timed iterations (outside the event pair)
        ev[i][0].record()
        step()
        ev[i][1].record()
    barrier()
    sampler.stop_flag = True
    sampler.join(timeout=2)
    step_ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    ms_per_step = float(total_ms.item()) / args.steps
    value = rays_per_frame / (ms_per_step * 1e-3) / 1e6

    # dominant kernel alone (CUDA events on its launching stream), for the roofline
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kms = []
    probe = torch.empty((D.tiles_for_rank(w, h, 0, world) * 256, 4), dtype=torch.float32, device=device)
    for i in range(args.steps):
        flush.fill_(i & 0xff)
        stp = torch.cuda.current_stream(device).cuda_stream
        k0.record()
        if primary_only:
            scene.trace_primary_device(cam, w, h, prim.data_ptr(), tbuf.data_ptr(), stream_ptr=stp, rank=rank, world=world, flags=args.flags)
        else:
            scene.render_tiles_device(cam, w, h, spp, bounces, probe.data_ptr(), stream_ptr=stp, integrator=args.integrator, flags=args.flags,
                                      rank=rank, world=world)
        k1.record()
        torch.cuda.synchronize()
        kms.append(k0.elapsed_time(k1))
    kernel_ms = torch.tensor([sum(kms) / len(kms)], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(kernel_ms, op=dist.ReduceOp.MAX)
    kernel_ms = float(kernel_ms.item())
    peak, peak_src = measured_peak()
    rays_per_launch = rays_per_frame / world
    achieved = rays_per_launch * bytes_per_ray / (kernel_ms * 1e-3) / 1e9

    # the dominant kernel by itself: one more frame with every trace / shade launch bracketed by CUDA events on its launching
    # stream (B200RT_FLAG_TIME_KERNELS; the tile groups then run one after the other, so this frame is slower than a timed step)
    dominant = None
    if not primary_only and args.integrator == 1:
        flush.fill_(7)
        kst = scene.render_tiles_device(cam, w, h, spp, bounces, probe.data_ptr(), stream_ptr=torch.cuda.current_stream(device).cuda_stream,
                                        integrator=args.integrator, flags=args.flags | rt.FLAG_TIME_KERNELS, rank=rank, world=world, want_stats=True)
        if kst["trace_launches"] > 0 and kst["trace_ms"] > 0:
            t_ach = kst["rays"] * bytes_per_ray / (kst["trace_ms"] * 1e-3) / 1e9
            dominant = {"kernel": "wf_trace_coop" if not (args.flags & (rt.FLAG_BVH2 | rt.FLAG_DIAG_SLABS | rt.FLAG_SIMPLE_TRACE)) else "wf_trace",
                        "launches": kst["trace_launches"], "avg_launch_ms": kst["trace_ms"] / kst["trace_launches"],
                        "rays_per_launch": kst["rays"] / kst["trace_launches"], "achieved": t_ach, "frac": t_ach / peak,
                        "share_of_kernel_time": kst["trace_ms"] / (kst["trace_ms"] + kst["shade_ms"]),
                        "shade_avg_launch_ms": kst["shade_ms"] / max(1, kst["shade_launches"])}

    # end to end through the public host API: host buffers in, host framebuffer out, copies inside the timed region
    e2e = None
    if world == 1:
        fb = rt.Image(w, h).pixels
        times = []
        stats = None
        for i in range(max(1, min(args.steps, 3))):
            fb[...] = (0.0, 0.0, 0.0, 1.0)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            if primary_only:
                _, _, stats = scene.trace_primary(cam, w, h, flags=args.flags)
            else:
                _, stats = scene.render(cam, w, h, spp, bounces, framebuffer=fb, integrator=args.integrator, flags=args.flags)
            times.append(time.perf_counter() - t0)
        e_sec = sum(times) / len(times)
        e2e = {"value": rays_per_frame / e_sec / 1e6, "unit": "Mrays/s", "h2d_bytes_per_step": int(stats["h2d_bytes"]),
               "d2h_bytes_per_step": int(stats["d2h_bytes"]), "api": "b200rt_trace_primary" if primary_only else "b200rt_render"}
    else:
        pinned = torch.empty((h, w, 4), dtype=torch.float32).pin_memory() if rank == 0 else None
        cam_host = torch.from_numpy(cam.as_array17()).pin_memory()
        cam_dev = torch.empty(17, dtype=torch.float32, device=device)
        barrier()
        t0 = time.perf_counter()
        n_e2e = max(1, min(args.steps, 3))
        for i in range(n_e2e):
            cam_dev.copy_(cam_host, non_blocking=True)          # the frame's input
            img = step()
            if rank == 0:
                pinned.copy_(img, non_blocking=True)
            torch.cuda.synchronize()
        barrier()
        e_sec = torch.tensor([(time.perf_counter() - t0) / n_e2e], dtype=torch.float64, device=device)
        dist.all_reduce(e_sec, op=dist.ReduceOp.MAX)
        e2e = {"value": rays_per_frame / float(e_sec.item()) / 1e6, "unit": "Mrays/s", "h2d_bytes_per_step": 68,
               "d2h_bytes_per_step": w * h * 16, "api": "distributed.FrameGatherer.frame + D2H (rank 0)"}

    # same frame with the rays that provably cannot change the image skipped (B200RT_FLAG_SKIP_DEAD_RAYS): reported next to
    # the headline, which keeps the reference's full ray set
    dead = None
    if not primary_only:
        g2 = D.make_cuda_gatherer(scene, cam, w, h, spp, bounces, rank, world, device, integrator=args.integrator,
                                  flags=args.flags | rt.binding.FLAG_SKIP_DEAD_RAYS)
        g2.frame()
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); g2.frame(); b.record()
        barrier()
        dms = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(dms, op=dist.ReduceOp.MAX)
        dead = {"ms_per_step": float(dms.item()), "spp_per_s": (w * h * spp) / (float(dms.item()) * 1e-3)}
        del g2

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            sw, sh, sspp = cpu_sample_shape(args.workload, args, w, h, spp)
            o, osc = make_oracle_scene(s)
            cam17 = cam.as_array17()
            # ray count of the sample from the GPU's own counter (identical paths => identical query count)
            if primary_only:
                srays = sw * sh
            else:
                _, sst = scene.render(cam, sw, sh, sspp, bounces, flags=args.flags & ~rt.FLAG_SKIP_DEAD_RAYS)
                srays = sst["rays"]
            sec = run_cpu_sample(osc, cam17, args.workload, sw, sh, sspp, bounces)
            cpu_baseline = {"value": srays / sec / 1e6, "unit": "Mrays/s", "cores": host_threads(), "kind": o.kind,
                            "sample": f"{sw}x{sh} px x {sspp} spp of the {w}x{h} x {spp} spp frame ({srays} rays, {sec:.2f} s)",
                            "spp_per_s": sw * sh * sspp / sec}
        except Exception as e:          # the baseline is reported context, never a reason to lose the GPU numbers
            cpu_baseline = {"value": None, "unit": "Mrays/s", "cores": 0, "kind": "port", "sample": f"failed: {e}"}

    # roofline of the dominant kernel (contract): algorithmic bytes per launch / its average launch duration, CUDA events on its
    # launching stream. Path tracing: wf_trace_coop, every launch of one frame timed alone (B200RT_FLAG_TIME_KERNELS: one tile
    # group at a time, full persistent grid); "frame" is the same arithmetic over the whole overlapped frame (trace + shade).
    traffic_total = profile_traffic(args.workload)
    frame_view = {"kernel": kernel_name, "achieved": achieved, "frac": achieved / peak, "kernel_ms": kernel_ms,
                  "traffic": (traffic_total / world) if traffic_total else None}
    if dominant:
        tpr = profile_traffic(args.workload + "_trace_bytes_per_ray")
        roofline = {"bound": "hbm", "kernel": dominant["kernel"], "achieved": dominant["achieved"], "peak": peak, "unit": "GB/s",
                    "frac": dominant["frac"], "traffic": (tpr * dominant["rays_per_launch"]) if tpr else None,
                    "bytes_per_ray": bytes_per_ray, "launches": dominant["launches"], "avg_launch_ms": dominant["avg_launch_ms"],
                    "rays_per_launch": dominant["rays_per_launch"], "share_of_kernel_time": dominant["share_of_kernel_time"],
                    "shade_avg_launch_ms": dominant["shade_avg_launch_ms"], "peak_source": peak_src, "frame": frame_view}
    else:
        roofline = {"bound": "hbm", "kernel": kernel_name, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": frame_view["traffic"], "bytes_per_ray": bytes_per_ray, "rays_per_launch": rays_per_launch,
                    "kernel_ms": kernel_ms, "peak_source": peak_src}

    if rank == 0:
        info = scene.bvh_info()
        line = {
            "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic (procedural stand-ins: the reference ships neither the Dragon OBJ nor the HDR skysphere)"
                    if args.workload != "c1" else "bundled cornell_pbr.obj (parsed by the reference, committed fixture)",
            "spp_per_s": (w * h * spp) / (ms_per_step * 1e-3),
            "rays_per_frame": rays_per_frame, "rays_per_sample": rays_per_frame / (w * h * spp),
            "config": {"workload": desc, "width": w, "height": h, "spp": spp, "max_bounces": bounces, "triangles": n_tri,
                       "integrator": "megakernel" if args.integrator == 0 else "wavefront", "flags": args.flags,
                       "partition": f"interleaved 16x16 tiles over {world} rank(s), scene replicated, NCCL gather to rank 0",
                       "l2": "256 MiB buffer written between timed iterations (L2 flush)",
                       "bvh": {k: info[k] for k in ("n_inner_nodes", "n_leaves", "max_depth", "has_diag_slabs", "n_wide_nodes", "wide_max_depth")},
                       "scene_build_s": build_s, "scene_device_bytes": scene.device_bytes()},
            "skip_dead_rays": dead, "verified_equal_to_1_rank": verified,
            "clocks": sampler.result(),
            "e2e": e2e,
            "gpu_launches": launches_per_step * args.steps,
            "roofline": roofline,
            "cpu_baseline": cpu_baseline,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
