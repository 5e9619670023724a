#!/usr/bin/env python
"""bench.py -- Mpaths/s of the wavefront path-tracing loop (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config 1..5] [--spp S]

Workload (config.workload): BASELINE.json configs[3], the configuration its metric is quoted on --
scenes/cornellSpaceship.txt at 1920x1080, depth 8 -- with the generated STAND-IN MESH (the reference's spaceship
OBJ is missing from its checkout) and the reference's four 4096x4096 textures when the build copied them,
procedural ones otherwise (config.textures says which).  --config selects any of the five BASELINE.json
configurations.

A step is ONE FRAME: one iteration (one sample per pixel through the whole depth loop) on EVERY rank.  Rank r of
a frame that starts at iteration i renders i + r; the frame is combined on the device (csrc/multi.cu:
k_frame_reduce over NVLink peer memory, or one NCCL reduce with --reduce nccl) into the running sum on rank 0,
once per frame.  Both legs go through mygpuraytracer_b200.distributed.FrameRenderer (b2pt_shard_*):
  value   frames stay on the device (no host pointers): N*K*W*H paths / max-over-ranks device time
  e2e     the same frames with the reference's output contract (apps/src/pathtrace.cu:662-668): after EVERY
          frame rank 0 copies the running sum to pinned host memory; d2h bytes counted per step
Each leg times the K-step window `--reps` times back to back (pipeline full at both ends) and reports the
median window; `windows_ms` holds all of them.

Keys beyond the base contract:
  roofline        the kernel with the largest device time in one iteration (CUDA events around every launch of
                  a lane context of the timed region, run alone): algorithmic bytes / average launch time against
                  the measured HBM copy bandwidth (MEASURED_PEAKS.json); `regime` says which grids were timed
  roofline_iter   Bytes_iter = 84*P + 280*S over the step
  roofline_kernels  every kernel: algorithmic bytes over its live launch time against measured HBM, and (default
                  workload) `issue`: its warp instructions (committed ncu pass) over the same time against the issue
                  capacity -- the bound that matters for the intersection kernels; issue_iter is the same for a step
  cpu_baseline    the reference's own intersections.h / interactions.h compiled for the host
                  (oracle/_ref/ref_cpu, kind "reference") or the C oracle (kind "port") on a bounded sample
  cpu_baseline_config1  BASELINE.md's B-CPU line: scenes/cornell.txt 800x800 depth 8, iteration 1, in full
  reference_gpu   the reference's unmodified pathtrace.cu built for sm_100a (oracle/_ref/ref_gpu) on the same
                  workload and GPU: 1 warm-up + 3 timed calls, min / median (the >=10x comparator of north_star)
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of each kernel, from the committed `ncu --set full`
# capture of this workload (cold caches: an upper bound on the warm traffic).  The capture is the DEPTH-1 launch
# (951 052 rays) of a lane context of the timed region; `roofline.traffic_vs_algorithmic` compares it with the
# algorithmic bytes of that same launch, not with the average launch.
TRAFFIC_SOURCE = "profiles/r02_ncu_depth1_shared.md (depth-1 launches, shared-SM grids); fused kernels: profiles/r01_final_ncu_depth01.md"
TRAFFIC = {
    "k_intersect_analytic": 32.02e6, "k_mesh_walk": 30.58e6, "k_mesh_walk_long": 8.67e6, "k_mesh_finish": 49.18e6,
    "k_sort_material": 1.99e6, "k_shade_compact": 170.02e6, "k_shade_trace": 197.83e6, "k_generate_trace": 123.39e6,
}
TRAFFIC_DEPTH = 1
# Warp instructions of ONE iteration per kernel (ncu smsp__inst_executed.sum summed over the 8 depths, lanes per
# instruction beside it) for the DEFAULT workload only (config 4, 250 500 triangles): profiles/r02_inst_per_kernel.md.
# Against them `roofline_kernels[...]["issue"]` states what the HBM fraction cannot say for the kernels that are
# bound by instruction issue and latency: warp instructions per second of the live launch times against the issue
# capacity of the GPU (SMs x 4 schedulers x SM clock).
INST_SOURCE = "profiles/r02_inst_per_kernel.md (end of round 2)"
INST_PER_STEP = {  # kernel: (million warp instructions, active lanes per instruction)
    "k_intersect_analytic": (183.06, 20.7), "k_mesh_walk": (164.00, 14.1), "k_mesh_walk_long": (26.32, 17.8),
    "k_mesh_finish": (20.43, 12.5), "k_sort_material": (17.58, 30.7), "k_shade_compact": (50.60, 30.3), "k_generate": (9.06, 32.0),
}

METRIC = "Mpaths/s"
# BASELINE.json "configs", in order.  4 is the configuration the metric is quoted on (the default).
CONFIGS = {
    1: dict(scene="cornell", width=800, height=800, depth=8, options={}, spp=1,
            what="configs[0]: Cornell box, analytic spheres / cubes, diffuse; the reference's CPU-runnable case"),
    2: dict(scene="cornellGlass", width=800, height=800, depth=8, options=dict(depth_of_field=1, antialiasing=1), spp=5000,
            what="configs[1]: refractive + reflective BSDFs with DOF and stochastic AA, 5000 spp on 1 B200"),
    3: dict(scene="cornellObj", width=1920, height=1080, depth=8, options={}, spp=0,
            what="configs[2]: triangle mesh with diffuse / specular / emission textures, LBVH build + traversal"),
    4: dict(scene="cornellSpaceship", width=1920, height=1080, depth=8, options={}, spp=0,
            what="configs[3]: heavy textured / bump-mapped mesh, material sort + compaction, at 1/2/4/8 B200"),
    5: dict(scene="cornellSpaceship", width=3840, height=2160, depth=12, options={}, spp=4096,
            what="configs[4]: synthetic scaling, 4096 spp, spp-sharded at 2/4/8 B200"),
}
CPU_SAMPLE = (48, 27)  # resolution of the bounded CPU sample (same scene, mesh, depth): ~10 s per iteration on 8 cores


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# The contract is ONE JSON line on stdout.  Libraries (NCCL prints its version
# banner) write to file descriptor 1 too, so everything but the result line is
# sent to stderr: fd 1 is pointed at fd 2 and the line goes to a saved copy.
_RESULT_FD = None


def capture_stdout():
    global _RESULT_FD
    if _RESULT_FD is None:
        sys.stdout.flush()
        _RESULT_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_RESULT_FD, data)


# ---------------------------------------------------------------------------------------
# workload
# ---------------------------------------------------------------------------------------
def prepare_assets(triangles: int, write: bool = True):
    """The run tree (scene files, stand-in mesh, textures).  Only ONE process may create / repoint it
    (write=True, rank 0); the other ranks look at the finished tree after a barrier."""
    from mygpuraytracer_b200 import assets

    if write:
        ref_tex = os.path.join(ROOT, "oracle", "_ref", "run", "textures")
        root = assets.prepare(reference_textures=ref_tex if os.path.isdir(ref_tex) else None)
        assets.set_mesh(root, triangles)
    else:
        root = assets.DEFAULT_ROOT
    return root, ("reference JPEGs" if assets.textures_are_reference(root) else "procedural 4096x4096 (reference JPEGs absent)")


def workload_config(args, n_tris: int, textures: str, world: int):
    from mygpuraytracer_b200 import scenes

    mesh = scenes.uses_mesh(args.scene)
    opts = CONFIGS[args.config]["options"]
    return {
        "workload": f"BASELINE.json configs[{args.config - 1}]: scenes/{args.scene}.txt {args.width}x{args.height} depth {args.depth}"
                    + (f", STAND-IN MESH {n_tris} triangles (reference OBJ missing)" if mesh else ", analytic geoms only")
                    + f", AA on, DOF {'on' if opts.get('depth_of_field') else 'off'}, material sort on",
        "baseline_config": args.config, "scene": args.scene, "width": args.width, "height": args.height, "depth": args.depth,
        "triangles": n_tris if mesh else 0, "textures": textures if mesh else "none",
        "step": "one frame = one iteration (1 spp) on every rank",
        "sharding": "samples-per-pixel: rank r of a frame starting at iteration i renders i + r; the frame is combined into "
                    "the running sum on rank 0 once per frame (" + ("k_frame_reduce over NVLink peer memory, in iteration order"
                                                                       if args.reduce == "p2p" else "one NCCL reduce") + ")",
        "reduce": args.reduce if world > 1 else "none (one rank)",
        "lanes_per_gpu": args.streams,
        "l2": "working set per step (path state 2x48 B + hits 32 B per path, 4 textures, BVH) > 126 MB L2; no flush",
        "rng": "slot-keyed minstd (reference mode)", "trig": "native",
    }


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region: NVML polled every few ms from a thread (the
    timed region is a fraction of a second); falls back to `nvidia-smi -lms` when pynvml is missing."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.rows = []       # nvidia-smi fallback rows
        self.sm = []         # MHz samples
        self.max_mhz = None
        self.reason_bits = 0
        self.proc = None
        self.thread = None
        self.stop_flag = threading.Event()
        self.nvml = None

    def start(self):
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _poll(self):
        n = self.nvml
        while not self.stop_flag.is_set():
            try:
                self.sm.append(float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)))
                self.reason_bits |= int(n.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
            except Exception:
                try:
                    self.reason_bits |= int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
                except Exception:
                    pass
            time.sleep(0.004)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.nvml is not None:
            self.stop_flag.set()
            self.thread.join(timeout=1)
            n = self.nvml
            names = {"hw_slowdown": getattr(n, "nvmlClocksEventReasonHwSlowdown", 0x8),
                     "hw_thermal_slowdown": getattr(n, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                     "sw_thermal_slowdown": getattr(n, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                     "sw_power_cap": getattr(n, "nvmlClocksEventReasonSwPowerCap", 0x4)}
            reasons = sorted(k for k, bit in names.items() if self.reason_bits & int(bit))
            return {"sm_mhz": statistics.median(self.sm) if self.sm else None, "sm_max_mhz": self.max_mhz,
                    "reasons": reasons, "samples": len(self.sm), "source": "nvml"}
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) > 8 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) > 8 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) > 8:
                for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi"}


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"



# ---------------------------------------------------------------------------------------
# CPU legs
# ---------------------------------------------------------------------------------------
def host_threads() -> int:
    """The host cores this process may use, whatever OMP_NUM_THREADS the launcher exported
    (torch.distributed.run sets it to 1)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cpu_sample(root: str, args, steps: int = 1, res=CPU_SAMPLE, scene=None, depth=None):
    """Time the reference's CPU code on a bounded sample: the same scene file,
    mesh, textures and depth at a reduced resolution, `steps` iterations."""
    from mygpuraytracer_b200 import assets

    w, h = res
    scene = scene or args.scene
    depth = depth or args.depth
    threads = host_threads()
    scene_txt = assets.scene_file(scene, w, h, depth=depth, root=root)
    sample = f"{scene} {w}x{h} depth {depth}, same mesh/textures, {steps} iteration(s) of {w * h} paths"
    try:
        from oracle import harness

        if harness.have("ref_cpu"):
            # run from the asset tree's bin/ so '../models/...' resolves to the same files
            cmd = [os.path.join(harness.REF_DIR, "ref_cpu"), "--scene", scene_txt, "--iters", str(steps), "--threads", str(threads)]
            env = dict(os.environ, OMP_NUM_THREADS=str(threads))
            t0 = time.time()
            p = subprocess.run(cmd, cwd=os.path.join(root, "bin"), capture_output=True, text=True, timeout=1800, env=env)
            wall = time.time() - t0
            for line in p.stdout.splitlines():
                if line.startswith("REF_CPU_RESULT"):
                    r = json.loads(line.split(" ", 1)[1])
                    return {"value": r["mpaths_per_s"], "unit": METRIC, "cores": r["threads"], "kind": "reference",
                            "sample": sample, "sample_width": w, "sample_height": h,
                            "ms_per_iteration": r["ms_per_iter"], "wall_s": round(wall, 2)}
            log("ref_cpu produced no result line:", p.stderr[-500:])
    except Exception as e:  # fall through to the port
        log("ref_cpu unavailable:", e)
    from mygpuraytracer_b200 import abi, api
    from oracle import oracle

    oracle.set_num_threads(threads)
    pod = api.Scene(scene_txt).pod
    t0 = time.time()
    oracle.render(pod, abi.default_options(**CONFIGS[args.config]["options"]), 1, steps, 1)
    dt = time.time() - t0
    return {"value": w * h * steps / dt / 1e6, "unit": METRIC, "cores": oracle.num_threads(), "kind": "port",
            "sample": sample, "sample_width": w, "sample_height": h, "ms_per_iteration": 1e3 * dt / steps, "wall_s": round(dt, 2)}


def reference_gpu(root: str, args):
    """The unmodified reference pathtrace.cu (sm_100a) on the full workload: warm-up + timed calls."""
    exe = os.path.join(ROOT, "oracle", "_ref", "ref_gpu")
    if not os.access(exe, os.X_OK) or args.ref_gpu_iters <= 0:
        return None
    from mygpuraytracer_b200 import assets

    scene_txt = assets.scene_file(args.scene, args.width, args.height, depth=args.depth, root=root)
    try:
        p = subprocess.run([exe, "--scene", scene_txt, "--time", "--iters", str(args.ref_gpu_iters), "--warmup", "1"],
                           cwd=os.path.join(root, "bin"), capture_output=True, text=True, timeout=args.ref_gpu_timeout)
        for line in p.stdout.splitlines():
            if line.startswith("REF_GPU_RESULT"):
                r = json.loads(line.split(" ", 1)[1])
                med = r.get("call_ms_median", r["call_ms_per_iter"])
                return {"what": "reference apps/src/pathtrace.cu, unmodified, nvcc -O3 sm_100a, same GPU and workload",
                        "warmup": r.get("warmup", 0), "iterations": r["iters"], "ms_per_iteration_loop": r["loop_ms_per_iter"],
                        "ms_per_iteration_call": r["call_ms_per_iter"], "ms_per_call_min": r.get("call_ms_min"),
                        "ms_per_call_median": med, "ms_per_call_max": r.get("call_ms_max"),
                        "mpaths_per_s": args.width * args.height / (med * 1e-3) / 1e6}
        return {"error": (p.stderr or p.stdout)[-300:]}
    except subprocess.TimeoutExpired:
        return {"error": f"timed out after {args.ref_gpu_timeout}s"}


def reference_host_adapter(root: str, args, iters: int):
    """The reference's own host code (Scene, scene.cpp, the pathtrace() call loop of main.cpp) linked with
    integration/pathtrace_b2pt.cpp + libb2pt.so instead of apps/src/pathtrace.cu (oracle/_ref/ref_adapter)."""
    exe = os.path.join(ROOT, "oracle", "_ref", "ref_adapter")
    if not os.access(exe, os.X_OK):
        return None
    from mygpuraytracer_b200 import assets

    scene_txt = assets.scene_file(args.scene, args.width, args.height, depth=args.depth, root=root)
    try:
        p = subprocess.run([exe, "--scene", scene_txt, "--iters", str(iters)], cwd=os.path.join(root, "bin"),
                           capture_output=True, text=True, timeout=300)
        for line in p.stdout.splitlines():
            if line.startswith("REF_ADAPTER_RESULT"):
                r = json.loads(line.split(" ", 1)[1])
                return {"what": "reference host code (scene.cpp loader + the pathtrace() loop of main.cpp:245-264) calling "
                                "integration/pathtrace_b2pt.cpp -> libb2pt.so; state.image / state.albedo on the host after every call",
                        "iterations": r["iters"], "ms_per_call": r["call_ms_per_iter"],
                        "timer_ms_per_call": r["timer_ms_per_iter"], "mpaths_per_s": r["mpaths_per_s_call"]}
        return {"error": (p.stderr or p.stdout)[-300:]}
    except subprocess.TimeoutExpired:
        return {"error": "timed out"}


# ---------------------------------------------------------------------------------------
# arms
# ---------------------------------------------------------------------------------------
def run_reference(args, rank: int, world: int):
    """The reference's CPU implementation of the path on this box's host cores (rank 0 only).  A step is one
    iteration of a BOUNDED SAMPLE of the workload (same scene, mesh, textures, depth; fewer pixels): the
    brute-force reference costs O(paths x triangles).  config.workload names the sample that really ran."""
    if rank != 0:
        return
    root, textures = prepare_assets(args.triangles)
    from mygpuraytracer_b200 import scenes, standin_mesh

    n_tris = len(standin_mesh.build(args.triangles)[3]) if scenes.uses_mesh(args.scene) else 0
    # Size the sample so that steps+warmup iterations fit in ~150 s.
    probe = cpu_sample(root, args, 1, (16, 9))
    per_pixel_s = probe["ms_per_iteration"] * 1e-3 / (16 * 9)
    budget_s = 150.0 / max(1, args.steps + min(args.warmup, 1))
    full = args.width * args.height
    pixels = max(16 * 9, min(full, int(budget_s / per_pixel_s)))
    if pixels >= full:
        w, h = args.width, args.height
    else:
        h = max(9, int((pixels * args.height / args.width) ** 0.5))
        w = max(16, pixels // h)
    r = cpu_sample(root, args, max(1, args.steps + min(args.warmup, 1)), (w, h))
    cfg = workload_config(args, n_tris, textures, world)
    same = (w, h) == (args.width, args.height)
    cfg["workload"] = (f"BOUNDED SAMPLE {w}x{h} of: " if not same else "") + cfg["workload"]
    cfg["sample"] = {"width": w, "height": h, "paths_per_step": w * h, "full_width": args.width, "full_height": args.height,
                     "same_config": same}
    cfg["sharding"] = "none: host cores of rank 0"
    cfg["reduce"] = "none"
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": METRIC, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": r["ms_per_iteration"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
        "cpu_baseline": {"value": r["value"], "unit": METRIC, "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]},
        "e2e": {"value": r["value"], "unit": METRIC, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": f"each step is one iteration of the bounded sample ({w}x{h}, {w * h} paths, {r['cores']} host threads); "
                "Mpaths/s of the brute-force reference falls with the triangle count, not with the pixel count",
    }
    emit(line)


def run_ours(args, rank: int, world: int, local_rank: int):
    import numpy as np
    import torch

    from mygpuraytracer_b200 import abi, api, assets, distributed, scenes

    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    scene_txt = None
    if rank == 0:
        root, textures = prepare_assets(args.triangles)
        scene_txt = assets.scene_file(args.scene, args.width, args.height, depth=args.depth, root=root)
    if dist:
        dist.barrier()
    if rank != 0:
        root, textures = prepare_assets(args.triangles, write=False)
        scene_txt = os.path.join(root, "scenes", f"{args.scene}_{args.width}x{args.height}_d{args.depth}.txt")
    t0 = time.time()
    scene = api.Scene(scene_txt)
    load_s = time.time() - t0
    pod = scene.pod
    has_mesh = scenes.uses_mesh(args.scene)
    n_tris = len(pod.face_pos)
    P = pod.n_pixels
    cfg_opts = CONFIGS[args.config]["options"]
    # rank 0's host buffers are two pinned arrays that nobody else writes: the albedo AOV, which only changes on
    # iteration 1, is copied when it changed (declared in e2e.albedo); the every-call form is measured beside it
    opt = abi.default_options(device=local_rank, persistent_host_albedo=1, **cfg_opts)
    KC = max(1, args.streams)
    fr = distributed.FrameRenderer(scene, opt, lanes=KC, reduce=args.reduce)
    shard = fr.shard
    stream = torch.cuda.ExternalStream(shard.stream_ptr(), device=torch.device("cuda", local_rank))
    lane0 = shard.lane(0)
    mesh_geom = int(np.nonzero(pod.geoms["type"] == abi.OBJ)[0][0]) if has_mesh else -1
    bvh = lane0.bvh_info(mesh_geom) if has_mesh else None

    W, K, R = args.warmup, args.steps, max(1, args.reps)
    host_img = torch.empty(P * 3, dtype=torch.float32).pin_memory()
    host_alb = torch.empty(P * 3, dtype=torch.float32).pin_memory()
    img_np, alb_np = host_img.numpy().reshape(P, 3), host_alb.numpy().reshape(P, 3)
    frame = [0]  # frames rendered so far: the sequence never restarts, so no call is mispredicted

    def barrier():
        if dist:
            dist.barrier()
        torch.cuda.synchronize()

    def frames(n, to_host):
        for _ in range(n):
            if to_host:
                fr.pathtrace(1 + frame[0] * world, img_np, alb_np)
            else:
                fr.pathtrace(1 + frame[0] * world, None, None)
            frame[0] += 1

    def timed_leg(to_host):
        """W warm-up frames, then R back-to-back windows of exactly K frames; device time on the shard's stream,
        max over ranks per window.  The lanes hold the next frames when a window starts and when it ends."""
        frames(max(W, KC), to_host)
        barrier()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(R + 1)]
        launches0 = shard.launch_count()
        ev[0].record(stream)
        for i in range(R):
            frames(K, to_host)
            ev[i + 1].record(stream)
        barrier()
        launches = shard.launch_count() - launches0
        ms = torch.tensor([ev[i].elapsed_time(ev[i + 1]) for i in range(R)], device="cuda")
        cnt = torch.tensor([launches], device="cuda", dtype=torch.int64)
        if dist:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
        return [float(x) for x in ms.tolist()], int(cnt.item())

    if args.spp > 0:
        converge(args, fr, shard, stream, dist, rank, world, P, img_np, alb_np, n_tris, textures, barrier)
        fr.close()
        if dist:
            dist.barrier()
            dist.destroy_process_group()
        return

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    dev_windows, dev_launches = timed_leg(False)
    e2e_windows, e2e_launches = timed_leg(True)
    clocks = sampler.stop() if rank == 0 else None
    misses = shard.misses()
    dev_ms, e2e_ms = statistics.median(dev_windows), statistics.median(e2e_windows)
    value = world * K * P / (dev_ms * 1e-3) / 1e6
    e2e_value = world * K * P / (e2e_ms * 1e-3) / 1e6
    checksum = float(np.float64(img_np.sum())) if rank == 0 else 0.0
    live = lane0.live_counts()
    segments = int(live[: args.depth].sum())

    extras = {}
    if rank == 0 and not args.no_extras:
        # ---- per-kernel times of one iteration: a lane context of the timed region (shared-SM grids, unfused
        # kernels), run alone, CUDA events around every launch.  ncu serialises launches, so its launch list of
        # this command times the same kernels in the same regime (profiles/).
        shard.sync()
        for k in range(KC):
            shard.lane(k).sync()  # the lanes hold speculated frames: wait for them, then lane 0 renders alone
        base_iter = 1 + (frame[0] + 4 * KC) * world
        prof = [lane0.profile_kernels(base_iter + i) for i in range(5)]
        prof = {k: statistics.median(p[k] for p in prof) for k in prof[0]}
        walks = lane0.walk_counts() if has_mesh else np.zeros(args.depth, np.int32)
        long_walks = lane0.walk_counts(long_walks=True) if has_mesh else np.zeros(args.depth, np.int32)
        extras.update(prof=prof, walks=walks, long_walks=long_walks)
    barrier()
    fr.close()

    if rank == 0:
        if not args.no_extras and world == 1:
            extras.update(single_process_legs(args, scene, opt, P, K, W, KC, img_np, alb_np))
        peak, peak_src = hbm_peak()
        line = {
            "metric": METRIC, "value": value, "unit": METRIC, "n_gpus": world, "steps": K, "warmup": max(W, KC),
            "ms_per_step": dev_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(args, n_tris, textures, world),
            "windows_ms": {"device": dev_windows, "e2e": e2e_windows, "reps": R,
                           "note": "each window is exactly `steps` frames, timed back to back; value / e2e use the median"},
            "e2e": {"value": e2e_value, "unit": METRIC, "h2d_bytes_per_step": 8 * world, "d2h_bytes_per_step": P * 12,
                    "ms_per_step": e2e_ms / K,
                    "call": "distributed.FrameRenderer.pathtrace(first_iter, host_image, host_albedo) = b2pt_shard_frame_begin / "
                            "_reduce / _end: N reference pathtrace() calls (apps/src/pathtrace.h:9) per step, one per rank; "
                            "rank 0 has the running sum in pinned host memory after EVERY step",
                    "albedo": "persistent_host_albedo=1: the albedo AOV (P*12 bytes more) is copied when it changed (iteration 1), "
                              "not every step; e2e_albedo_every_call is the reference's literal 2 x P*12 per call",
                    "lanes": KC, "mispredicted_calls": int(misses)},
            "gpu_launches": int(e2e_launches), "gpu_launches_device_leg": int(dev_launches), "lanes_per_gpu": KC,
            "clocks": clocks,
            "segments_per_step": segments * world, "segments_per_iteration": segments,
            "live_paths_per_depth": [int(x) for x in live[: args.depth + 1]],
            "scene_load_s": round(load_s, 2), "image_checksum": checksum,
        }
        if args.config == 4 and args.triangles == 250_000 and clocks and clocks.get("sm_mhz"):
            # the whole step against the issue capacity: what four lanes together reach
            tot = sum(v[0] for v in INST_PER_STEP.values()) * 1e6
            cap = torch.cuda.get_device_properties(local_rank).multi_processor_count * 4 * clocks["sm_mhz"] * 1e6
            line["issue_iter"] = {"warp_inst_per_iteration": tot, "achieved": tot / (dev_ms / K * 1e-3) / 1e9, "peak": cap / 1e9,
                                  "unit": "G warp inst/s", "frac": tot / (dev_ms / K * 1e-3) / cap, "source": INST_SOURCE,
                                  "note": "per GPU; instruction counts of one iteration from the committed ncu pass, time from this run"}
        iter_bytes = 84.0 * P + 280.0 * segments
        iter_gbs = iter_bytes * world / (dev_ms / K * 1e-3) / 1e9
        line["roofline_iter"] = {"bytes_per_step": iter_bytes * world, "achieved": iter_gbs, "peak": peak * world, "unit": "GB/s",
                                 "frac": iter_gbs / (peak * world), "formula": "N * (84*P + 280*S)"}
        if bvh is not None:
            line["bvh"] = {"triangles": int(bvh.n_faces), "nodes": int(bvh.n_nodes), "max_depth": int(bvh.max_depth),
                           "build_ms": float(bvh.build_ms)}
        if "prof" in extras:
            prof = extras["prof"]
            n_walks = int(extras["walks"][: args.depth].sum())
            n_long = int(extras["long_walks"][: args.depth].sum())
            # algorithmic bytes per launch unit (kernel table)
            kern = {
                "k_intersect_analytic": ("analytic", 58.0 * segments),   # ray 24 B, hit record 32 B, key 1 B, flag 1 B
                "k_mesh_walk": ("walk", 57.0 * n_walks),                 # queue 4, ray 24, analytic hit 8, (t, bary, ids) 20 + key 1
                "k_mesh_walk_long": ("walk_long", 57.0 * n_long),
                "k_mesh_finish": ("finish", 57.0 * n_walks),             # queue 4 + partial record 20 read, record 32 + flag 1 written
                "k_sort_material": ("sort", 10.0 * segments),            # key 1 + flag 1 read, permutation 4 + rank 4 written
                                                                         # (k_sort_material_few when the scene has <= 8 materials)
                "k_shade_compact": ("shade", 132.0 * segments),          # permutation 4 + hit 32 + state 48 read, state 48 written
                "k_generate": ("generate", 44.0 * P),
            }
            roofline_kernels = {}
            for name, (key, nbytes) in kern.items():
                kms = prof.get(key, 0.0)
                gbs = nbytes / (kms * 1e-3) / 1e9 if kms > 0 else 0.0
                roofline_kernels[name] = {"ms_per_step": kms, "algorithmic_bytes_per_step": nbytes, "achieved": gbs,
                                          "unit": "GB/s", "frac": gbs / peak}
                if args.config == 4 and args.triangles == 250_000 and name in INST_PER_STEP and kms > 0 and clocks and clocks.get("sm_mhz"):
                    minst, lanes = INST_PER_STEP[name]
                    cap = torch.cuda.get_device_properties(local_rank).multi_processor_count * 4 * clocks["sm_mhz"] * 1e6
                    roofline_kernels[name]["issue"] = {
                        "warp_inst_per_step": minst * 1e6, "lanes_per_inst": lanes, "achieved": minst * 1e6 / (kms * 1e-3) / 1e9,
                        "peak": cap / 1e9, "unit": "G warp inst/s", "frac": minst * 1e6 / (kms * 1e-3) / cap, "source": INST_SOURCE}
            dom = max((k for k in kern if k != "k_generate"), key=lambda k: prof.get(kern[k][0], 0.0))
            dom_ms, dom_bytes = prof[kern[dom][0]], kern[dom][1]
            dom_gbs = dom_bytes / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
            # algorithmic bytes of the launch the ncu capture holds (depth TRAFFIC_DEPTH), for traffic_vs_algorithmic
            per_unit = dom_bytes / max(1, {"k_mesh_walk": n_walks, "k_mesh_finish": n_walks, "k_mesh_walk_long": n_long}.get(dom, segments))
            units_d = {"k_mesh_walk": int(extras["walks"][TRAFFIC_DEPTH]), "k_mesh_finish": int(extras["walks"][TRAFFIC_DEPTH]),
                       "k_mesh_walk_long": int(extras["long_walks"][TRAFFIC_DEPTH])}.get(dom, int(live[TRAFFIC_DEPTH]))
            line["roofline"] = {
                "kernel": dom, "bound": "hbm", "achieved": dom_gbs, "peak": peak, "unit": "GB/s", "frac": dom_gbs / peak,
                "traffic": TRAFFIC.get(dom), "traffic_source": TRAFFIC_SOURCE, "peak_source": peak_src,
                "traffic_launch": f"depth {TRAFFIC_DEPTH}", "algorithmic_bytes_traffic_launch": per_unit * units_d,
                "traffic_vs_algorithmic": (TRAFFIC.get(dom, 0.0) / (per_unit * units_d)) if units_d else None,
                "regime": "a lane context of the timed region (grids sized to a share of the SMs, unfused kernels) rendering one "
                          "iteration ALONE, CUDA events around every launch: the regime ncu's serialised launch list of this "
                          "command measures",
                "algorithmic_bytes_per_launch": dom_bytes / args.depth, "launches_per_step": args.depth,
                "ms_per_launch": dom_ms / args.depth, "walks_per_step": n_walks, "long_walks_per_step": n_long,
                "note": "achieved = algorithmic bytes per launch / average launch duration.  The BVH walk is issue / latency "
                        "bound, not HBM bound: see profiles/ for issue-slot, FP32-pipe and divergence counters"}
            if "issue" in roofline_kernels.get(dom, {}):
                # the bound that does apply to this kernel, next to the HBM figure the contract asks for
                line["roofline"]["issue"] = roofline_kernels[dom]["issue"]
            line["roofline_kernels"] = roofline_kernels
            line["kernel_ms_per_step"] = prof
        for k in ("kernel_ms_per_step_full_width", "e2e_single_context", "e2e_albedo_every_call", "e2e_single_process_multi"):
            if k in extras:
                line[k] = extras[k]
        if world == 1 and not args.no_cpu:
            ra = reference_host_adapter(root, args, max(20, min(K, 200)))
            if ra is not None:
                line["e2e_reference_host"] = ra
            line["cpu_baseline"] = cpu_sample(root, args, 1)
            if args.config != 1:
                # BASELINE.md B-CPU: the reference's CPU-runnable case, in full
                line["cpu_baseline_config1"] = cpu_sample(root, args, 1, (800, 800), scene="cornell", depth=8)
            rg = reference_gpu(root, args)
            if rg is not None:
                line["reference_gpu"] = rg
        emit(line)
    if dist:
        dist.barrier()
        dist.destroy_process_group()


def single_process_legs(args, scene, opt, P, K, W, KC, img_np, alb_np):
    """N = 1 only, after the shard is gone: the full-width single context (kernel split and the unoverlapped
    b2pt_pathtrace call) and the pipelined call with the albedo AOV copied on every call."""
    import torch

    from mygpuraytracer_b200 import abi, api

    out = {}
    opt1 = abi.default_options(device=opt.device, persistent_host_albedo=1, **CONFIGS[args.config]["options"])
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stream = torch.cuda.Stream()
    with api.Renderer(scene, opt1) as r:
        r.set_stream_ptr(stream.cuda_stream)
        for i in range(3):
            r.pathtrace(1 + i, img_np, alb_np)
        torch.cuda.synchronize()
        e0.record(stream)
        for i in range(K):
            r.pathtrace(4 + i, img_np, alb_np)
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        out["e2e_single_context"] = {"value": K * P / (ms * 1e-3) / 1e6, "unit": METRIC, "ms_per_step": ms / K,
                                     "call": "b2pt_pathtrace(ctx, iter, host_image, host_albedo): render, then copy, nothing overlapped"}
        prof = [r.profile_kernels(10 + K + i) for i in range(5)]
        out["kernel_ms_per_step_full_width"] = {k: statistics.median(p[k] for p in prof) for k in prof[0]}
    opt2 = abi.default_options(device=opt.device, persistent_host_albedo=0, **CONFIGS[args.config]["options"])
    with api.Pipeline(scene, opt2, lanes=KC) as pipe:
        for i in range(max(W, KC)):
            pipe.pathtrace(1 + i, img_np, alb_np)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(K):
            pipe.pathtrace(1 + max(W, KC) + i, img_np, alb_np)
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) * 1e3
        out["e2e_albedo_every_call"] = {"value": K * P / (ms * 1e-3) / 1e6, "unit": METRIC, "ms_per_step": ms / K,
                                        "d2h_bytes_per_step": P * 24,
                                        "call": "b2pt_pipe_pathtrace with persistent_host_albedo=0: image AND albedo to the host "
                                                "on every call, the literal apps/src/pathtrace.cu:662-668 (host wall clock)"}
    return out


def converge(args, fr, shard, stream, dist, rank, world, P, img_np, alb_np, n_tris, textures, barrier):
    """--spp S: render iterations 1 .. S (S rounded up to whole frames) through the e2e path from a zeroed accumulator,
    rank 0 reading the running sum after every frame, and optionally save the final image."""
    import numpy as np
    import torch

    n_frames = (args.spp + world - 1) // world
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record(stream)
    for f in range(n_frames):
        fr.pathtrace(1 + f * world, img_np, alb_np)
    e1.record(stream)
    barrier()
    wall = time.perf_counter() - t0
    ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    if dist:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    spp = n_frames * world
    if rank == 0:
        if args.save_image:
            np.save(args.save_image, img_np / np.float32(spp))
        emit({"metric": METRIC, "mode": "converge", "value": spp * P / (ms * 1e-3) / 1e6, "unit": METRIC, "n_gpus": world,
              "spp": spp, "frames": n_frames, "ms_per_frame": ms / n_frames, "ms_per_iteration": ms / spp, "seconds": ms * 1e-3,
              "wall_s": round(wall, 3), "higher_is_better": True, "scaling": "strong", "dtype": "f32", "data": "synthetic",
              "config": workload_config(args, n_tris, textures, world),
              "e2e": {"value": spp * P / (ms * 1e-3) / 1e6, "unit": METRIC, "h2d_bytes_per_step": 8 * world, "d2h_bytes_per_step": P * 12},
              "image_checksum": float(np.float64(img_np.sum())), "image_file": args.save_image or None,
              "note": "the whole render from a zeroed accumulator, rank 0 has the running sum on the host after every frame"})


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--reps", type=int, default=5, help="back-to-back timed windows of `steps` frames; the median is reported")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=4, choices=sorted(CONFIGS), help="BASELINE.json configuration (1-based)")
    ap.add_argument("--width", type=int, default=0)
    ap.add_argument("--height", type=int, default=0)
    ap.add_argument("--depth", type=int, default=0)
    ap.add_argument("--triangles", type=int, default=250_000)
    ap.add_argument("--streams", type=int, default=6,
                    help="lanes (contexts rendering ahead) per GPU; 4 / 5 / 6 lanes: 2559 / 2604 / 2610 Mpaths/s on one B200")
    ap.add_argument("--reduce", default="p2p", choices=["p2p", "nccl"],
                    help="how a frame is combined across ranks: k_frame_reduce over NVLink peer memory, or one NCCL reduce")
    ap.add_argument("--spp", type=int, default=0, help="render this many samples per pixel from scratch instead of timing windows")
    ap.add_argument("--save-image", default="", help="with --spp: write image / spp (float32 .npy) here")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline / reference_gpu legs")
    ap.add_argument("--no-extras", action="store_true", help="skip the per-kernel profile and the single-process legs")
    ap.add_argument("--ref-gpu-iters", type=int, default=3)
    ap.add_argument("--ref-gpu-timeout", type=int, default=300)
    args = ap.parse_args()
    cfg = CONFIGS[args.config]
    args.scene = cfg["scene"]
    args.width = args.width or cfg["width"]
    args.height = args.height or cfg["height"]
    args.depth = args.depth or cfg["depth"]
    capture_stdout()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
