#!/usr/bin/env python
"""bench.py -- Mpaths/s of the wavefront path-tracing loop (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (config.workload): scenes/cornellSpaceship.txt at 1920x1080, depth 8
-- the configuration BASELINE.json's metric is quoted on -- with the generated
STAND-IN MESH (the reference's spaceship OBJ is missing from its checkout) and
the reference's four 4096x4096 textures when the build copied them, procedural
ones otherwise (config.textures says which).

A step is ONE iteration (one sample per pixel through the whole depth loop) on
every rank.  Ranks shard samples per pixel: rank r of N renders iteration
indices r+1, r+1+N, ... into its own accumulator; after the K timed steps the
accumulators are combined with one NCCL reduce (inside the timed region).
`value` = N*K*W*H paths / max-over-ranks device time, scaling "weak".

Keys beyond the base contract:
  e2e           the same metric through the reference-facing call
                b2pt_pathtrace(): every step renders one iteration and copies
                the running sum and the albedo AOV to pinned HOST buffers, as
                the reference's pathtrace() does (apps/src/pathtrace.cu:663-668)
  roofline      the dominant kernel (k_intersect): algorithmic bytes (56 B per
                path segment: 24 B ray read + 32 B hit record written) over its
                device time measured with CUDA events in this run, against the
                measured HBM copy bandwidth (MEASURED_PEAKS.json)
  roofline_iter Bytes_iter = 84*P + 280*S over the step
  cpu_baseline  the reference's own intersections.h / interactions.h compiled
                for the host (oracle/_ref/ref_cpu, kind "reference") or the C
                oracle (kind "port") on a bounded sample of the same scene
  reference_gpu the reference's unmodified pathtrace.cu built for sm_100a
                (oracle/_ref/ref_gpu) on the same workload, same GPU, when the
                binary travelled to the box (the >=10x comparator of north_star)
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

# dram__bytes_read.sum + dram__bytes_write.sum per launch of each kernel, from the committed `ncu --set full`
# capture of this workload (profiles/, cold caches: an upper bound on the warm traffic)
TRAFFIC = {  # bytes per launch, depth-1 launches of profiles/r01_final_ncu_depth01.md and r01_final_ncu_unfused_depth1.md
    "k_intersect_analytic": 31.24e6, "k_mesh_walk": 71.61e6, "k_mesh_walk_long": 20.69e6, "k_mesh_finish": 50.24e6,
    "k_sort_material": 1.99e6, "k_shade_compact": 169.64e6, "k_shade_trace": 197.83e6, "k_generate_trace": 123.39e6,
}

METRIC = "Mpaths/s"
SCENE = "cornellSpaceship"
WIDTH, HEIGHT, DEPTH = 1920, 1080, 8
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


def workload_config(args, n_tris: int, textures: str):
    return {
        "workload": f"scenes/{SCENE}.txt {args.width}x{args.height} depth {args.depth}, STAND-IN MESH {n_tris} triangles "
                    f"(reference OBJ missing), AA on, DOF off, material sort on",
        "scene": SCENE, "width": args.width, "height": args.height, "depth": args.depth,
        "triangles": n_tris, "textures": textures,
        "step": "one iteration (1 spp) per rank", "sharding": "samples-per-pixel, one NCCL reduce per frame",
        "streams_per_gpu": args.streams,
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
def cpu_sample(root: str, args, steps: int = 1, res=CPU_SAMPLE):
    """Time the reference's CPU code on a bounded sample: the same scene file,
    mesh, textures and depth at a reduced resolution, `steps` iterations."""
    from mygpuraytracer_b200 import assets

    w, h = res
    scene_txt = assets.scene_file(SCENE, w, h, depth=args.depth, root=root)
    sample = f"{SCENE} {w}x{h} depth {args.depth}, same mesh/textures, {steps} iteration(s) of {w * h} paths"
    try:
        from oracle import harness

        if harness.have("ref_cpu"):
            # run from the asset tree's bin/ so '../models/...' resolves to the same files
            cmd = [os.path.join(harness.REF_DIR, "ref_cpu"), "--scene", scene_txt, "--iters", str(steps)]
            t0 = time.time()
            p = subprocess.run(cmd, cwd=os.path.join(root, "bin"), capture_output=True, text=True, timeout=1800)
            wall = time.time() - t0
            for line in p.stdout.splitlines():
                if line.startswith("REF_CPU_RESULT"):
                    r = json.loads(line.split(" ", 1)[1])
                    return {"value": r["mpaths_per_s"], "unit": METRIC, "cores": r["threads"], "kind": "reference",
                            "sample": sample, "ms_per_iteration": r["ms_per_iter"], "wall_s": round(wall, 2)}
            log("ref_cpu produced no result line:", p.stderr[-500:])
    except Exception as e:  # fall through to the port
        log("ref_cpu unavailable:", e)
    from mygpuraytracer_b200 import abi, api
    from oracle import oracle

    pod = api.Scene(scene_txt).pod
    t0 = time.time()
    oracle.render(pod, abi.default_options(), 1, steps, 1)
    dt = time.time() - t0
    return {"value": w * h * steps / dt / 1e6, "unit": METRIC, "cores": oracle.num_threads(), "kind": "port",
            "sample": sample, "ms_per_iteration": 1e3 * dt / steps, "wall_s": round(dt, 2)}


def reference_gpu(root: str, args):
    """The unmodified reference pathtrace.cu (sm_100a) on the full workload."""
    exe = os.path.join(ROOT, "oracle", "_ref", "ref_gpu")
    if not os.access(exe, os.X_OK) or args.ref_gpu_iters <= 0:
        return None
    from mygpuraytracer_b200 import assets

    scene_txt = assets.scene_file(SCENE, args.width, args.height, depth=args.depth, root=root)
    try:
        p = subprocess.run([exe, "--scene", scene_txt, "--time", "--iters", str(args.ref_gpu_iters), "--warmup", "0"],
                           cwd=os.path.join(root, "bin"), capture_output=True, text=True, timeout=args.ref_gpu_timeout)
        for line in p.stdout.splitlines():
            if line.startswith("REF_GPU_RESULT"):
                r = json.loads(line.split(" ", 1)[1])
                return {"what": "reference apps/src/pathtrace.cu, unmodified, nvcc -O3 sm_100a, same GPU and workload",
                        "iterations": r["iters"], "ms_per_iteration_loop": r["loop_ms_per_iter"],
                        "ms_per_iteration_call": r["call_ms_per_iter"], "mpaths_per_s": r["mpaths_per_s_call"]}
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

    scene_txt = assets.scene_file(SCENE, args.width, args.height, depth=args.depth, root=root)
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
    if rank != 0:
        return
    root, textures = prepare_assets(args.triangles)
    from mygpuraytracer_b200 import standin_mesh

    n_tris = len(standin_mesh.build(args.triangles)[3])
    # Size the sample so that steps+warmup iterations fit in ~150 s: the brute-force
    # reference costs O(paths x triangles), so time scales with the pixel count.
    probe = cpu_sample(root, args, 1, (16, 9))
    per_pixel_s = probe["ms_per_iteration"] * 1e-3 / (16 * 9)
    budget_s = 150.0 / max(1, args.steps + min(args.warmup, 1))
    pixels = max(16 * 9, min(CPU_SAMPLE[0] * CPU_SAMPLE[1], int(budget_s / per_pixel_s)))
    h = max(9, int((pixels * 9 / 16) ** 0.5))
    w = max(16, pixels // h)
    r = cpu_sample(root, args, max(1, args.steps + min(args.warmup, 1)), (w, h))
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": METRIC, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": r["ms_per_iteration"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, n_tris, textures),
        "cpu_baseline": {"value": r["value"], "unit": METRIC, "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]},
        "e2e": {"value": r["value"], "unit": METRIC, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": f"each step is one iteration of the bounded sample ({w}x{h}); Mpaths/s is size independent",
    }
    emit(line)


def run_ours(args, rank: int, world: int, local_rank: int):
    import numpy as np
    import torch

    from mygpuraytracer_b200 import abi, api, assets, standin_mesh

    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    scene_txt = None
    if rank == 0:
        root, textures = prepare_assets(args.triangles)
        scene_txt = assets.scene_file(SCENE, args.width, args.height, depth=args.depth, root=root)
    if dist:
        dist.barrier()
    if rank != 0:
        root, textures = prepare_assets(args.triangles, write=False)
        scene_txt = os.path.join(root, "scenes", f"{SCENE}_{args.width}x{args.height}_d{args.depth}.txt")
    t0 = time.time()
    scene = api.Scene(scene_txt)
    load_s = time.time() - t0
    pod = scene.pod
    n_tris = len(pod.face_pos)
    P = pod.n_pixels
    opt = abi.default_options(device=local_rank)
    # the KC throughput contexts share the GPU (grids sized to a share of the SMs); the e2e leg below uses its
    # own full-width context, because one pathtrace() call at a time wants the lowest latency
    opt_shared = abi.default_options(device=local_rank, concurrent_contexts=max(1, args.streams))
    # KC contexts per GPU, each on its own stream with its own accumulator, render
    # interleaved iteration indices (the same samples-per-pixel sharding used
    # across GPUs).  One iteration is 26 short dependent kernels that cannot fill
    # 148 SMs on their own (a depth-7 launch has 250 k rays); independent
    # iterations in flight overlap each other's tails.
    KC = max(1, args.streams)
    rs = [api.Renderer(scene, opt_shared)]
    if args.private_scenes:
        rs += [api.Renderer(scene, opt_shared) for _ in range(KC - 1)]
    else:  # one copy of the scene (BVH, triangles, textures) on the device, shared by the KC contexts
        rs += [api.Renderer(scene, opt_shared, share=rs[0]) for _ in range(KC - 1)]
    r = rs[0]
    mesh_geom = int(np.nonzero(pod.geoms["type"] == abi.OBJ)[0][0])
    # every context builds its own BVH; the first build of a process also pays CUDA's lazy kernel loading
    bvh = min((x.bvh_info(mesh_geom) for x in rs), key=lambda b: b.build_ms)

    streams = [torch.cuda.Stream() for _ in range(KC)]
    accs = [torch.zeros(P * 3, dtype=torch.float32, device="cuda") for _ in range(KC)]
    for rk, st, ac in zip(rs, streams, accs):
        rk.set_stream_ptr(st.cuda_stream)
        rk.set_device_image_ptr(ac.data_ptr())
    stream, acc = streams[0], accs[0]

    W = args.warmup
    K = args.steps                          # iterations per rank, split over the KC contexts
    lanes = world * KC                      # independent iteration streams in the whole job

    def first_of(c):
        return rank * KC + c + 1

    def barrier():
        if dist:
            dist.barrier()
        torch.cuda.synchronize()

    def run(iter_base, total):
        """Queue `total` iterations over the contexts (exactly), fork/join around stream 0."""
        for c in range(1, KC):
            streams[c].wait_stream(stream)
        for c in range(KC):
            count = total // KC + (1 if c < total % KC else 0)
            with torch.cuda.stream(streams[c]):
                rs[c].render(first_of(c) + iter_base * lanes, count, lanes)
        for c in range(1, KC):
            stream.wait_stream(streams[c])

    # ---- device-resident throughput --------------------------------------------------------
    with torch.cuda.stream(stream):
        run(0, max(W, KC))
        barrier()
        for ac in accs:
            ac.zero_()
        barrier()
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        launches0 = sum(x.launch_count() for x in rs)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        run(max(W, KC), K)
        for c in range(1, KC):
            acc.add_(accs[c])
        if dist:
            dist.reduce(acc, dst=0)
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        clocks = sampler.stop() if rank == 0 else None
        launches = sum(x.launch_count() for x in rs) - launches0
        live = r.live_counts()
    t = torch.tensor([ms], device="cuda")
    if dist:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = world * K * P / (ms_max * 1e-3) / 1e6
    segments = int(live[: args.depth].sum())
    first = first_of(0)

    # ---- per-kernel times of one iteration (CUDA events around every launch) ------------------
    with torch.cuda.stream(stream):
        prof = [r.profile_kernels(first + (2 * W + K + KC + i) * lanes) for i in range(5)]
        barrier()
        walks = r.walk_counts()
        long_walks = r.walk_counts(long_walks=True)
    prof = {k: statistics.median(p[k] for p in prof) for k in prof[0]}

    # ---- end to end: the reference-facing call with host buffers ------------------------------
    host_img = torch.empty(P * 3, dtype=torch.float32).pin_memory()
    host_alb = torch.empty(P * 3, dtype=torch.float32).pin_memory()
    img_np, alb_np = host_img.numpy().reshape(P, 3), host_alb.numpy().reshape(P, 3)
    r_e2e = api.Renderer(scene, opt)
    r_e2e.set_stream_ptr(stream.cuda_stream)
    # the first LBVH build of a process also pays CUDA's lazy kernel loading; this context builds its own
    bvh = min([bvh, r_e2e.bvh_info(mesh_geom)], key=lambda b: b.build_ms)
    with torch.cuda.stream(stream):
        r = r_e2e
        for i in range(min(W, 3)):
            r.pathtrace(first + i * lanes, img_np, alb_np)
        barrier()
        e0.record(stream)
        for i in range(K):
            r.pathtrace(first + (W + i) * lanes, img_np, alb_np)
        if dist:
            dist.reduce(acc, dst=0)
        e1.record(stream)
        barrier()
        e2e_ms = e0.elapsed_time(e1)
    t = torch.tensor([e2e_ms], device="cuda")
    if dist:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * K * P / (float(t.item()) * 1e-3) / 1e6
    checksum = float(np.float64(img_np.sum()))

    # ---- end to end, pipelined: the same one-iteration-per-call contract served by b2pt_pipe_pathtrace -------
    # (csrc/pipe.cu: KC lanes render the next iterations while the host consumes the current one; every call
    # still returns the running sum in host memory, bit-identical to the single-context call above)
    pipe = api.Pipeline(scene, opt, lanes=KC)
    pipe_warm = max(3, min(W, 8))            # call 1 starts the lanes, call 2 learns the stride, call 3 is steady state
    with torch.cuda.stream(stream):
        for i in range(pipe_warm):
            pipe.pathtrace(first + i * lanes, img_np, alb_np)
        barrier()
        misses0 = pipe.misses()
        e0.record(stream)
        for i in range(K):
            pipe.pathtrace(first + (pipe_warm + i) * lanes, img_np, alb_np)
        if dist:
            dist.reduce(acc, dst=0)
        e1.record(stream)
        barrier()
        pipe_ms = e0.elapsed_time(e1)
        pipe_misses = pipe.misses() - misses0
    tp = torch.tensor([pipe_ms], device="cuda")
    if dist:
        dist.all_reduce(tp, op=dist.ReduceOp.MAX)
    pipe_value = world * K * P / (float(tp.item()) * 1e-3) / 1e6
    pipe.close()
    with torch.cuda.stream(stream):
        prof_full = [r_e2e.profile_kernels(first + (3 * W + 2 * K + KC + i) * lanes) for i in range(5)]
        barrier()
    prof_full = {k: statistics.median(p[k] for p in prof_full) for k in prof_full[0]}

    if rank == 0:
        peak, peak_src = hbm_peak()
        n_walks = int(walks[: args.depth].sum())
        n_long = int(long_walks[: args.depth].sum())
        # algorithmic bytes per launch unit (kernel table)
        kern = {
            # ray read 24 B (32 as stored), hit record 32 B, key 1 B, survival flag 1 B
            "k_intersect_analytic": ("analytic", 58.0 * segments),
            # per queued ray: queue slot 4 B, ray 24 B, closest analytic t 4 B + geom 4 B read; (t, bary, ids) 20 B + key 1 B written
            "k_mesh_walk": ("walk", 57.0 * n_walks),
            "k_mesh_walk_long": ("walk_long", 57.0 * n_long),
            # per queued ray: queue slot 4 B + partial record 20 B read, record 32 B + flag 1 B written
            "k_mesh_finish": ("finish", 57.0 * n_walks),
            "k_sort_material": ("sort", 10.0 * segments),      # key 1 B + flag 1 B read, permutation 4 B + compaction rank 4 B written
            "k_shade_compact": ("shade", 132.0 * segments),    # permutation 4 B + hit 32 B + state 48 B read, state 48 B written
        }
        dom = max(kern, key=lambda k: prof[kern[k][0]])
        dom_ms = prof[kern[dom][0]]
        dom_bytes = kern[dom][1]
        dom_gbs = dom_bytes / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
        # every kernel of the iteration against the HBM roofline: algorithmic bytes (kernel table) over
        # its device time in one iteration of a context running alone with the shared-SM grids of the timed region
        kern_all = dict(kern)
        kern_all["k_generate"] = ("generate", 44.0 * P)
        roofline_kernels = {}
        for name, (key, nbytes) in kern_all.items():
            kms = prof.get(key, 0.0)
            gbs = nbytes / (kms * 1e-3) / 1e9 if kms > 0 else 0.0
            roofline_kernels[name] = {"ms_per_step": kms, "algorithmic_bytes_per_step": nbytes, "achieved": gbs,
                                      "unit": "GB/s", "frac": gbs / peak}
        iter_bytes = 84.0 * P + 280.0 * segments
        iter_gbs = iter_bytes / (ms_max / K * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": METRIC, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(args, n_tris, textures),
            "e2e": {"value": pipe_value, "unit": METRIC, "h2d_bytes_per_step": 8, "d2h_bytes_per_step": P * 12,
                    "d2h_note": "running sum every step; the albedo AOV (P*12 more) only when it changed (iteration 1)",
                    "call": "b2pt_pipe_pathtrace(pipe, iter, host_image, host_albedo) -- pathtrace() of apps/src/pathtrace.h:9, "
                            "one iteration per call, host image after every call",
                    "ms_per_step": float(tp.item()) / K, "lanes": KC, "mispredicted_calls": int(pipe_misses),
                    "note": "steady state of the pipeline: the lanes hold the next %d iterations when the timed region "
                            "starts and when it ends; K iterations are rendered and K consumed inside it.  The copy to the host "
                            "and the merge overlap the other lanes' kernels, so this leg is bound by the same %d-context GPU "
                            "throughput as `value` (e2e_single_context is the unoverlapped form)" % (KC, KC)},
            "e2e_single_context": {"value": e2e_value, "unit": METRIC, "ms_per_step": float(t.item()) / K,
                                   "call": "b2pt_pathtrace(ctx, iter, host_image, host_albedo): render, then copy, nothing overlapped"},
            "gpu_launches": int(launches), "streams_per_gpu": KC,
            "clocks": clocks,
            "roofline": {"kernel": dom, "bound": "hbm", "achieved": dom_gbs, "peak": peak, "unit": "GB/s",
                         "frac": dom_gbs / peak, "traffic": TRAFFIC.get(dom), "peak_source": peak_src,
                         "traffic_note": "dram__bytes_read+write of the depth-1 launch under ncu --set full (cold caches); it "
                                         "includes the scene traffic (BVH nodes, triangles, texels) that the algorithmic "
                                         "bytes leave out (SURVEY.md 8d)",
                         "algorithmic_bytes_per_launch": dom_bytes / args.depth, "launches_per_step": args.depth,
                         "ms_per_launch": dom_ms / args.depth, "walks_per_step": n_walks, "long_walks_per_step": n_long,
                         "note": "dominant kernel by device time measured with CUDA events in this run; achieved = algorithmic "
                                 "bytes per launch / average launch duration.  The BVH walk is issue/latency bound, not HBM "
                                 "bound: see profiles/ for issue-slot, FP32-pipe and divergence counters"},
            "roofline_iter": {"bytes_per_step": iter_bytes, "achieved": iter_gbs, "peak": peak, "unit": "GB/s",
                              "frac": iter_gbs / peak, "formula": "84*P + 280*S"},
            "roofline_kernels": roofline_kernels,
            "kernel_ms_per_step": prof,
            "kernel_ms_per_step_full_width": prof_full,
            "kernel_ms_note": "CUDA events around every launch of one iteration run alone: with the grids of the timed "
                              "region (contexts share the SMs) and with the full-width grids of the e2e context, where "
                              "the analytic intersection is fused into generate / shade (k_generate_trace, k_shade_trace)",
            "segments_per_step": segments, "live_paths_per_depth": [int(x) for x in live[: args.depth + 1]],
            "bvh": {"triangles": int(bvh.n_faces), "nodes": int(bvh.n_nodes), "max_depth": int(bvh.max_depth),
                    "build_ms": float(bvh.build_ms), "build_ms_note": "device time, fastest build of this process (the first one also pays CUDA's lazy kernel loading)"},
            "scene_load_s": round(load_s, 2), "image_checksum": checksum,
        }
        if world == 1 and not args.no_cpu:
            ra = reference_host_adapter(root, args, max(20, min(K, 200)))
            if ra is not None:
                line["e2e_reference_host"] = ra
            line["cpu_baseline"] = {k: v for k, v in cpu_sample(root, args, 1).items()}
            rg = reference_gpu(root, args)
            if rg is not None:
                line["reference_gpu"] = rg
        emit(line)
    for x in rs + [r_e2e]:
        x.close()
    if dist:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--width", type=int, default=WIDTH)
    ap.add_argument("--height", type=int, default=HEIGHT)
    ap.add_argument("--depth", type=int, default=DEPTH)
    ap.add_argument("--triangles", type=int, default=250_000)
    ap.add_argument("--streams", type=int, default=4, help="concurrent iteration streams (contexts) per GPU")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline / reference_gpu legs")
    ap.add_argument("--private-scenes", action="store_true", help="every context uploads its own copy of the scene (A/B of b2pt_create_shared)")
    ap.add_argument("--ref-gpu-iters", type=int, default=1)
    ap.add_argument("--ref-gpu-timeout", type=int, default=240)
    args = ap.parse_args()
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
