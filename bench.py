#!/usr/bin/env python3
"""bench.py - the trace path on BASELINE.json's headline configuration.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2|cfg1]

One "step" = one 1920x1080 frame of the workload through the trace path (no anti-aliasing):
  cfg2 (default)  synthetic 999,698-triangle mesh2 height field with its BBox tree, 2 point lights with shadows,
                  reflection to max_trace_level 5, checker floor + 2 mirror spheres      (BASELINE.json configs[1])
  cfg1            1024 spheres + checker plane, 1 point light with shadows               (BASELINE.json configs[0])

metric  = Mrays/s, rays counted like the reference's statistics page: Number_Of_Rays + Shadow_Ray_Tests.
value   = device-timed (CUDA events), scene tables and output buffer resident in HBM.
e2e     = same frame through pvgpu_render() with HOST buffers (rectangle list in, RGBT float frame out),
          wall clock around the C-ABI call.
N > 1   = one process per GPU (torchrun); the frame's 32x32 tiles are dealt round-robin to the ranks, the scene is
          replicated, and the finished tiles are gathered to rank 0 (the only NVLink traffic).  Total work is one
          frame whatever N is -> "strong" scaling.  e2e at N > 1 ends with the ASSEMBLED frame in rank 0's host memory
          (NCCL gather + one D2H inside the timed region).  --inproc-gpus M instead shards one process's frame over M
          devices inside the library (pvgpu_scene_finalize_multi: atomic tile counter + peer copies).
--impl reference = the UNMODIFIED reference binary (oracle/_ref/*/povray, built from /root/reference by
          oracle/build_ref.sh) rendering the same scene file with +WT<host threads>; its "Trace Time" and its
          Rays / Shadow Ray Tests counters give the same metric.  Rank 0 only.
"""
import argparse
import json
import os
import re
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W, H, BLOCK = 1920, 1080, 32
WORKLOADS = {
    "cfg2": "synthetic 999698-triangle mesh2 height field + BBox tree, 2 lights with shadows, reflection max_trace_level 5, 1920x1080 -A",
    "cfg1": "synthetic 1024 spheres + checker plane, 1 point light with shadows, 1920x1080 -A",
    "cfg3": "synthetic 4096 CSG objects (difference / intersection / merge of boxes, spheres, quadrics), refraction ior 1.3-1.6, "
            "2 lights, max_trace_level 6, 1920x1080 +A0.3 +AM2 +R3 +J (scene tables from the reference parser through the adapter)",
    "cfg3_noaa": "config 3 without anti-aliasing (-A)",
    "cfg4": "synthetic 2048 tori (half sturm) with granite / bozo pigments, colour maps, turbulence, noise_generator 2 and 3, 1920x1080 -A "
            "(scene tables from the reference parser through the adapter)",
}
# anti-aliasing of the workload: (method, depth, threshold, jitter amount) and the reference's switches for it
WORKLOAD_AA = {"cfg3": ((2, 3, 0.3, 1.0), ["+A0.3", "+AM2", "+R3", "+J"])}
# Algorithmic work of a traversal kernel (SURVEY.md section 8d / DESIGN.md "Roofline"), from the DEVICE's own counters of the timed
# frames: every bounding-box test reads one 32 B node and costs 24 flop (6 sub, 6 mul, 12 compares: boundingbox.cpp:566-625);
# every primitive test reads the record(s) below and costs the FP64 flop below; every ray moves its queue records through HBM.
NODE_BYTES, NODE_FLOP = 32, 24
PRIM_BYTES = {"cfg1": 168, "cfg2": 64, "cfg3": 4 * 168, "cfg3_noaa": 4 * 168, "cfg4": 168 + 256}     # object / triangle record (+ transform)
PRIM_FLOP = {"cfg1": 40, "cfg2": 45, "cfg3": 250, "cfg3_noaa": 250, "cfg4": 400}
RAY_BYTES = {"k_closest": 96 + 48, "k_shadow": 96 + 16}        # PRay in + HitRec out; SRay in + one RGBT accumulation
ADAPTER = os.path.join(ROOT, "oracle", "_ref", "parity", "povray-gpu")


def make_builder(workload):
    from povray_b200 import synth
    return synth.mesh_scene(708) if workload == "cfg2" else synth.spheres_scene(1024)


def pov_text(workload):
    from povray_b200 import synth
    return synth.csg_scene_pov(4096) if workload.startswith("cfg3") else synth.torus_scene_pov(2048)


def build_scene(workload):
    """The flattened scene of the workload: configs 1 and 2 from the Python table builder, configs 3 and 4 from the reference
    parser through the reference-side adapter (povray-gpu dumps the tables it hands to pvgpu_scene_* and renders 8x8 pixels)."""
    import povray_b200 as pv
    if workload in ("cfg1", "cfg2"):
        return make_builder(workload).build()
    if not os.path.exists(ADAPTER):
        raise SystemExit(f"bench.py: workload {workload} needs the reference-side adapter {ADAPTER}")
    with tempfile.TemporaryDirectory() as d:
        pov, pvs = os.path.join(d, "s.pov"), os.path.join(d, "s.pvs")
        open(pov, "w").write(pov_text(workload))
        r = subprocess.run([ADAPTER, "+I" + pov, "+O" + os.path.join(d, "o.png"), "+W8", "+H8", "-A", "-D", "+WT1", "-GA"],
                           env=dict(os.environ, PVGPU_RENDER="gpu", PVGPU_DUMP_SCENE=pvs, PVGPU_DEVICE=os.environ.get("LOCAL_RANK", "0")),
                           capture_output=True, text=True, cwd=d)
        if r.returncode != 0 or not os.path.exists(pvs):
            raise SystemExit("bench.py: adapter failed:\n" + (r.stdout + r.stderr)[-2000:])
        return pv.Scene.load(pvs)


# ----------------------------------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.25)
            self.proc.terminate()
            self.thread.join(timeout=2)

    def summary(self):
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx = max(mx, float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------------
# reference arm (CPU)
# ----------------------------------------------------------------------------------------------------------
REF_FLAGS = {"fast": "-O3 -march=x86-64-v3 -fno-fast-math, portable noise", "parity": "-O2 -fno-fast-math -ffp-contract=off, portable noise",
             "stock": "-O3 -ffast-math -march=x86-64-v3 + the reference's AVX / AVX2-FMA3 noise (unix/configure.ac:750-763 defaults, -march=native replaced by the portable x86-64-v3)"}


def reference_binary():
    for variant in ("fast", "parity"):
        p = os.path.join(ROOT, "oracle", "_ref", variant, "povray")
        if os.path.exists(p):
            return p, variant
    return None, None


def run_reference_once(binary, pov, threads, width=W, height=H, aa_flags=("-A",)):
    """One render of `pov` by the unmodified reference; returns (trace seconds, rays, shadow tests, parse seconds)."""
    with tempfile.TemporaryDirectory() as d:
        r = subprocess.run([binary, "+I" + pov, "+O" + os.path.join(d, "o.png"), f"+W{width}", f"+H{height}", "-D", f"+WT{threads}",
                            "-GD", "-GR", "-GW", "-GF", "+GS"] + list(aa_flags), capture_output=True, text=True, cwd=d)
    out = (r.stdout + r.stderr).replace("\r", "\n")
    if r.returncode != 0:
        raise RuntimeError("reference render failed:\n" + out[-2000:])
    trace = float(re.search(r"Trace Time:.*?\(([\d.]+) seconds\)", out).group(1))
    parse = float(re.search(r"Parse Time:.*?\(([\d.]+) seconds\)", out).group(1))
    rays = int(re.search(r"Rays:\s+(\d+)", out).group(1))
    m = re.search(r"Shadow Ray Tests:\s+(\d+)", out)
    shadow = int(m.group(1)) if m else 0
    return trace, rays, shadow, parse


def write_pov(workload, directory):
    path = os.path.join(directory, workload + ".pov")
    with open(path, "w") as f:
        if workload in ("cfg1", "cfg2"):
            make_builder(workload).to_pov(f)
        else:
            f.write(pov_text(workload))
    return path


def aa_flags(workload):
    return WORKLOAD_AA[workload][1] if workload in WORKLOAD_AA else ["-A"]


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    binary, variant = reference_binary()
    threads = os.cpu_count() or 1
    base = {"impl": "reference", "metric": "Mrays/s", "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOADS[args.workload].replace("1920x1080", f"{W}x{H}"), "width": W, "height": H, "aa": " ".join(aa_flags(args.workload)), "tile": BLOCK}}
    if binary is None:
        base["unavailable"] = "oracle/_ref/*/povray not built (needs /root/reference at build time)"
        print(json.dumps(base))
        return
    with tempfile.TemporaryDirectory() as d:
        pov = write_pov(args.workload, d)
        for _ in range(args.warmup):
            run_reference_once(binary, pov, threads, aa_flags=aa_flags(args.workload))
        t, rays, shadow, parse = 0.0, 0, 0, 0.0
        for _ in range(args.steps):
            tr, r, s, p = run_reference_once(binary, pov, threads, aa_flags=aa_flags(args.workload))
            t += tr; rays += r; shadow += s; parse += p
    value = (rays + shadow) / t / 1e6
    flags = REF_FLAGS[variant]
    base.update({"value": value, "ms_per_step": 1e3 * t / args.steps, "sec_per_frame": t / args.steps,
                 "rays_per_step": (rays + shadow) / args.steps, "gpu_launches": 0,
                 "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": threads, "kind": "reference",
                                  "sample": f"{args.steps} full {W}x{H} frames by oracle/_ref/{variant}/povray ({flags}) +WT{threads}; "
                                            f"time = the reference's own 'Trace Time' (parse {parse / args.steps:.1f} s/frame excluded)"},
                 "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
    print(json.dumps(base))


# ----------------------------------------------------------------------------------------------------------
# our arm (GPU)
# ----------------------------------------------------------------------------------------------------------
def load_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


def roofline_of(workload, stats, steps, frame_ms, fp64_peak_tflops):
    """Roofline entries of the two traversal kernel families from the library's own per-launch CUDA events and the device's own
    node / primitive test counters (pvgpu_stats), for the frames that were timed."""
    peaks = load_peaks()
    peak, which = (peaks["hbm_gbs"], "measured") if "hbm_gbs" in peaks else (6650.0, "fallback")
    fam = {}
    for name, ms_key, n_key, rays_key, node_key, prim_key in (
            ("k_closest", "closest_ms", "closest_launches", "closest_rays", "node_tests_closest", "prim_tests_closest"),
            ("k_shadow", "shadow_ms", "shadow_launches", "shadow_rays", "node_tests_shadow", "prim_tests_shadow")):
        ms = sum(s[ms_key] for s in stats)
        n = sum(s[n_key] for s in stats)
        rays = sum(s[rays_key] for s in stats)
        nodes = sum(s.get(node_key, 0) for s in stats)
        prims = sum(s.get(prim_key, 0) for s in stats)
        if ms <= 0 or n == 0:
            continue
        alg_bytes = nodes * NODE_BYTES + prims * PRIM_BYTES[workload] + rays * RAY_BYTES[name]
        alg_flop = nodes * NODE_FLOP + prims * PRIM_FLOP[workload]
        fam[name] = {"ms": ms, "launches": n, "rays": rays, "node_tests": nodes, "prim_tests": prims, "alg_bytes": alg_bytes, "alg_flop": alg_flop}
    if not fam:
        return None
    dom = max(fam, key=lambda k: fam[k]["ms"])
    f = fam[dom]
    achieved = f["alg_bytes"] / (f["ms"] * 1e-3) / 1e9
    traffic, traffic_note = None, "no ncu capture of this build committed"
    try:      # DRAM bytes per launch of this kernel family from the committed ncu --set full capture of THIS source revision
        tj = json.load(open(os.path.join(ROOT, "profiles", "r2_traffic.json")))
        if workload in tj.get("workloads", {}) and dom in tj["workloads"][workload]:
            b = tj["workloads"][workload][dom]["dram_bytes_per_launch"]
            traffic = sum(b) / len(b)
            traffic_note = f"mean dram__bytes_read+write of the launches captured by ncu --set full at commit {tj.get('commit', '?')} (profiles/r2_traffic.json)"
    except Exception:
        pass
    out = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "peak_source": which,
           "traffic": traffic, "traffic_note": traffic_note,
           "achieved_dram_gbs": (traffic / (f["ms"] / f["launches"] * 1e-3) / 1e9) if traffic else None,
           "alg_bytes_per_launch": f["alg_bytes"] / f["launches"], "alg_bytes_per_ray": f["alg_bytes"] / max(f["rays"], 1),
           "node_tests_per_ray": f["node_tests"] / max(f["rays"], 1), "prim_tests_per_ray": f["prim_tests"] / max(f["rays"], 1),
           "counts": "device counters of the timed frames (pvgpu_stats.node_tests_* / prim_tests_*)",
           "rays_per_launch": f["rays"] / f["launches"], "avg_launch_ms": f["ms"] / f["launches"],
           "share_of_step": f["ms"] / (frame_ms * steps),
           "share_note": "k_shadow_* of wave k runs concurrently with k_closest / k_shade of wave k+1 on a second stream: the shares of the families add up to more than 1",
           "kernel_ms_per_step": {"k_primary": sum(s["primary_ms"] for s in stats) / steps, "k_closest": sum(s["closest_ms"] for s in stats) / steps,
                                  "k_shade": sum(s["shade_ms"] for s in stats) / steps, "k_shadow": sum(s["shadow_ms"] for s in stats) / steps}}
    if fp64_peak_tflops:
        fl = f["alg_flop"] / (f["ms"] * 1e-3) / 1e12
        out["fp64"] = {"achieved": fl, "peak": fp64_peak_tflops, "unit": "TFLOP/s", "frac": fl / fp64_peak_tflops,
                       "peak_source": "DFMA loop timed in this run (pvgpu_fp64_peak)", "alg_flop_per_ray": f["alg_flop"] / max(f["rays"], 1)}
    return out


def bench_workload(args, workload, steps, warmup, rank, local_rank, world, dist, inproc):
    """Device-timed value + e2e of one workload; returns (dict for the JSON line, scene, aa) on rank 0 (None elsewhere)."""
    import torch
    import povray_b200 as pv
    from povray_b200 import _abi as A
    from povray_b200 import shard
    from povray_b200.scene import _rect_array, _area

    dev = torch.device("cuda", local_rank)
    t0 = time.time()
    scene = build_scene(workload)
    t_build = time.time() - t0
    aa = None
    if workload in WORKLOAD_AA:
        m, dep, thr, jit = WORKLOAD_AA[workload][0]
        aa = A.AA()
        aa.method, aa.depth, aa.threshold, aa.jitter_scale, aa.gamma = m, dep, thr, jit, 2.5
    t0 = time.time()
    if inproc > 1:
        scene.finalize_multi(n_devices=inproc)
    else:
        scene.finalize(local_rank)
    t_upload = time.time() - t0

    all_tiles = pv.tiles(W, H, BLOCK)
    mine = shard.deal(all_tiles, rank, world)          # round-robin deal: statistically balanced, no exchange needed
    if world == 1 and args.emulate_world > 1:          # development aid: the share of rank 0 of an N-rank run, on one GPU
        mine = shard.deal(all_tiles, 0, args.emulate_world)
    rect_arr = _rect_array(mine)
    n_px = _area(mine)
    max_px = shard.padded_pixels(all_tiles, world)
    out = torch.zeros(max_px * 4, dtype=torch.float32, device=dev)
    gathered = [torch.empty_like(out) for _ in range(world)] if (world > 1 and rank == 0) else None
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2
    stream = torch.cuda.current_stream()

    def step_device():
        st = scene.render_device(W, H, rect_arr, out.data_ptr(), stream.cuda_stream, aa=aa)
        if world > 1:
            dist.gather(out, gathered, dst=0)
        return st

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(warmup, 3)):
        step_device()
    barrier()
    stats = []
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    with ClockSampler(local_rank) as clocks:
        for k in range(steps):
            flush.zero_()                              # evict the scene tables and queues from L2 between timed frames
            barrier()
            ev[k][0].record()
            stats.append(step_device())
            ev[k][1].record()
        barrier()
    ms = sum(a.elapsed_time(b) for a, b in ev)
    rays = sum(s["rays"] + s["shadow_ray_tests"] for s in stats)
    launches = sum(s["kernel_launches"] for s in stats)

    # end to end: HOST buffers through the C-ABI call (rectangle list in, RGBT float frame out), wall clock.  The frame lands in
    # page-locked memory obtained from pvgpu_host_alloc (what the adapter uses); the same call with an ordinary pageable numpy array
    # goes through the library's staging copy and is reported as e2e.pageable_ms_per_step.  At N > 1 (torchrun) every rank renders
    # its tiles into device memory, NCCL gathers them on rank 0, and rank 0 copies the assembled frame to its host: the region
    # ends when the whole frame is in ONE host buffer.
    e2e_rays = 0
    if world == 1:
        hb = pv.HostBuffer(max_px * 4)
        for _ in range(2):
            scene.render(W, H, mine, out=hb.array, aa=aa)
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            px, st = scene.render(W, H, mine, out=hb.array, aa=aa)
            e2e_rays += st["rays"] + st["shadow_ray_tests"]
        e2e_s = time.perf_counter() - t0
        scene.render(W, H, mine, aa=aa)
        t0 = time.perf_counter()
        for _ in range(min(steps, 3)):
            scene.render(W, H, mine, aa=aa)
        e2e_pageable_s = (time.perf_counter() - t0) / min(steps, 3)
        d2h = n_px * 16
    else:
        host_frame = torch.empty(world * max_px * 4, dtype=torch.float32).pin_memory() if rank == 0 else None
        frame_dev = torch.empty(world * max_px * 4, dtype=torch.float32, device=dev) if rank == 0 else None
        chunks = list(frame_dev.split(max_px * 4)) if rank == 0 else None

        def step_e2e():
            st = scene.render_device(W, H, rect_arr, out.data_ptr(), stream.cuda_stream, aa=aa)
            dist.gather(out, chunks, dst=0)
            if rank == 0:
                host_frame.copy_(frame_dev, non_blocking=True)
            torch.cuda.synchronize()
            return st
        for _ in range(2):
            step_e2e()
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            st = step_e2e()
            e2e_rays += st["rays"] + st["shadow_ray_tests"]
        barrier()
        e2e_s = time.perf_counter() - t0
        e2e_pageable_s = None
        d2h = world * max_px * 16

    vals = torch.tensor([ms, float(rays), float(launches), e2e_s, float(e2e_rays)], dtype=torch.float64, device=dev)
    if world > 1:
        mx = vals.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = vals.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        ms, e2e_s = float(mx[0]), float(mx[3])
        rays, launches, e2e_rays = float(sm[1]), float(sm[2]), float(sm[4])
    if rank != 0:
        return None, scene, aa
    res = {"value": rays / (ms * 1e-3) / 1e6, "ms_per_step": ms / steps, "rays_per_frame": rays / steps, "gpu_launches": int(launches),
           "clocks": clocks.summary(), "t_build": t_build, "t_upload": t_upload, "scene_device_bytes": scene.device_bytes,
           "n_tiles": len(all_tiles), "stats": stats,
           "e2e": {"value": e2e_rays / e2e_s / 1e6, "unit": "Mrays/s", "ms_per_step": 1e3 * e2e_s / steps,
                   "h2d_bytes_per_step": len(mine) * 16 * world, "d2h_bytes_per_step": d2h,
                   "region": ("pvgpu_render: rectangle list in, RGBT frame out into page-locked host memory" if world == 1 else
                              "per rank pvgpu_render_device, NCCL gather to rank 0, D2H of the assembled frame into one pinned host buffer")}}
    if e2e_pageable_s is not None:
        res["e2e"]["pageable_ms_per_step"] = 1e3 * e2e_pageable_s
    return res, scene, aa


def ours(args):
    import torch
    from povray_b200 import _abi as A

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the trace path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        # NCCL prints its version banner on stdout at the first collective; the contract is ONE JSON line there, so
        # stdout is pointed at stderr (file-descriptor level) until the communicator is up
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            warm = torch.zeros(1, device=torch.device("cuda", local_rank))
            dist.all_reduce(warm)
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)

    inproc = args.inproc_gpus if world == 1 else 1
    res, scene, aa = bench_workload(args, args.workload, args.steps, args.warmup, rank, local_rank, world, dist, inproc)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    n_gpus = world if world > 1 else max(1, inproc)
    sharding = (f"{res['n_tiles']} tiles dealt round-robin over {world} GPU(s) (one process each), scene replicated, NCCL gather to rank 0" if inproc <= 1 else
                f"{res['n_tiles']} tiles sharded over {inproc} GPUs inside the library: one host thread per GPU, chunks from one atomic counter, peer-copy gather")
    line = {"metric": "Mrays/s", "value": res["value"], "unit": "Mrays/s", "n_gpus": n_gpus, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": res["ms_per_step"], "sec_per_frame": res["ms_per_step"] / 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOADS[args.workload].replace("1920x1080", f"{W}x{H}"), "width": W, "height": H, "aa": " ".join(aa_flags(args.workload)), "tile": BLOCK,
                       "sharding": sharding,
                       "l2": "256 MB flush write between timed frames; scene tables + ray queues exceed the 126 MB L2",
                       "scene_device_bytes": res["scene_device_bytes"], "scene_build_s": round(res["t_build"], 2), "scene_upload_s": round(res["t_upload"], 3),
                       "rays_per_frame": res["rays_per_frame"]},
            "gpu_launches": res["gpu_launches"], "clocks": res["clocks"], "e2e": res["e2e"]}

    # FP64 vector peak of this GPU, measured in this run by a DFMA loop of the library (the second roofline bound of SURVEY 8d)
    fp64_peak = None
    try:
        v = A.C.c_double(0.0)
        if A.lib().pvgpu_fp64_peak(local_rank, A.C.byref(v)) == 0:
            fp64_peak = v.value
    except Exception:
        pass
    line["roofline"] = roofline_of(args.workload, res["stats"], args.steps, res["ms_per_step"], fp64_peak)

    # the other BASELINE.json configurations at the same frame size, short runs, in the same line (N = 1 only)
    if world == 1 and inproc <= 1 and not args.no_other_workloads:
        others = {}
        for wl in ("cfg1", "cfg3_noaa", "cfg3", "cfg4"):
            if wl == args.workload:
                continue
            try:
                del scene
                r, scene, _ = bench_workload(args, wl, 3, 3, rank, local_rank, world, dist, 1)
                rf = roofline_of(wl, r["stats"], 3, r["ms_per_step"], fp64_peak)
                others[wl] = {"workload": WORKLOADS[wl], "value": r["value"], "unit": "Mrays/s", "ms_per_step": r["ms_per_step"], "steps": 3,
                              "e2e_ms_per_step": r["e2e"]["ms_per_step"], "rays_per_frame": r["rays_per_frame"],
                              "roofline": {k: rf[k] for k in ("kernel", "frac", "achieved", "alg_bytes_per_ray", "node_tests_per_ray", "prim_tests_per_ray", "kernel_ms_per_step")} if rf else None,
                              "roofline_fp64_frac": rf["fp64"]["frac"] if rf and "fp64" in rf else None}
            except BaseException as e:      # e.g. the reference-side adapter is missing for configs 3 / 4
                others[wl] = {"error": str(e)[:200]}
        line["config"]["other_workloads"] = others

    # CPU baseline on this box's host cores: the unmodified reference on a bounded sample of the same workload
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        base = []
        for variant in ("fast", "stock"):
            binary = os.path.join(ROOT, "oracle", "_ref", variant, "povray")
            if not os.path.exists(binary):
                continue
            try:
                with tempfile.TemporaryDirectory() as d:
                    pov = write_pov(args.workload, d)
                    tr, r, s, parse = run_reference_once(binary, pov, threads, aa_flags=aa_flags(args.workload))
                base.append({"value": (r + s) / tr / 1e6, "unit": "Mrays/s", "cores": threads, "kind": "reference", "build": REF_FLAGS[variant],
                             "sec_per_frame": tr, "parse_s": parse,
                             "sample": f"1 full {W}x{H} frame by oracle/_ref/{variant}/povray +WT{threads}, the reference's own Trace Time"})
            except Exception as e:
                base.append({"error": f"{variant}: {e}"[:300]})
        good = [b for b in base if "value" in b]
        if good:
            line["cpu_baseline"] = good[0]
            if len(good) > 1:
                line["cpu_baseline_stock_flags"] = good[1]
        else:      # the oracle port is the fallback checker-as-baseline
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            import oracle_lib
            with tempfile.TemporaryDirectory() as d:
                p = os.path.join(d, "s.pvs")
                scene.save(p)
                o = oracle_lib.OracleScene(p)
                t0 = time.perf_counter()
                _, ost = o.render(W, H, rect=(0, 476, W - 1, 603), threads=threads)      # 128 rows through the middle of the frame (no AA)
                dt = time.perf_counter() - t0
            line["cpu_baseline"] = {"value": (ost["rays"] + ost["shadow_ray_tests"]) / dt / 1e6, "unit": "Mrays/s", "cores": threads, "kind": "port",
                                    "sample": f"rows 476-603 of the {W}x{H} frame by oracle/libpvoracle.so with {threads} threads ({base})"}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-workloads", action="store_true", help="skip the short runs of the other configurations (config.other_workloads)")
    ap.add_argument("--emulate-world", type=int, default=1, help="development aid: render only rank 0's share of an N-rank run (one GPU)")
    ap.add_argument("--inproc-gpus", type=int, default=1, help="shard the frame over this many GPUs inside ONE process (pvgpu_scene_finalize_multi)")
    ap.add_argument("--width", type=int, default=1920, help="frame width (BASELINE.json's metric is quoted at 1920x1080)")
    ap.add_argument("--height", type=int, default=1080)
    args = ap.parse_args()
    global W, H
    W, H = args.width, args.height
    if args.impl == "reference":
        reference_arm(args)
    else:
        ours(args)


if __name__ == "__main__":
    main()
