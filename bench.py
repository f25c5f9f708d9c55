#!/usr/bin/env python
"""bench.py — the hot-path benchmark (driver contract).

    python bench.py --gpus N --steps K --warmup W              # our CUDA path
    python bench.py --impl reference --gpus N --steps K ...    # the reference's own CPU code (oracle/_ref)

Workload (BASELINE.json configs[1], the configuration the metric is quoted on): forward render of a
256^3 SH-degree-2 dense grid (synthetic LCG grid of SURVEY 8d, density scale D=20), 800x800, 100 orbit
views of main.cpp:99-102 (360/100 degrees apart).  One STEP = one pass over the 100 views = 64 M rays.
With N GPUs each rank holds a grid replica and renders its own 100 views of a 100*N-view orbit (views
interleaved over ranks): per-GPU work fixed => "scaling": "weak", no data-path collective.

  value     Mrays/s with everything resident in HBM (grid, output frames), CUDA events on the launching
            stream, max over ranks.  A step is ONE call of lrf_render_views_dev_async: the step's views in one launch
            of the forward kernel (no idle SMs between views), frame v of a views x frame device buffer.
  e2e       same metric through the public host API (RayMarcherExample.render_views == lrf_render_views): the view
            matrices go host->device and every finished frame device->pinned host inside the timed region,
            synchronous like kernel2D_RayMarch; three of the frames are compared with kernel2D_RayMarch's afterwards.
  roofline  dominant kernel render_lean_kernel (lrf_lean.cuh): algorithmic bytes per launch = sum over the launch's views of
            52 B x (distinct grid nodes read by shaded samples of the view, counted on the device) + 4 B x rays, / mean launch time.
  cpu_baseline  the reference's kernel2D_RayMarch compiled from /root/reference (oracle/_ref/libref_O3.so,
            -O3 -fopenmp -ffp-contract=off) on the same grid + camera, bounded sample (a 96x96 frame of view 0).
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
sys.path.insert(0, os.path.join(ROOT, "tests"))

GS = 256
W = H = 800
VIEWS = 100
DENSITY_SCALE = 20.0
ASPECT = 1.0
FILL_MODE = None         # None: host LCG grid uploaded (default workload); 0/1/2: generated on the device (lrf_fill_synthetic)
SPARSE = False
WORKLOAD = "render256"
# --workload: the default is the configuration the metric is quoted on (BASELINE.json configs[1]); the others are the larger
# configurations, measured the same way (their lines are kept under profiles/, the driver only runs the default)
WORKLOADS = {
    "render256": dict(gs=256, w=800, h=800, views=100, D=20.0, aspect=1.0, fill=None, sparse=False,
                      text="forward render, dense 256^3 SH-degree-2 grid (13 fp32/node, synthetic LCG, density scale 20), 800x800, 100 orbit views per GPU (BASELINE.json configs[1])"),
    "sparse512": dict(gs=512, w=800, h=800, views=20, D=20.0, aspect=1.0, fill=2, sparse=True,
                      text="forward render, sparse box set at 512^3 (density > 0 inside a sphere of radius 0.288, boxes kept by the RebuildOctree criterion: ~10 % of 511^3), 800x800, 20 orbit views per GPU (BASELINE.json configs[3])"),
    "hires512_1080p": dict(gs=512, w=1920, h=1080, views=8, D=20.0, aspect=16.0 / 9.0, fill=0, sparse=False,
                           text="forward render, dense 512^3 grid (synthetic LCG, density scale 20), 1920x1080 frames, 8 orbit views per GPU (BASELINE.json configs[4])"),
    "hires512_4k": dict(gs=512, w=3840, h=2160, views=8, D=20.0, aspect=16.0 / 9.0, fill=0, sparse=False,
                        text="forward render, dense 512^3 grid (synthetic LCG, density scale 20), 3840x2160 frames, 8 orbit views per GPU (BASELINE.json configs[4])"),
}


def select_workload(name):
    global GS, W, H, VIEWS, DENSITY_SCALE, ASPECT, FILL_MODE, SPARSE, WORKLOAD
    wl = WORKLOADS[name]
    GS, W, H, VIEWS, DENSITY_SCALE, ASPECT, FILL_MODE, SPARSE, WORKLOAD = wl["gs"], wl["w"], wl["h"], wl["views"], wl["D"], wl["aspect"], wl["fill"], wl["sparse"], name
CPU_SAMPLE = 96          # cpu_baseline / reference arm: one CPU_SAMPLE x CPU_SAMPLE frame per step
METRIC = "forward_render_Mrays_per_s"


def lcg_grid(gs, D):
    import libs  # numpy generator only (no oracle code is executed by it)
    return libs.lcg_grid(gs, D)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (profiling recipe's clocks line)"""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for (t, r) in self.rows if t0 <= t <= t1] or [r for (_, r) in self.rows]
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1])); pw.append(float(r[2]))
                for nm, v in zip(names, r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "power_w_max": max(pw) if pw else None, "samples": len(sm)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(views_per_launch):
    """(dram bytes per launch of the dominant kernel, provenance) from the committed ncu --set full summaries, or (None, why)"""
    p = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    src = "from file: profiles/roofline_traffic.json (dram__bytes_read + dram__bytes_write of this kernel in a committed ncu --set full capture); not measured in this run"
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            b = d.get("render_lean_kernel_batch", {})
            if b.get("views_per_launch") == views_per_launch:
                return b.get("dram_bytes_per_launch"), src + "; capture of one %d-view launch" % views_per_launch
            v = d.get("render_lean_kernel", {}).get("dram_bytes_per_launch")
            if v is not None:
                return v * views_per_launch, src + "; capture of a one-view launch x %d views" % views_per_launch
        except Exception:
            pass
    return None, "no committed capture"


def sharding():
    import importlib.util
    spec = importlib.util.spec_from_file_location("lrf_sharding", os.path.join(ROOT, "literf-radiancefields_b200", "sharding.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def cameras(pkg, rank, world, strong=False):
    """weak scaling: a VIEWS*world-view orbit dealt round-robin, VIEWS views per rank; strong: the VIEWS-view orbit dealt round-robin"""
    proj_inv = pkg.inverse4x4(pkg.perspective(45.0, ASPECT, 0.1, 100.0))
    total = VIEWS if strong else VIEWS * world
    views = [pkg.inverse4x4(pkg.orbit_view(k, total)) for k in sharding().shard_views(total, rank, world)]
    return proj_inv, views


# --------------------------------------------------------------------------------------------------
def cpu_reference_rate(grid, steps, warmup, quiet=False):
    """times the reference's own kernel2D_RayMarch (oracle/_ref) — or the oracle port if the reference build is
    absent — on a CPU_SAMPLE^2 frame of view 0 of the same workload.  Returns (rays_per_s list, info)."""
    import libs
    from libs import fp, u32p, np_view
    n = CPU_SAMPLE * CPU_SAMPLE
    out = np.zeros(n, np.uint32)
    rates = []
    try:
        host_cores = len(os.sched_getaffinity(0))
    except Exception:
        host_cores = os.cpu_count() or 1
    if libs.have_ref("O3"):
        R = libs.ref("O3")
        R.ref_set_omp_threads(host_cores)     # torchrun exports OMP_NUM_THREADS=1; the CPU arm runs on all the host cores it can use
        kind, cores = "reference", R.ref_omp_threads()
        h = R.ref_create()
        R.ref_init_grid(h, GS)   # builds the 255^3-box octree (~15 s, not timed: setup, like InitGrid in main.cpp)
        np_view(R.ref_grid_ptr(h), GS ** 3 * 13)[:] = grid.ravel()
        proj, view = np.zeros(16, np.float32), np.zeros(16, np.float32)
        R.ref_make_perspective(45.0, 1.0, 0.1, 100.0, fp(proj))
        R.ref_make_orbit_view(0, VIEWS, fp(view))
        R.ref_set_proj_matrix(h, fp(proj))
        R.ref_set_view_matrix(h, fp(view))
        run = lambda: R.ref_render(h, u32p(out), CPU_SAMPLE, CPU_SAMPLE, 0)
        what = "oracle/_ref/libref_O3.so: reference example_tracer.cpp compiled -O3 -fopenmp -ffp-contract=off"
    else:
        O = libs.oracle()
        O.orc_set_omp_threads(host_cores)
        kind, cores = "port", O.orc_omp_threads()
        h = O.orc_create(GS)
        np_view(O.orc_grid_ptr(h), GS ** 3 * 13)[:] = grid.ravel()
        proj, view, vi, pi = (np.zeros(16, np.float32) for _ in range(4))
        O.orc_perspective(45.0, 1.0, 0.1, 100.0, fp(proj))
        O.orc_orbit_view(0, VIEWS, fp(view))
        O.orc_inverse4x4(fp(view), fp(vi))
        O.orc_inverse4x4(fp(proj), fp(pi))
        O.orc_set_view_inv(h, fp(vi))
        O.orc_set_proj_inv(h, fp(pi))
        run = lambda: O.orc_render(h, u32p(out), CPU_SAMPLE, CPU_SAMPLE, 0)
        what = "oracle/_build/liblrf_oracle.so: CPU restatement (reference build absent)"
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        run()
        dt = time.perf_counter() - t0
        if i >= warmup:
            rates.append(n / dt)
    cpu_reference_rate.last_frame = out.copy()    # the sample frame itself: bench compares the GPU's render of the same frame with it
    info = {"kind": kind, "cores": int(cores),
            "sample": "%dx%d frame of orbit view 0 (%d rays) per step on the %d^3 D=%g grid, %d timed step(s); %s" % (CPU_SAMPLE, CPU_SAMPLE, n, GS, DENSITY_SCALE, steps, what)}
    return rates, info



def cpu_config1_rates():
    """BASELINE.json configs[0], the reference as it runs today (main.cpp:49,67-68,99-102): 128^3 grid, 256x256 frames of the 7-view
    orbit, its own kernel2D_RayMarch — timed for the build the reference ships (-O0, CMakeLists.txt:62) and for -O3, on a bounded
    sample (views 0 and 3 at -O3, view 0 at -O0)."""
    import libs
    from libs import fp, u32p, np_view
    gs, w = 128, 256
    grid = lcg_grid(gs, DENSITY_SCALE)
    try:
        host_cores = len(os.sched_getaffinity(0))
    except Exception:
        host_cores = os.cpu_count() or 1
    out = {"workload": "128^3 synthetic LCG grid (D=%g), 256x256 frames, orbit of 7 views (BASELINE.json configs[0])" % DENSITY_SCALE}
    for kind, views in (("O3", (0, 3)), ("O0", (0,))):
        if not libs.have_ref(kind):
            continue
        R = libs.ref(kind)
        R.ref_set_omp_threads(host_cores)
        h = R.ref_create()
        R.ref_init_grid(h, gs)
        np_view(R.ref_grid_ptr(h), gs ** 3 * 13)[:] = grid.ravel()
        proj, view = np.zeros(16, np.float32), np.zeros(16, np.float32)
        R.ref_make_perspective(45.0, 1.0, 0.1, 100.0, fp(proj))
        R.ref_set_proj_matrix(h, fp(proj))
        img = np.zeros(w * w, np.uint32)
        secs = 0.0
        for k in views:
            R.ref_make_orbit_view(k, 7, fp(view))
            R.ref_set_view_matrix(h, fp(view))
            t0 = time.perf_counter()
            R.ref_render(h, u32p(img), w, w, 0)
            secs += time.perf_counter() - t0
        R.ref_destroy(h)
        out[kind] = {"value": w * w * len(views) / secs / 1e6, "unit": "Mrays/s", "cores": int(R.ref_omp_threads()), "views_timed": list(views),
                     "build": "-O0 -fopenmp (as shipped, CMakeLists.txt:62)" if kind == "O0" else "-O3 -fopenmp -ffp-contract=off"}
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    select_workload("render256")       # the CPU arm is timed on the configuration the metric is quoted on
    grid = lcg_grid(GS, DENSITY_SCALE)
    steps, warmup = max(1, args.steps), max(1, min(args.warmup, 3))
    rates, info = cpu_reference_rate(grid, steps, warmup)
    n = CPU_SAMPLE * CPU_SAMPLE
    total_t = sum(n / r for r in rates)
    value = n * len(rates) / total_t / 1e6
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
        "ms_per_step": 1e3 * total_t / len(rates), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": workload_config(1),
        "cpu_baseline": dict(info, value=value, unit="Mrays/s"),
        "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def workload_config(world, strong=False):
    vpg = (VIEWS + world - 1) // world if strong else VIEWS
    return {"workload": WORKLOADS[WORKLOAD]["text"], "name": WORKLOAD,
            "grid": GS, "width": W, "height": H, "views_per_gpu": vpg, "rays_per_step_per_gpu": W * H * vpg,
            "density_scale": DENSITY_SCALE, "parallelism": "views sharded over %d GPU(s), grid replica per GPU, no collective" % world,
            "l2": "inputs exceed L2: the %d MB grid is re-read by every view (L2 = 126 MB); no explicit flush needed" % (52 * GS ** 3 // 10 ** 6)}


# --------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-train", action="store_true", help="skip the (extra) training-iteration measurement")
    ap.add_argument("--workload", default="render256", choices=sorted(WORKLOADS))
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak (default, the driver's contract): VIEWS views per GPU; strong: the VIEWS views of ONE orbit dealt over the GPUs")
    ap.add_argument("--no-extra-configs", action="store_true", help="skip the short sparse512 / hires512 measurements appended to extra")
    args = ap.parse_args()
    select_workload(args.workload)
    if args.workload != "render256":
        args.no_train = True          # the training extra and the CPU arm belong to the default workload
        args.no_cpu_baseline = True
    if args.impl == "reference":
        return run_reference(args)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    steps, warmup = max(1, args.steps), max(3, args.warmup)
    # the contract is ONE JSON line on stdout: NCCL prints its version banner there, so everything but the final line
    # is sent to stderr
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    import __graft_entry__ as ge
    pkg = ge.package()
    t = pkg.RayMarcherExample(local)          # raises if the CUDA library / a B200 is missing: no fallback
    t.InitGrid(GS)
    t.SetBoundingBox([0, 0, 0], [1, 1, 1])
    grid = None
    if FILL_MODE is None:
        grid = lcg_grid(GS, DENSITY_SCALE)
        t.set_grid(grid)
    else:
        t.fill_synthetic(DENSITY_SCALE, FILL_MODE)
    if SPARSE:
        t.RebuildOctree()              # keeps the boxes whose centre density is > 0 (ref example_tracer.cpp:449-489)
    strong = args.scaling == "strong"
    proj_inv, views = cameras(pkg, rank, world, strong)
    t.set_proj_inv(proj_inv)
    frame = pkg.DeviceBuffer(t, W * H * 4)

    def barrier():
        t.sync()
        if dist is not None:
            dist.barrier()

    def max_over_ranks(x):
        if dist is None:
            return x
        import torch
        v = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(v, op=dist.ReduceOp.MAX)
        return float(v.item())

    # ---- value: device-resident ----
    # one step = the VIEWS views of this GPU's shard through lrf_render_views_dev_async: one launch of the forward kernel over all views
    # (frame v of `frames`), the view matrices travel inside the call
    frames = pkg.DeviceBuffer(t, len(views) * W * H * 4)

    def step_resident():
        t.render_views_dev_async(views, frames.ptr, W, H)

    for _ in range(warmup):
        step_resident()
    barrier()
    sampler = ClockSampler(local)
    time.sleep(0.3)
    l0 = t.launch_count()
    tc0 = time.perf_counter()
    t.timer_start()
    for _ in range(steps):
        step_resident()
    ms = t.timer_stop_ms()
    tc1 = time.perf_counter()
    barrier()
    launches = t.launch_count() - l0
    clocks = sampler.stop(tc0, tc1)
    ms = max_over_ranks(ms)
    rays_per_step = W * H * (VIEWS if strong else VIEWS * world)
    value = rays_per_step * steps / (ms * 1e-3) / 1e6

    # ---- e2e: host API, camera h2d + frame d2h per view inside the timed region ----
    pinned = pkg.PinnedBuffer(len(views) * W * H, np.uint32)
    e2e_steps = max(1, min(steps, 2))

    def step_host():
        # host matrices in, every finished frame out into pinned host memory, synchronous (lrf_render_views: batches of views, a finished
        # batch is copied out while the next one renders)
        t.render_views(views, pinned.array, W, H)

    step_host()
    barrier()
    te0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_host()
    t.sync()
    e2e_s = max_over_ranks(time.perf_counter() - te0)
    checksum = int(pinned.array[-W * H:].astype(np.uint64).sum())      # the last view's frame, as before the batch call existed
    e2e_value = rays_per_step * e2e_steps / e2e_s / 1e6
    # the batch frames are the frames the per-view call (kernel2D_RayMarch == lrf_render) delivers: checked on three views, outside the timed regions
    # the same step view by view through the reference's own entry point (kernel2D_RayMarch == lrf_render: set the view, render, frame to
    # pinned host memory, synchronous, once per view), timed once, for comparison with the batch call's e2e
    one_pinned = pkg.PinnedBuffer(W * H, np.uint32)
    barrier()
    tv0 = time.perf_counter()
    for vi in views:
        t.set_view_inv(vi)
        t.kernel2D_RayMarch(one_pinned.array, W, H)
    t.sync()
    e2e_per_view_value = rays_per_step / max_over_ranks(time.perf_counter() - tv0) / 1e6
    one = np.zeros(W * H, np.uint32)
    batch_ok = True
    for k in sorted({0, len(views) // 2, len(views) - 1}):
        t.set_view_inv(views[k])
        t.kernel2D_RayMarch(one, W, H)
        batch_ok = batch_ok and bool(np.array_equal(one, pinned.array[k * W * H:(k + 1) * W * H]))

    # ---- roofline of the dominant kernel ----
    sel = list(range(0, VIEWS, 10 if VIEWS >= 20 else 2))
    touched = []
    for k in sel:
        t.set_view_inv(views[k])
        touched.append(t.count_touched_nodes(W, H))
    n_touched = float(np.mean(touched))
    alg_bytes_view = 52.0 * n_touched + 4.0 * W * H
    alg_bytes = alg_bytes_view * len(views)        # ONE launch renders the step's views
    kernel_ms = ms / steps                         # the step is that launch (plus the 3 us fallback kernel): mean launch duration
    peak, peak_src = measured_peak()
    achieved = alg_bytes / (kernel_ms * 1e-3) / 1e9
    traffic, traffic_src = ncu_traffic(len(views))
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "kernel": "render_image_kernel<SPARSE>" if SPARSE else "render_lean_kernel", "kernel_ms": kernel_ms, "views_per_launch": len(views), "kernel_ms_per_view": kernel_ms / len(views),
                "algorithmic_bytes_per_launch": alg_bytes, "algorithmic_bytes_per_view": alg_bytes_view,
                "n_touched_nodes_mean": n_touched, "n_touched_sampled_views": len(sel), "peak_source": peak_src,
                "traffic_source": traffic_src}
    if WORKLOAD != "render256":
        roofline["traffic"] = None     # the committed ncu capture is of the default workload

    extra = {"e2e_checksum": checksum, "batch_frames_equal_per_view_frames": batch_ok,
             "e2e_view_by_view": {"value": e2e_per_view_value, "unit": "Mrays/s", "call": "kernel2D_RayMarch (lrf_render) once per view, pinned host frame", "steps": 1}}
    if SPARSE:
        extra["boxes"] = int(t.box_count())
        extra["occupancy"] = t.box_count() / float((GS - 1) ** 3)
    if WORKLOAD == "render256":
        try:
            extra["latency_mode"] = latency_bench(pkg, t, frame, rank, world, dist, max_over_ranks, barrier)
        except Exception as e:
            extra["latency_mode"] = {"error": repr(e)}
    if not args.no_train:
        try:
            extra["train"] = train_bench(pkg, t, grid, views, proj_inv, rank, world, dist, max_over_ranks, barrier)
        except Exception as e:  # the headline line must still be printed
            extra["train"] = {"error": repr(e)}
    if WORKLOAD == "render256" and not args.no_extra_configs:
        try:
            extra["configs"] = other_configs(pkg, t, local, rank, world, max_over_ranks, barrier)
        except Exception as e:
            extra["configs"] = {"error": repr(e)}

    line = {
        "metric": METRIC, "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": steps, "warmup": warmup,
        "ms_per_step": ms / steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(world, strong), "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": len(views) * 2 * 64, "d2h_bytes_per_step": len(views) * W * H * 4, "steps": e2e_steps},
        "gpu_launches": int(launches), "roofline": roofline, "extra": extra,
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            rates, info = cpu_reference_rate(grid, 2, 1)
            line["cpu_baseline"] = dict(info, value=float(np.mean(rates)) / 1e6, unit="Mrays/s")
            # the same CPU_SAMPLE^2 frame through the CUDA path, compared with what the CPU code just produced (full-size grid)
            mine = np.zeros(CPU_SAMPLE * CPU_SAMPLE, np.uint32)
            t.set_grid(grid)                      # the training extra above has moved the parameters: back to the CPU arm's grid
            t.set_view_inv(views[0])
            t.kernel2D_RayMarch(mine, CPU_SAMPLE, CPU_SAMPLE)
            theirs = cpu_reference_rate.last_frame
            ch = lambda f: np.stack([(f >> sft) & 255 for sft in (0, 8, 16, 24)]).astype(np.int32)
            dlt = np.abs(ch(mine) - ch(theirs))
            line["cpu_baseline"]["frame_check"] = {"pixels": int(mine.size), "max_channel_diff_of_255": int(dlt.max()),
                                                   "channels_differing": int((dlt != 0).sum()),
                                                   "note": "GPU render of the sample frame vs the CPU arm's own output; RGB tolerance 1e-4 allows +-1 of 255"}
        except Exception as e:
            line["cpu_baseline"] = {"error": repr(e)}
        try:
            line["cpu_baseline"]["config1"] = cpu_config1_rates()
        except Exception as e:
            line["cpu_baseline"]["config1"] = {"error": repr(e)}
    sys.stdout.flush()
    os.dup2(real_stdout, 1)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if dist is not None:
        os.dup2(2, 1)
        dist.destroy_process_group()
    return 0



def latency_bench(pkg, t, frame, rank, world, dist, max_over_ranks, barrier, reps=10):
    """strong scaling of the render (SURVEY 8e latency mode): (a) ONE 800x800 frame split over the GPUs in row bands (the reference
    splits the rows of a frame over its OpenMP threads, example_tracer.cpp:630), bands gathered on rank 0 — the only collective;
    (b) the VIEWS views of one orbit dealt over the GPUs.  Times are max over ranks, CUDA events around the whole loop."""
    shard = sharding()
    view = pkg.inverse4x4(pkg.orbit_view(7, 100))
    t.set_view_inv(view)
    y0, y1 = shard.shard_rows(H, rank, world)
    ftensor = None
    if dist is not None:
        import torch

        class _Arr:
            __cuda_array_interface__ = {"shape": (W * H,), "typestr": "<i4", "data": (frame.ptr, False), "version": 2}
        ftensor = torch.as_tensor(_Arr(), device="cuda")
        ext = torch.cuda.ExternalStream(t.lib.lrf_stream(t.h))

    def one_frame():
        t.render_rows_dev_async(frame.ptr, W, H, y0, y1)
        if ftensor is not None:
            import torch
            with torch.cuda.stream(ext):
                shard.gather_frame(ftensor, H, W, dist)

    for _ in range(3):
        one_frame()
    barrier()
    t.timer_start()
    for _ in range(reps):
        one_frame()
    ms_frame = max_over_ranks(t.timer_stop_ms()) / reps
    barrier()
    # check on rank 0: the assembled frame equals the frame one GPU renders alone
    ok = None
    if rank == 0:
        host = np.zeros(W * H, np.uint32)
        frame.download(host)
        alone = np.zeros(W * H, np.uint32)
        t.kernel2D_RayMarch(alone, W, H)
        ok = bool(np.array_equal(host, alone))
    # (b) one orbit of VIEWS views over all GPUs
    mine = [pkg.inverse4x4(pkg.orbit_view(k, VIEWS)) for k in shard.shard_views(VIEWS, rank, world)]
    for vi in mine[:3]:
        t.set_view_inv(vi)
        t.render_dev_async(frame.ptr, W, H)
    barrier()
    t.timer_start()
    for vi in mine:
        t.set_view_inv(vi)
        t.render_dev_async(frame.ptr, W, H)
    ms_orbit = max_over_ranks(t.timer_stop_ms())
    barrier()
    return {"scaling": "strong", "n_gpus": world,
            "one_frame_ms": ms_frame, "one_frame_Mrays_per_s": W * H / ms_frame / 1e3, "one_frame_rows_per_gpu": y1 - y0,
            "one_frame_collective": "row bands gathered on rank 0 (NCCL send/recv of %d-row bands)" % (y1 - y0) if world > 1 else "none (1 GPU)",
            "one_frame_equals_single_gpu_frame": ok,
            "orbit_views": VIEWS, "orbit_ms": ms_orbit, "orbit_Mrays_per_s": W * H * VIEWS / ms_orbit / 1e3}


def other_configs(pkg, t0, local, rank, world, max_over_ranks, barrier):
    """the larger configurations of BASELINE.json, short measurements in the default run so that the driver sees them: configs[4]
    (dense 512^3 at 1920x1080 and 3840x2160, 8 views per GPU, view-sharded) and configs[3] (sparse 512^3, ~10 % occupancy, 800x800).
    value = aggregate Mrays/s over the GPUs, device-timed, max over ranks."""
    out = {}
    t = pkg.RayMarcherExample(local)
    gs = 512
    t.InitGrid(gs)
    t.SetBoundingBox([0, 0, 0], [1, 1, 1])
    shard = sharding()

    def run(name, w, h, aspect, n_views, reps=2):
        proj_inv = pkg.inverse4x4(pkg.perspective(45.0, aspect, 0.1, 100.0))
        views = [pkg.inverse4x4(pkg.orbit_view(k, n_views * world)) for k in shard.shard_views(n_views * world, rank, world)]
        t.set_proj_inv(proj_inv)
        fr = pkg.DeviceBuffer(t, len(views) * w * h * 4)     # one frame per view: the batch call (one launch for dense sets, view by view for sparse ones)
        t.render_views_dev_async(views, fr.ptr, w, h)          # warm-up with the timed call's own shape: its buffers are sized here, not inside the timed region
        barrier()
        t.sync()
        t.timer_start()
        for _ in range(reps):
            t.render_views_dev_async(views, fr.ptr, w, h)
        ms = max_over_ranks(t.timer_stop_ms())
        barrier()
        fr.free()
        out[name] = {"value": w * h * n_views * world * reps / ms / 1e3, "unit": "Mrays/s", "ms_per_view": ms / (reps * n_views), "width": w, "height": h,
                     "views_per_gpu": n_views, "n_gpus": world, "grid": gs, "workload": WORKLOADS[name]["text"]}

    t.fill_synthetic(20.0, 0)
    run("hires512_1080p", 1920, 1080, 16.0 / 9.0, 8)
    run("hires512_4k", 3840, 2160, 16.0 / 9.0, 8)
    t.fill_synthetic(20.0, 2)
    t.RebuildOctree()
    run("sparse512", 800, 800, 1.0, 8)
    out["sparse512"]["boxes"] = int(t.box_count())
    out["sparse512"]["occupancy"] = t.box_count() / float((gs - 1) ** 3)
    t.close()
    return out


def exchange_check(pkg, t, grid, dist, torch, shard, one_nccl, one_peer_factory, n_it=4):
    """the three exchange paths from the same state: replicas identical within a path (all-gathered checksum of the parameter bits),
    and each fused path against the NCCL path (max |difference| over the parameter planes of this rank, max over ranks)"""
    n = t.lib.lrf_param_count(t.h)

    def params_tensor():
        class _P:
            __cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (t.lib.lrf_param_device_ptr(t.h), False), "version": 2}
        return torch.as_tensor(_P(), device="cuda")

    def reset():
        if FILL_MODE is None:
            t.set_grid(grid)
        else:
            t.fill_synthetic(DENSITY_SCALE, FILL_MODE)
        t.optimizerInit()
        t.zeroGrad()
        t.sync()

    def checksum():
        t.sync()
        p = params_tensor()
        bits = p.view(torch.int32).to(torch.int64)
        cs = int((bits * (torch.arange(n, device="cuda", dtype=torch.int64) % 1000003 + 1)).sum().item())
        alls = [None] * dist.get_world_size()
        dist.all_gather_object(alls, cs)
        return p, len(set(alls)) == 1

    out = {}
    reset()
    for it in range(n_it):
        one_nccl(it)
    p, same = checksum()
    ref = p.clone()
    out["nccl_replicas_identical"] = same
    for name, cls in (("peer", shard.PeerExchange), ("multicast", shard.McExchange)):
        if name == "multicast" and not t.mc_supported():
            continue
        try:
            reset()
            e = cls(t, dist, torch)
            step = one_peer_factory(e)
            for it in range(n_it):
                step(it)
            p, same = checksum()
            d = torch.tensor([float((p - ref).abs().max().item())], device="cuda")
            dist.all_reduce(d, op=dist.ReduceOp.MAX)
            out[name + "_replicas_identical"] = same
            out[name + "_max_abs_diff_vs_nccl"] = float(d.item())
            e.close()
        except Exception as ex_:
            out[name + "_error"] = "%s: %s" % (type(ex_).__name__, ex_)
    out["note"] = "%d iterations from the same parameters through each path; the NCCL ring and the switch add in different orders than the rank-order sum (last-bit differences in the gradient sum)" % n_it
    return out


def train_bench(pkg, t, grid, views, proj_inv, rank, world, dist, max_over_ranks, barrier, iters=20, batch=4096):
    """BASELINE.json configs[2]: grid optimisation on the 256^3 grid, 4096-ray batches per GPU against synthetic targets:
    one iteration = ray generation + backward(4096 rays) + per-cell pass ({TV -> Adam(density)} || Adam(colour), zeroing
    fused) on one GPU; on N GPUs either
    backward + NCCL all-reduce of the 13*N-float gradient + replicated TV + Adam + zero, or backward || slab TV + the
    peer-memory fused exchange (barrier, slab {sum over ranks + Adam -> every replica}, barrier, zero)."""
    import ctypes as C
    n_views = len(views)
    rng = np.random.default_rng(1234 + rank)
    vi_dev = pkg.DeviceBuffer(t, n_views * 64)
    vi_dev.upload(np.stack(views).astype(np.float32))
    picks = np.stack([rng.integers(0, n_views, (iters + 3) * batch), rng.integers(0, W * H, (iters + 3) * batch)], axis=1).astype(np.uint32)
    picks_dev = pkg.DeviceBuffer(t, picks.nbytes)
    picks_dev.upload(picks)
    ref = rng.random(((iters + 3) * batch, 4)).astype(np.float32)   # synthetic target colours
    ref_dev = pkg.DeviceBuffer(t, ref.nbytes)
    ref_dev.upload(ref)
    o_dev, d_dev = pkg.DeviceBuffer(t, batch * 12), pkg.DeviceBuffer(t, batch * 12)
    t.optimizerInit()
    t.zeroGrad()
    grad_tensor = None
    if dist is not None:
        import torch
        n = t.lib.lrf_param_count(t.h)

        def grad_view():   # the planes move when a multicast group is joined or left: never keep the view across such a call
            class _Arr:
                __cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (t.lib.lrf_grad_device_ptr(t.h), False), "version": 2}
            return torch.as_tensor(_Arr(), device="cuda")
        grad_tensor = grad_view()
        ext = torch.cuda.ExternalStream(t.lib.lrf_stream(t.h))

    shard = sharding()

    def one(it, exchange=True):
        off = it * batch
        t._ck(t.lib.lrf_gen_rays_dev_async(t.h, batch, vi_dev.ptr, picks_dev.ptr + off * 8, W, H, o_dev.ptr, d_dev.ptr))
        t._ck(t.lib.lrf_rays_grad_dev_async(t.h, batch, o_dev.ptr, d_dev.ptr, ref_dev.ptr + off * 16, 0))
        if grad_tensor is not None and exchange:
            import torch
            with torch.cuda.stream(ext):
                shard.allreduce_gradient(grad_view(), dist)
        t._ck(t.lib.lrf_train_percell_async(t.h, it))

    for it in range(3):
        one(it)
    barrier()
    t.timer_start()
    for it in range(3, 3 + iters):
        one(it)
    ms = max_over_ranks(t.timer_stop_ms())
    barrier()
    ms_noex = None
    ms_peer = ms_peer_exchange = ms_barrier = None
    ms_mc = ms_mc_exchange = mc_error = replica_check = None
    if grad_tensor is not None:   # same iterations without the exchange step: the all-reduce share of an iteration
        t.timer_start()
        for it in range(3, 3 + iters):
            one(it, exchange=False)
        ms_noex = max_over_ranks(t.timer_stop_ms())
        barrier()
        # the fused exchange over NVLink peer memory: barrier, own slab {TV + sum of all ranks' gradients + Adam, parameters
        # stored into every replica}, barrier, gradient zeroing (sharding.PeerExchange / lrf_train_percell_peer_async)
        import torch
        t.optimizerInit()
        t.zeroGrad()
        ex = shard.PeerExchange(t, dist, torch)

        def one_peer(it):
            off = it * batch
            t._ck(t.lib.lrf_gen_rays_dev_async(t.h, batch, vi_dev.ptr, picks_dev.ptr + off * 8, W, H, o_dev.ptr, d_dev.ptr))
            t._ck(t.lib.lrf_rays_grad_tv_dev_async(t.h, batch, o_dev.ptr, d_dev.ptr, ref_dev.ptr + off * 16, 0))   # backward || TV of the own slab
            ex.percell(it, tv_done=True)

        for it in range(3):
            one_peer(it)
        barrier()
        t.timer_start()
        for it in range(3, 3 + iters):
            one_peer(it)
        ms_peer = max_over_ranks(t.timer_stop_ms())
        barrier()
        # the exchange alone (no backward): barrier + fused pass + barrier + zero
        t.timer_start()
        for it in range(5):
            ex.percell(200 + it, tv_done=True)
        ms_peer_exchange = max_over_ranks(t.timer_stop_ms()) / 5
        t.timer_start()
        for it in range(5):
            ex.barrier()
        ms_barrier = max_over_ranks(t.timer_stop_ms()) / 5
        barrier()
        ex.close()
        # the same exchange through the NVSwitch (NVLink multicast objects: multimem.ld_reduce / multimem.st, sharding.McExchange)
        if t.mc_supported():
            try:
                exm = shard.McExchange(t, dist, torch)
                ex = exm

                for it in range(3):
                    one_peer(it)
                barrier()
                t.timer_start()
                for it in range(3, 3 + iters):
                    one_peer(it)
                ms_mc = max_over_ranks(t.timer_stop_ms())
                barrier()
                t.timer_start()
                for it in range(5):
                    exm.percell(300 + it, tv_done=True)
                ms_mc_exchange = max_over_ranks(t.timer_stop_ms()) / 5
                barrier()
                exm.close()
            except Exception as e:   # reported, not fatal: the point-to-point exchange above stands
                mc_error = "%s: %s" % (type(e).__name__, e)
        # replicas and paths agree: the same 4 iterations from the same state through each path; per path the parameter planes of all
        # ranks are compared (bit pattern checksum all-gathered), and the paths against the NCCL one (max |difference|)
        replica_check = exchange_check(pkg, t, grid, dist, torch, shard, one, one_peer_factory=lambda e: (lambda it: (
            t._ck(t.lib.lrf_gen_rays_dev_async(t.h, batch, vi_dev.ptr, picks_dev.ptr + it * batch * 8, W, H, o_dev.ptr, d_dev.ptr)),
            t._ck(t.lib.lrf_rays_grad_tv_dev_async(t.h, batch, o_dev.ptr, d_dev.ptr, ref_dev.ptr + it * batch * 16, 0)),
            e.percell(it, tv_done=True))))
    # break-down: backward alone / per-cell alone
    t.timer_start()
    for it in range(5):
        t._ck(t.lib.lrf_rays_grad_dev_async(t.h, batch, o_dev.ptr, d_dev.ptr, ref_dev.ptr, 0))
    ms_bwd = t.timer_stop_ms() / 5
    t.timer_start()
    for it in range(5):
        t._ck(t.lib.lrf_train_percell_async(t.h, 100 + it))
    ms_cell = t.timer_stop_ms() / 5
    N = GS ** 3
    peak, _ = measured_peak()
    # algorithmic bytes of one iteration: per-cell pass 428 B/node (dense, parity semantics) + the backward's
    # node traffic (52 B read + 2 x 52 B gradient read-modify-write per touched node; counted as touched-node upper bound
    # rays x samples x 8 is NOT used — only the dense per-cell term, which dominates, enters the roofline figure)
    iter_bytes = 428.0 * N
    cands = [(ms, "NCCL all-reduce + replicated per-cell pass")]
    if ms_peer:
        cands.append((ms_peer, "peer-memory fused, point to point (lrf_train_percell_peer_async, adam_peer_kernel)"))
    if ms_mc:
        cands.append((ms_mc, "NVLink multicast fused: multimem.ld_reduce + multimem.st through the switch (lrf_mc_*, adam_mc_kernel)"))
    best, best_name = min(cands)
    out = {"metric": "train_iters_per_s", "value": iters / (best * 1e-3), "ms_per_iter": best / iters, "rays_per_iter": batch * world, "n_gpus": world,
           "exchange": best_name if world > 1 else "none (1 GPU)",
           "ms_per_iter_nccl_allreduce": ms / iters, "ms_per_iter_peer_fused": (ms_peer / iters) if ms_peer else None,
           "ms_per_iter_multicast_fused": (ms_mc / iters) if ms_mc else None, "ms_multicast_exchange_alone": ms_mc_exchange, "multicast_error": mc_error,
           "iter_frac_of_hbm_peak_multicast_fused": ((428.0 * GS ** 3 / (ms_mc / iters * 1e-3) / 1e9) / measured_peak()[0]) if ms_mc else None,
           "replica_check": replica_check,
           "ms_peer_exchange_alone": ms_peer_exchange, "ms_stream_barrier": ms_barrier,
           "scaling": "weak (4096 rays per GPU per iteration)",
           "ms_backward_4096_rays": ms_bwd, "ms_percell_tv_adam_zero": ms_cell,
           "percell_algorithmic_bytes": 428.0 * N, "percell_frac_of_hbm_peak": (428.0 * N / (ms_cell * 1e-3) / 1e9) / peak,
           "iter_frac_of_hbm_peak_incl_allreduce": (iter_bytes / (ms / iters * 1e-3) / 1e9) / peak,
           "iter_frac_of_hbm_peak_peer_fused": ((iter_bytes / (ms_peer / iters * 1e-3) / 1e9) / peak) if ms_peer else None,
           "ms_per_iter_without_allreduce": (ms_noex / iters) if ms_noex else ms / iters,
           "allreduce_ms_per_iter": ((ms - ms_noex) / iters) if ms_noex else 0.0,
           "allreduce": "NCCL all_reduce (sum) of the 13*N fp32 gradient planes (%d MB) per iteration, on the context stream" % (52 * N // 2 ** 20) if world > 1 else "none (1 GPU)"}
    for b in (vi_dev, picks_dev, ref_dev, o_dev, d_dev):
        b.free()
    return out


if __name__ == "__main__":
    sys.exit(main())
