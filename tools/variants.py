"""Kernel-variant experiments for the forward march.
    python tools/variants.py build  name=DEF1,DEF2 ...     (here; parallel nvcc, output under _build/variants/)
    python tools/variants.py run [n_views]                 (on the GPU box; times every built variant + the default library)
Each run renders the bench workload's views (256^3 D=20, 800x800), prints ms/view (CUDA events on the context stream) and a
checksum of the frames so that variants can be checked for identical output."""
import concurrent.futures as cf
import glob
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VDIR = os.path.join(ROOT, "literf-radiancefields_b200", "_build", "variants")


def build(specs):
    sys.path.insert(0, os.path.join(ROOT, "literf-radiancefields_b200"))
    import build as b
    with cf.ThreadPoolExecutor(max_workers=6) as ex:
        futs = {}
        for spec in specs:
            name, _, defs = spec.partition("=")
            futs[ex.submit(b.build_variant, name, [d for d in defs.split(",") if d])] = name
        for f in cf.as_completed(futs):
            lib = f.result()
            log = open(lib + ".log").read()
            info = [l for l in log.splitlines() if "render_lean_kernelILb0ELb1" in l]
            # the ptxas line after "Compiling entry function ... render_image_kernel<0,0>" holds registers / spills
            lines = log.splitlines()
            used = ""
            for i, l in enumerate(lines):
                if "Compiling entry function" in l and "render_lean_kernelILb0ELb1" in l:
                    used = " | ".join(x.strip() for x in lines[i + 1:i + 4])
                    break
            print(futs[f], "->", used)


def time_one(n_views):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import numpy as np
    import __graft_entry__ as ge
    import bench
    pkg = ge.package()
    if os.environ.get("LRF_VARIANTS_WORKLOAD"):     # e.g. sparse512, hires512_1080p (bench.py's workload table)
        bench.select_workload(os.environ["LRF_VARIANTS_WORKLOAD"])
    t = pkg.RayMarcherExample(0)
    t.InitGrid(bench.GS)
    if bench.FILL_MODE is None:
        t.set_grid(bench.lcg_grid(bench.GS, bench.DENSITY_SCALE))
    else:
        t.fill_synthetic(bench.DENSITY_SCALE, bench.FILL_MODE)
    if bench.SPARSE:
        t.RebuildOctree()
    proj_inv, views = bench.cameras(pkg, 0, 1)
    t.set_proj_inv(proj_inv)
    frame = pkg.DeviceBuffer(t, bench.W * bench.H * 4)
    host = np.zeros(bench.W * bench.H, np.uint32)
    pick = [(k * 13) % len(views) for k in range(n_views)]
    for k in pick[:3]:
        t.set_view_inv(views[k])
        t.render_dev_async(frame.ptr, bench.W, bench.H)
    t.sync()
    ck = 0
    t.timer_start()
    for k in pick:
        t.set_view_inv(views[k])
        t.render_dev_async(frame.ptr, bench.W, bench.H)
    ms = t.timer_stop_ms()
    for k in pick[:4]:
        t.set_view_inv(views[k])
        t.render_dev_async(frame.ptr, bench.W, bench.H)
        frame.download(host)
        ck = (ck * 1000003 + int(host.astype(np.uint64).sum())) % (1 << 61)
    out = {"lib": os.path.basename(os.environ.get("LRF_B200_LIB", "default")), "ms_per_view": ms / len(pick), "Mrays_s": bench.W * bench.H * len(pick) / ms / 1e3,
           "checksum": ck}
    if os.environ.get("LRF_VARIANTS_BATCH"):   # the bench's step: all views in one launch (lrf_render_views_dev_async)
        frames = pkg.DeviceBuffer(t, len(views) * bench.W * bench.H * 4)
        t.render_views_dev_async(views, frames.ptr, bench.W, bench.H)
        t.sync()
        t.timer_start()
        for _ in range(2):
            t.render_views_dev_async(views, frames.ptr, bench.W, bench.H)
        out["batch_ms_per_view"] = t.timer_stop_ms() / (2 * len(views))
        out["batch_Mrays_s"] = bench.W * bench.H / out["batch_ms_per_view"] / 1e3
    if os.environ.get("LRF_VARIANTS_TRAIN"):   # packed-batch backward: 4096 rays, CUDA events
        rng = np.random.default_rng(1)
        n = 4096
        vi_dev = pkg.DeviceBuffer(t, len(views) * 64)
        vi_dev.upload(np.stack(views).astype(np.float32))
        picks = np.stack([rng.integers(0, len(views), n), rng.integers(0, bench.W * bench.H, n)], axis=1).astype(np.uint32)
        pk, rf = pkg.DeviceBuffer(t, picks.nbytes), pkg.DeviceBuffer(t, n * 16)
        pk.upload(picks)
        rf.upload(rng.random((n, 4)).astype(np.float32))
        od, dd = pkg.DeviceBuffer(t, n * 12), pkg.DeviceBuffer(t, n * 12)
        t.zeroGrad()
        t._ck(t.lib.lrf_gen_rays_dev_async(t.h, n, vi_dev.ptr, pk.ptr, bench.W, bench.H, od.ptr, dd.ptr))
        for _ in range(2):
            t._ck(t.lib.lrf_rays_grad_dev_async(t.h, n, od.ptr, dd.ptr, rf.ptr, 0))
        t.sync()
        t.timer_start()
        for _ in range(10):
            t._ck(t.lib.lrf_rays_grad_dev_async(t.h, n, od.ptr, dd.ptr, rf.ptr, 0))
        out["ms_backward_4096"] = t.timer_stop_ms() / 10
        out["grad_abs_sum"] = float(np.abs(t.get_grad().astype(np.float64)).sum())
        if os.environ.get("LRF_VARIANTS_SORT"):   # experiment: longest rays first (chord through the unit cube x L1 norm of the direction ~ cells crossed)
            o = np.zeros(n * 3, np.float32); d = np.zeros(n * 3, np.float32); r = np.zeros(n * 4, np.float32)
            od.download(o); dd.download(d); rf.download(r)
            o = o.reshape(n, 3); d = d.reshape(n, 3); r = r.reshape(n, 4)
            with np.errstate(divide="ignore", invalid="ignore"):
                t0 = (0.0 - o) / d; t1 = (1.0 - o) / d
            tn = np.nanmax(np.minimum(t0, t1), axis=1); tx = np.nanmin(np.maximum(t0, t1), axis=1)
            key = np.clip(tx - np.maximum(tn, 0.0), 0.0, None) * np.abs(d).sum(axis=1)
            for name, order in (("desc", np.argsort(-key, kind="stable")), ("asc", np.argsort(key, kind="stable"))):
                od.upload(np.ascontiguousarray(o[order])); dd.upload(np.ascontiguousarray(d[order])); rf.upload(np.ascontiguousarray(r[order]))
                t.zeroGrad()
                t._ck(t.lib.lrf_rays_grad_dev_async(t.h, n, od.ptr, dd.ptr, rf.ptr, 0))
                t.sync()
                t.timer_start()
                for _ in range(10):
                    t._ck(t.lib.lrf_rays_grad_dev_async(t.h, n, od.ptr, dd.ptr, rf.ptr, 0))
                out["ms_backward_4096_sorted_" + name] = t.timer_stop_ms() / 10
    if os.environ.get("LRF_VARIANTS_IMGGRAD"):   # whole-image backward, 800x800
        t.zeroGrad()
        t.targets_resize(1, bench.W, bench.H)
        t.target_upload(0, np.random.default_rng(0).random((bench.W * bench.H, 4)).astype(np.float32))
        t.set_view_inv(views[7])
        t.render_grad_target_async(0); t.sync()
        t.timer_start()
        for _ in range(3):
            t.render_grad_target_async(0)
        out["ms_image_backward"] = t.timer_stop_ms() / 3
        out["grad_abs_sum"] = float(np.abs(t.get_grad().astype(np.float64)).sum())
    print(json.dumps(out))


if __name__ == "__main__":
    if sys.argv[1] == "build":
        build(sys.argv[2:])
    elif sys.argv[1] == "one":
        time_one(int(sys.argv[2]))
    else:
        n = sys.argv[2] if len(sys.argv) > 2 else "30"
        libs = [None] + sorted(glob.glob(os.path.join(VDIR, "*.so"))) + [None]   # the default library first AND last: the first process on a fresh box runs ~2 % slow
        for lib in libs:
            env = dict(os.environ)
            if lib:
                env["LRF_B200_LIB"] = lib
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "one", n], env=env, capture_output=True, text=True)
            print(r.stdout.strip().splitlines()[-1] if r.stdout.strip() else "FAILED " + str(lib) + r.stderr[-400:])
