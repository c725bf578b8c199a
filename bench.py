#!/usr/bin/env python
"""bench.py — headline benchmark of the B200 wavefront path tracer.

Metric (BASELINE.json): Mrays/s over all bounces (+ spp/s).  Headline workload = configs[1]: the
bundled OBJ triangle-mesh scene (assets/scenes/bunny.json with a procedurally generated bunny.obj —
the reference's models are git-LFS stubs), 1920x1080, 64 spp, max depth 8.  The same JSON line
carries a `secondary` list with the other BASELINE configs: the 10 M-triangle terrain at 4K /
16 spp (configs[3]), the 800x800 sphere frame (configs[0]) and the 1 spp + 5 x A-Trous 1080p
interactive frame (configs[2]), each with its own e2e, roofline fraction and the reference's number.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
                  [--workload bunny|bunny_1m|three_balls|terrain] [--no-secondary]

A "step" is one render of the frame at the workload's spp (all bounces of every wavefront pass).
`value` is timed with CUDA events on the launching stream with the scene resident in HBM; `e2e`
goes through the public PathTracer API from a host camera struct to a host RGBA8 image (tonemap +
D2H inside the timed region).

N > 1: one process per GPU (torchrun), STRONG scaling: the frame's spp iterations are split into N
sample ranges (rank r renders its share with the seeds a single GPU would use), then one NCCL
reduce of the radiance/G-buffer sums to rank 0; the line reports `image_rmse_vs_single`, the RMSE
between the reduced N-GPU frame and the same frame rendered by rank 0 alone.

`--impl reference` times the reference's own CUDA build (oracle/_ref/libref_cuda_stock.so: compiled
from /root/reference for sm_100 with the reference's own 24-entry traversal stack; the stack-64
build only where the reference's tree is deeper than its stack allows) on the same configs; if no
reference build or device is available, the CPU oracle port.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

WORKLOADS = {
    # name: builder kwargs, width, height, spp, max depth
    "bunny": dict(kind="bunny", subdiv=4, w=1920, h=1080, spp=64, depth=8),
    "bunny_82k": dict(kind="bunny", subdiv=6, w=1920, h=1080, spp=64, depth=8),
    "bunny_1m": dict(kind="bunny", subdiv=8, w=1920, h=1080, spp=64, depth=8),
    "three_balls": dict(kind="balls", w=800, h=800, spp=1, depth=5),
    "terrain": dict(kind="terrain", n=2236, w=3840, h=2160, spp=16, depth=8),
    "terrain_small": dict(kind="terrain", n=700, w=3840, h=2160, spp=16, depth=8),
    "many_materials": dict(kind="materials", w=1920, h=1080, spp=16, depth=8),
    "many_spheres": dict(kind="spheres", n=2000, w=1920, h=1080, spp=8, depth=8),
}
L2_NOTE = "no flush: the per-step path state (hundreds of MB to GB) exceeds the 126 MB L2"
METRIC = "Mrays/s (all bounces)"
REF_PATCHES = ["run-time max_bounces (reference: compile-time 50)", "host-side ray counter",
               "traversal stack size is a build parameter (24 = reference; 64 for trees deeper than 23)"]


def config_of(name):
    wl = WORKLOADS[name]
    return {"workload": name, **{k: v for k, v in wl.items() if k != "kind"}, "l2": L2_NOTE}


def make_scene(wl):
    import cuda_path_tracer_b200 as pt
    if wl["kind"] == "bunny":
        return pt.bunny_scene(pt.bunny_like(wl["subdiv"]), wl["w"], wl["h"], wl["spp"])
    if wl["kind"] == "balls":
        return pt.three_balls(wl["w"], wl["h"], wl["spp"])
    if wl["kind"] == "materials":
        return pt.many_materials_scene(wl["w"], wl["h"], wl["spp"])
    if wl["kind"] == "spheres":
        return pt.many_spheres_scene(wl["n"], wl["w"], wl["h"], wl["spp"])
    return pt.terrain_scene(wl["n"], wl["w"], wl["h"], wl["spp"])


class ClockSampler:
    """Samples SM clocks and throttle reasons during the timed region (NVML)."""

    def __init__(self, device_index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {
            nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
            nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.05)

    def start(self):
        if self.nv:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()

    def stop(self):
        self._stop.set()
        if self._thread:
            self._thread.join()
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(workload):
    """DRAM bytes per traverse_kernel launch from the committed ncu capture of this command
    (profiles/*traverse_traffic*.json, newest round first); None for workloads not captured."""
    pdir = os.path.join(ROOT, "profiles")
    try:
        names = sorted((n for n in os.listdir(pdir) if "traverse_traffic" in n and n.endswith(".json")), reverse=True)
    except OSError:
        return None
    for n in names:
        try:
            with open(os.path.join(pdir, n)) as f:
                j = json.load(f)
            if j.get("workload") == workload:
                return float(j["dram_bytes_per_launch"])
        except Exception:
            continue
    return None


def cpu_baseline(sd, wl):
    """The oracle's CPU port (reference megakernel semantics, OpenMP over pixels) on a bounded
    sample of the workload: the same scene and depth at quarter resolution, as many spp as fit
    in about 12 s."""
    from tests.oracle_lib import load_oracle
    o = load_oracle()
    w, h = max(16, wl["w"] // 4), max(16, wl["h"] // 4)
    osc = o.scene(sd)
    t0 = time.perf_counter()
    _, _, _, rays = osc.render(sd.camera, w, h, 1, wl["depth"])   # probe: sizes the sample
    probe = max(time.perf_counter() - t0, 1e-4)
    spp = int(min(4096, max(1, round(12.0 / probe))))               # ~12 s of CPU work
    t0 = time.perf_counter()
    _, _, _, rays = osc.render(sd.camera, w, h, spp, wl["depth"])
    dt = time.perf_counter() - t0
    return {"value": rays / dt * 1e-6, "unit": "Mrays/s", "cores": o.num_threads(), "kind": "port",
            "sample": f"{w}x{h}, {spp} spp, depth {wl['depth']}, {rays} rays in {dt:.2f} s "
                      f"(oracle/liboracle.so, reference algorithm restated in C, OpenMP over pixels)"}


# ------------------------------------------------------------------------------ distributed
class Dist:
    def __init__(self):
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.multi = self.world > 1
        self.dist = None

    def init(self):
        import torch
        torch.cuda.set_device(self.local_rank)
        if self.multi:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
            self.dist = dist

    def barrier(self):
        import torch
        if self.multi:
            self.dist.barrier()
        torch.cuda.synchronize()

    def share(self, spp):
        """Strong scaling: this rank's iterations [first, first + n) of the frame's spp."""
        base, extra = divmod(spp, self.world)
        n = base + (1 if self.rank < extra else 0)
        first = self.rank * base + min(self.rank, extra)
        return first, n

    def max_sum(self, values):
        import torch
        t = torch.tensor(values, dtype=torch.float64, device="cuda")
        if not self.multi:
            return list(values), list(values)
        tmax, tsum = t.clone(), t.clone()
        self.dist.all_reduce(tmax, op=self.dist.ReduceOp.MAX)
        self.dist.all_reduce(tsum, op=self.dist.ReduceOp.SUM)
        return tmax.tolist(), tsum.tolist()


# ------------------------------------------------------------------------------ our arm
def bench_render(name, steps, warmup, D: Dist, with_clocks=True):
    """Times `steps` renders of workload `name` (strong-scaled over D.world ranks)."""
    import torch

    import cuda_path_tracer_b200 as pt
    from cuda_path_tracer_b200 import DisplayBufferType as DB

    wl = WORKLOADS[name]
    sd = make_scene(wl)
    W, H, spp, depth = wl["w"], wl["h"], wl["spp"], wl["depth"]
    scene = pt.Scene.from_description(sd, device=D.local_rank)
    stream = torch.cuda.Stream()
    tracer = pt.PathTracer(max_depth=depth, profile=True, stream=stream.cuda_stream,
                           samples_per_pass=int(os.environ.get("PT_SPP_PASS", "0")),
                           sort_rays=os.environ.get("PT_SORT_RAYS", "0") == "1")
    tracer.max_iterations = 1 << 30
    tracer.create_buffers((W, H), scene)
    sums = torch.zeros(W * H * 8, dtype=torch.float32, device="cuda")
    tracer.bind_sums(sums.data_ptr())
    first, n_local = D.share(spp)

    radiance = sums[: W * H * 4]   # colour sums + sample count (16 B/pixel); the G-buffer sums follow

    def step():
        sums.zero_()
        if n_local:
            tracer.render_range(sd.camera, first, n_local)
        if D.multi:
            # the frame is tonemapped, not denoised: only the radiance plane is reduced (the
            # G-buffer sums would ride along only for a denoiser, north_star / SURVEY 8e)
            D.dist.reduce(radiance, dst=0)

    out = {}
    with torch.cuda.stream(stream):
        for _ in range(warmup):
            step()
        D.barrier()
        tracer.reset_stats()
        sampler = ClockSampler(D.local_rank) if with_clocks else None
        if sampler:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        D.barrier()
        e0.record(stream)
        for _ in range(steps):
            step()
        e1.record(stream)
        D.barrier()
        clocks = sampler.stop() if sampler else None
        ms = e0.elapsed_time(e1)
        st = tracer.stats()
        rays = int(st.rays)
        launches_timed = int(st.kernel_launches)

        # end to end through the public API: host camera in, host RGBA8 image out
        pinned = torch.empty((H, W, 4), dtype=torch.uint8, pin_memory=True).numpy()  # D2H target
        host_img = None
        D.barrier()
        tracer.reset_stats()
        wall0 = time.perf_counter()
        for _ in range(steps):
            step()
            tracer.set_sample_count(spp)
            if D.rank == 0:
                host_img = tracer.send_to_preview(type=DB.color, out=pinned)  # tonemap + D2H (pinned) + sync
        D.barrier()
        wall_ms = (time.perf_counter() - wall0) * 1e3
        rays_e2e = int(tracer.stats().rays)

        # per-kernel times for the roofline: a one-lane pass, where every launch has the GPU to
        # itself (in the timed region above two half-passes run concurrently, so the event time of
        # one launch is not its own cost)
        alone = pt.PathTracer(max_depth=depth, profile=True, stream=stream.cuda_stream, lanes=1,
                              samples_per_pass=int(os.environ.get("PT_SPP_PASS", "0")))
        alone.max_iterations = 1 << 30
        tracer.close()
        alone.create_buffers((W, H), scene)
        k_steps = max(1, min(steps, 2))
        for i in range(k_steps + 1):
            if i == 1:
                torch.cuda.synchronize()
                alone.reset_stats()
            if n_local:
                alone.render_range(sd.camera, first, n_local)
        torch.cuda.synchronize()
        st = alone.stats()
        alone.close()
        tracer = pt.PathTracer(max_depth=depth, stream=stream.cuda_stream,
                               samples_per_pass=int(os.environ.get("PT_SPP_PASS", "0")))
        tracer.max_iterations = 1 << 30
        tracer.create_buffers((W, H), scene)
        tracer.bind_sums(sums.data_ptr())

        rmse = None
        if D.multi:
            # the reduced N-GPU frame (rank 0's sums after the last step) vs rank 0 alone
            multi_img = (sums[: W * H * 4].view(-1, 4)[:, :3] / float(spp)).clone() if D.rank == 0 else None
            D.barrier()
            if D.rank == 0:
                sums.zero_()
                tracer.render_range(sd.camera, 0, spp)
                torch.cuda.synchronize()
                single = sums[: W * H * 4].view(-1, 4)[:, :3] / float(spp)
                rmse = float(torch.sqrt(torch.mean((multi_img - single) ** 2)))
            D.barrier()

    (ms, wall_ms, _, _), (_, _, rays_all, rays_e2e_all) = D.max_sum([ms, wall_ms, float(rays), float(rays_e2e)])
    rays_all, rays_e2e_all = int(rays_all), int(rays_e2e_all)
    hbm, peak_src = peaks()
    gpu_launches = launches_timed
    ext_ms = float(st.ms_extend)
    ms_alone = ext_ms + float(st.ms_raygen_extend0) + float(st.ms_shade) + float(st.ms_accumulate)
    n_ext = max(1, int(st.n_extend_launches) - int(st.passes))
    achieved = (int(st.rays_traversed) * 48) / (ext_ms * 1e-3) * 1e-9 if ext_ms > 0 else None
    out.update(
        value=rays_all / (ms * 1e-3) * 1e-6, ms_per_step=ms / steps,
        spp_per_s=spp * steps / (ms * 1e-3), rays_per_step=rays_all // steps, clocks=clocks,
        e2e={"value": rays_e2e_all / (wall_ms * 1e-3) * 1e-6, "unit": "Mrays/s",
             "h2d_bytes_per_step": C.sizeof(pt._abi.pt_camera), "d2h_bytes_per_step": W * H * 4,
             "ms_per_step": wall_ms / steps},
        gpu_launches=gpu_launches,
        roofline={"bound": "hbm", "achieved": achieved, "peak": hbm, "unit": "GB/s",
                  "frac": (achieved / hbm) if achieved else None,
                  "traffic": ncu_traffic(name) if D.world == 1 else None,
                  "algorithmic_bytes_per_launch": int(st.rays_traversed) * 48 / n_ext,
                  "kernel": "traverse_kernel", "launches": n_ext,
                  "rays_traversed_per_step": int(st.rays_traversed) // k_steps,
                  "avg_launch_ms": ext_ms / n_ext,
                  "extend_share_of_step": ext_ms / ms_alone if ms_alone > 0 else None,
                  "timed": "launches of %d one-lane step(s) after the timed region, events on the launching stream; "
                           "the timed region itself runs two concurrent half-passes" % k_steps,
                  "wavefront_bytes_per_ray": 164,
                  "wavefront_frac": rays_all / D.world * 164 / (ms * 1e-3) * 1e-9 / hbm,
                  "peak_source": peak_src,
                  # HBM is the contract's denominator, not this kernel's limiter (ncu capture in
                  # profiles/r2_traverse_bunny_final_ncu_full_summary.csv)
                  "limiter": "instruction issue / ALU pipe on L2-resident scenes (ncu: issue-active 64-75 %, ALU "
                             "pipe 54-65 %, DRAM 56 B per ray); memory latency on scenes far beyond the L2",
                  "per_rank": "rank 0's launches" if D.multi else None},
        kernel_ms={"raygen_classify": st.ms_raygen_extend0 / k_steps, "traverse": st.ms_extend / k_steps,
                   "shade_classify_compact": st.ms_shade / k_steps, "accumulate": st.ms_accumulate / k_steps,
                   "note": "per step, one-lane pass (each launch alone on the GPU)"},
        scene={"triangles_world": int(scene.info.n_world_triangles), "bvh_nodes": int(scene.info.n_bvh_nodes),
               "scene_build_ms": round(float(scene.info.build_ms), 2),
               "scene_upload_ms": round(float(scene.info.upload_ms), 2),
               "path_state_mb": tracer_state_mb(W, H, n_local)},
        image_rmse_vs_single=rmse,
        image_mean=float(host_img[..., :3].mean()) if host_img is not None else None,
    )
    tracer.close()
    scene.close()
    return out, sd


def bench_frame(name, sd, w, h, spp, depth, filter_size, reps=30):
    """One interactive frame = restart + spp x path_trace (+ A-Trous) + tonemap.  `value`: device
    time (CUDA events, image left on the device); e2e: wall clock with the RGBA8 frame copied to
    pinned host memory."""
    import torch

    import cuda_path_tracer_b200 as pt
    stream = torch.cuda.Stream()
    tr = pt.PathTracer(max_depth=depth, stream=stream.cuda_stream)
    tr.max_iterations = 1 << 30
    tr.create_buffers((w, h), sd)
    tr.atrous_denoiser.filter_size = max(1, filter_size)
    pinned = torch.empty((h, w, 4), dtype=torch.uint8, pin_memory=True).numpy()
    dev_img = torch.empty((h, w, 4), dtype=torch.uint8, device="cuda")

    def frame(to_host):
        tr.restart()
        tr.render(sd.camera, spp)
        if filter_size:
            tr.denoise()
        if to_host:
            tr.send_to_preview(out=pinned)
        else:
            tr.send_to_preview(dev_pbo=dev_img.data_ptr())

    with torch.cuda.stream(stream):
        for _ in range(5):
            frame(True)
        dev_ms = []
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            frame(False)
            e1.record(stream)
            torch.cuda.synchronize()
            dev_ms.append(e0.elapsed_time(e1))
        wall = []
        for _ in range(reps):
            t0 = time.perf_counter()
            frame(True)
            wall.append((time.perf_counter() - t0) * 1e3)
        tr.reset_stats()
        frame(True)
        rays = int(tr.stats().rays)
        launches = int(tr.stats().kernel_launches)
        dn_ms = None
        if filter_size:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(reps):
                tr.denoise()
            e1.record(stream)
            torch.cuda.synchronize()
            dn_ms = e0.elapsed_time(e1) / reps
    hbm, _ = peaks()
    ms = float(np.median(dev_ms))
    iters = int(np.floor(np.log2(filter_size))) + 1 if filter_size else 0
    entry = {"workload": name, "metric": "ms per frame (device)", "unit": "ms", "higher_is_better": False,
             "value": ms, "value_min": float(min(dev_ms)), "mrays_per_s": rays / (ms * 1e-3) * 1e-6,
             "rays_per_frame": rays, "gpu_launches_per_frame": launches,
             "e2e": {"value": float(np.median(wall)), "min": float(min(wall)), "unit": "ms",
                     "h2d_bytes_per_step": 32, "d2h_bytes_per_step": w * h * 4},
             "config": {"w": w, "h": h, "spp": spp, "depth": depth, "denoise_filter_size": filter_size}}
    if dn_ms:
        alg = 40.0 * w * h * iters
        entry["denoise_ms"] = dn_ms
        entry["roofline"] = {"bound": "hbm", "kernel": "atrous_kernel x%d (+ prepare)" % iters,
                             "achieved": alg / (dn_ms * 1e-3) * 1e-9, "peak": hbm, "unit": "GB/s",
                             "frac": alg / (dn_ms * 1e-3) * 1e-9 / hbm,
                             "algorithmic_bytes": alg, "note": "40 B per pixel per iteration (SURVEY 8d)"}
    else:
        entry["roofline"] = {"bound": "hbm", "kernel": "whole frame (164 B per ray-bounce, SURVEY 8d)",
                             "achieved": rays * 164 / (ms * 1e-3) * 1e-9, "peak": hbm, "unit": "GB/s",
                             "frac": rays * 164 / (ms * 1e-3) * 1e-9 / hbm,
                             "note": "launch-latency bound at this size: %d launches per frame" % launches}
    tr.close()
    return entry


# ------------------------------------------------------------------------------ reference arm
def ref_library_for(sd, wl):
    """The stock (stack 24) build wherever the reference's own tree fits its stack, else stack 64."""
    from tests import ref_lib
    depth = None
    mesh = sd.meshes[sorted(sd.meshes)[0]] if sd.meshes else None
    if mesh is None:
        return ref_lib.load_ref_cuda_stock(), 0
    if mesh.triangle_count <= 3_000_000 and ref_lib.have_ref_cuda_stock():
        lib = ref_lib.load_ref_cuda_stock()
        depth = lib.bvh_depth(mesh)
        if depth <= 23:
            return lib, depth
    return ref_lib.load_ref_cuda(), depth


def ref_render(sd, wl, spp, steps, warmup):
    lib, depth = ref_library_for(sd, wl)
    t0 = time.perf_counter()
    rt = lib.tracer(sd, wl["w"], wl["h"], wl["depth"])   # build_scene(): bvh_from_mesh on one host thread + uploads
    create_s = time.perf_counter() - t0
    times, rays = [], 0
    for i in range(warmup + steps):
        ms, r = rt.render_timed(sd.camera, spp)
        if i >= warmup:
            times.append(ms)
            rays += r
    rt.close()
    total = sum(times)
    return {"value": rays / (total * 1e-3) * 1e-6, "unit": "Mrays/s", "ms_per_step": total / len(times),
            "spp_per_step": spp, "steps": steps, "scene_create_s": round(create_s, 3),
            "step_ms": [round(t, 1) for t in times], "best_step_mrays": max(0.0, rays / len(times) / (min(times) * 1e-3) * 1e-6),
            "build": {"library": os.path.basename(lib.path), "traversal_stack": lib.stack_size,
                      "reference_bvh_depth": depth, "flags": "-O3 -DNDEBUG -arch=sm_100 -rdc=true (the reference's)",
                      "patches": REF_PATCHES, "mode": "streaming (CLI default)"}}


def ref_frame(sd, w, h, spp, depth, filter_size, reps=10):
    from tests import ref_lib
    lib = ref_lib.load_ref_cuda_stock() if ref_lib.have_ref_cuda_stock() else ref_lib.load_ref_cuda()
    rt = lib.tracer(sd, w, h, depth)
    ts, gpu = [], []
    for i in range(reps + 3):
        t0 = time.perf_counter()
        rt.restart()
        ms, rays = rt.render_timed(sd.camera, spp, depth)
        dms = rt.denoise(filter_size) if filter_size else 0.0
        rt.preview(0)
        if i >= 3:
            ts.append((time.perf_counter() - t0) * 1e3)
            gpu.append(ms + dms)
    rt.close()
    return {"frame_ms_median": float(np.median(ts)), "frame_ms_min": float(min(ts)),
            "gpu_ms_min": float(min(gpu)), "library": os.path.basename(lib.path)}


def reference_secondary(which):
    """The reference's own numbers on the secondary configs (bounded samples: its render time
    per spp does not depend on how many spp follow)."""
    import cuda_path_tracer_b200 as pt
    out = []
    for name in which:
        try:
            if name == "terrain":
                wl = WORKLOADS["terrain"]
                r = ref_render(make_scene(wl), wl, 1, 2, 1)
                r["sample"] = "1 spp per step of the 16-spp workload (the reference renders one spp per call)"
                out.append({"workload": "terrain", **r})
            elif name == "three_balls_frame":
                out.append({"workload": name, **ref_frame(pt.three_balls(800, 800), 800, 800, 1, 5, 0)})
            elif name == "interactive_frame":
                out.append({"workload": name, **ref_frame(pt.bunny_scene(pt.bunny_like(4), 1920, 1080), 1920, 1080, 1, 8, 16)})
        except Exception as e:
            out.append({"workload": name, "reference_fails": f"{type(e).__name__}: {e}".splitlines()[0]})
    return out


def run_reference(args):
    """Reference arm: the reference's own CUDA build through its own PathTracer class."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    wl = WORKLOADS[args.workload]
    sd = make_scene(wl)
    line = {"impl": "reference", "metric": METRIC, "unit": "Mrays/s", "n_gpus": args.gpus, "n_gpus_used": 1,
            "steps": args.steps, "warmup": args.warmup, "higher_is_better": True,
            "scaling": "strong" if args.gpus > 1 else "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": config_of(args.workload)}
    why = None
    try:
        import torch
        torch.cuda.init()
        r = ref_render(sd, wl, wl["spp"], args.steps, args.warmup)
        v = r["value"]
        line.update(value=v, ms_per_step=r["ms_per_step"], spp_per_s=wl["spp"] / (r["ms_per_step"] * 1e-3),
                    best_step_value=r["best_step_mrays"], step_ms=r["step_ms"],  # it varies 2-3x between runs
                    gpu_launches=None, reference_build=r["build"],
                    cpu_baseline={"value": v, "unit": "Mrays/s", "cores": 1, "kind": "reference",
                                  "sample": "full workload on the reference's own CUDA build (oracle/_ref/%s, "
                                            "sm_100, streaming mode); 1 host thread drives it" % r["build"]["library"]},
                    e2e={"value": v, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0})
        if not args.no_secondary and args.workload == "bunny":
            line["secondary"] = reference_secondary(["terrain", "three_balls_frame", "interactive_frame"])
    except Exception as e:  # no reference build or no device on this box: time the CPU oracle port instead
        why = f"{type(e).__name__}: {e}".splitlines()[0]
    if why is not None:
        cb = cpu_baseline(sd, wl)
        cb["sample"] += f" [reference CUDA build unavailable: {why}]"
        line.update(value=cb["value"], ms_per_step=None, cpu_baseline=cb,
                    e2e={"value": cb["value"], "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0})
    print(json.dumps(line), flush=True)


def tracer_state_mb(W, H, n_local):
    spp_pass = int(os.environ.get("PT_SPP_PASS", "0")) or max(1, min(64, (1 << 27) // (W * H)))
    return int(W * H * min(spp_pass, max(1, n_local)) * 168 / 1e6)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="bunny", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true")
    args = ap.parse_args()

    if args.impl == "reference":
        run_reference(args)
        return

    D = Dist()
    if D.multi:
        # before the library (and its OpenMP runtime) loads: N processes share the host's threads
        # instead of each starting one per core (scene build 5 -> 97 ms across ranks in round 1)
        os.environ.setdefault("OMP_NUM_THREADS", str(max(1, (os.cpu_count() or 8) // D.world)))
    import cuda_path_tracer_b200 as pt
    pt.load_library()  # raises if the CUDA extension is missing: no fallback
    D.init()
    wl = WORKLOADS[args.workload]
    head, sd = bench_render(args.workload, args.steps, args.warmup, D)
    line = {"metric": METRIC, "value": head.pop("value"), "unit": "Mrays/s", "n_gpus": D.world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": head.pop("ms_per_step"),
            "higher_is_better": True, "scaling": "strong" if D.multi else "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": config_of(args.workload),
            "parallelism": f"sample ranges: {wl['spp']} spp split over {D.world} GPUs, one NCCL reduce of the "
                           f"radiance sums ({wl['w'] * wl['h'] * 16 / 1e6:.0f} MB) to rank 0" if D.multi else "single",
            **head}
    secondary = []
    if not args.no_secondary and args.workload == "bunny":
        t, _ = bench_render("terrain", 3, 3, D, with_clocks=False)
        secondary.append({"workload": "terrain", "metric": METRIC, "unit": "Mrays/s", "higher_is_better": True,
                          "config": config_of("terrain"), "steps": 3, "warmup": 3, **t})
        if not D.multi:
            secondary.append(bench_frame("three_balls_frame", pt.three_balls(800, 800), 800, 800, 1, 5, 0))
            secondary.append(bench_frame("interactive_frame", pt.bunny_scene(pt.bunny_like(4), 1920, 1080),
                                         1920, 1080, 1, 8, 16))
            from tests import ref_lib
            if ref_lib.have_ref_cuda():
                refs = {r["workload"]: r for r in reference_secondary([s["workload"] for s in secondary])}
                for s in secondary:
                    s["reference"] = refs.get(s["workload"])
            else:
                for s in secondary:
                    s["reference"] = {"reference_fails": "oracle/_ref/libref_cuda.so not built"}
    if D.rank == 0:
        if secondary:
            line["secondary"] = secondary
        if not args.no_cpu_baseline and D.world == 1:
            line["cpu_baseline"] = cpu_baseline(sd, wl)
        print(json.dumps(line), flush=True)
    if D.multi:
        D.dist.destroy_process_group()


if __name__ == "__main__":
    main()
