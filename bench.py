#!/usr/bin/env python
"""bench.py — headline benchmark of the B200 wavefront path tracer.

Metric (BASELINE.json): Mrays/s over all bounces (+ spp/s) on the bundled OBJ triangle-mesh
scene (assets/scenes/bunny.json with a procedurally generated bunny.obj — the reference's
models are git-LFS stubs), 1920x1080, 64 spp, max depth 8.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
                  [--workload bunny|bunny_1m|three_balls|terrain]

A "step" is one 64-spp render of the frame (all bounces of every wavefront pass).  `value` is
timed with CUDA events on the launching stream with the scene resident in HBM; `e2e` goes
through the public PathTracer API from a host camera struct to a host RGBA8 image (tonemap +
D2H inside the timed region).  N > 1: one process per GPU (torchrun), sample-range sharding —
rank r renders iterations [r*spp, (r+1)*spp) of the same frame — then one NCCL reduce of the
radiance/G-buffer sums to rank 0 ("weak": per-GPU work fixed, N x the samples).
`--impl reference` times the reference's own CUDA build (oracle/_ref/libref_cuda.so, compiled
from /root/reference) on the same scene/config; if it is absent, the CPU oracle port.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

WORKLOADS = {
    # name: (builder kwargs, width, height, spp, max_depth)
    "bunny": dict(kind="bunny", subdiv=4, w=1920, h=1080, spp=64, depth=8),
    "bunny_82k": dict(kind="bunny", subdiv=6, w=1920, h=1080, spp=64, depth=8),
    "bunny_1m": dict(kind="bunny", subdiv=8, w=1920, h=1080, spp=64, depth=8),
    "three_balls": dict(kind="balls", w=800, h=800, spp=1, depth=5),
    "terrain": dict(kind="terrain", n=2236, w=3840, h=2160, spp=16, depth=8),
    "terrain_small": dict(kind="terrain", n=700, w=3840, h=2160, spp=16, depth=8),
}


def make_scene(wl):
    import cuda_path_tracer_b200 as pt
    if wl["kind"] == "bunny":
        return pt.bunny_scene(pt.bunny_like(wl["subdiv"]), wl["w"], wl["h"], wl["spp"])
    if wl["kind"] == "balls":
        return pt.three_balls(wl["w"], wl["h"], wl["spp"])
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
            self._stop.wait(0.1)

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
    (profiles/r1_traverse_traffic_v7.json); None for workloads that were not captured."""
    p = os.path.join(ROOT, "profiles", "r1_traverse_traffic_v7.json")
    try:
        with open(p) as f:
            j = json.load(f)
        return float(j["dram_bytes_per_launch"]) if j.get("workload") == workload else None
    except Exception:
        return None


def cpu_baseline(sd, wl):
    """The oracle's CPU port (reference megakernel semantics, OpenMP over pixels) on a bounded
    sample of the workload: the same scene and depth at quarter resolution, 1 spp."""
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


def run_reference(args, wl, sd):
    """Reference arm: the reference's own CUDA build through its own PathTracer class."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    line = {"impl": "reference", "metric": "Mrays/s (all bounces)", "unit": "Mrays/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, **{k: v for k, v in wl.items() if k != "kind"}}}
    why = None
    try:
        from tests.ref_lib import load_ref_cuda
        ref = load_ref_cuda()
        import torch
        torch.cuda.init()
        rt = ref.tracer(sd, wl["w"], wl["h"], wl["depth"])
        times, rays = [], 0
        for i in range(args.warmup + args.steps):
            ms, r = rt.render_timed(sd.camera, wl["spp"])
            if i >= args.warmup:
                times.append(ms)
                rays += r
        total_ms = sum(times)
        v = rays / (total_ms * 1e-3) * 1e-6
        line.update(value=v, ms_per_step=total_ms / len(times), spp_per_s=wl["spp"] * len(times) / (total_ms * 1e-3),
                    gpu_launches=None,
                    cpu_baseline={"value": v, "unit": "Mrays/s", "cores": 1, "kind": "reference",
                                  "sample": "full workload on the reference's own CUDA build (oracle/_ref/libref_cuda.so, "
                                            "sm_100, streaming mode); 1 host thread drives it"},
                    e2e={"value": v, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0})
    except Exception as e:  # no reference build or no device on this box: time the CPU oracle port instead
        why = f"{type(e).__name__}: {e}".splitlines()[0]
    if why is not None:
        cb = cpu_baseline(sd, wl)
        cb["sample"] += f" [reference CUDA build unavailable: {why}]"
        line.update(value=cb["value"], ms_per_step=None, cpu_baseline=cb,
                    e2e={"value": cb["value"], "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0})
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="bunny", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]

    import cuda_path_tracer_b200 as pt
    from cuda_path_tracer_b200 import DisplayBufferType as DB

    sd = make_scene(wl)
    if args.impl == "reference":
        run_reference(args, wl, sd)
        return

    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    multi = world > 1
    torch.cuda.set_device(local_rank)
    if multi:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    pt.load_library()
    W, H, spp, depth = wl["w"], wl["h"], wl["spp"], wl["depth"]
    scene = pt.Scene.from_description(sd, device=local_rank)
    stream = torch.cuda.Stream()
    tracer = pt.PathTracer(max_depth=depth, profile=True, stream=stream.cuda_stream,
                           samples_per_pass=int(os.environ.get("PT_SPP_PASS", "0")))
    tracer.max_iterations = 1 << 30
    tracer.create_buffers((W, H), scene)
    sums = torch.zeros(W * H * 8, dtype=torch.float32, device="cuda")
    tracer.bind_sums(sums.data_ptr())
    first = rank * spp  # sample-range sharding: this rank's iterations

    def step():
        sums.zero_()
        tracer.render_range(sd.camera, first, spp)
        if multi:
            dist.reduce(sums, dst=0)

    def barrier():
        if multi:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.cuda.stream(stream):
        for _ in range(args.warmup):
            step()
        barrier()
        tracer.reset_stats()
        sampler = ClockSampler(local_rank)
        sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record(stream)
        for _ in range(args.steps):
            step()
        e1.record(stream)
        barrier()
        clocks = sampler.stop()
        ms = e0.elapsed_time(e1)
        st = tracer.stats()
        rays = int(st.rays)

        # end to end through the public API: host camera in, host RGBA8 image out
        tracer.set_sample_count(spp)
        host_img = None
        pinned = torch.empty((H, W, 4), dtype=torch.uint8, pin_memory=True).numpy()  # D2H target
        t_e2e0 = torch.cuda.Event(enable_timing=True)
        t_e2e1 = torch.cuda.Event(enable_timing=True)
        barrier()
        launches_before = int(tracer.stats().kernel_launches)
        tracer.reset_stats()
        wall0 = time.perf_counter()
        t_e2e0.record(stream)
        for _ in range(args.steps):
            step()
            tracer.set_sample_count(spp * world)
            host_img = tracer.send_to_preview(type=DB.color, out=pinned)  # tonemap + D2H (pinned) + sync
        t_e2e1.record(stream)
        barrier()
        wall_e2e = time.perf_counter() - wall0
        st2 = tracer.stats()
        rays_e2e = int(st2.rays)

    t = torch.tensor([ms, float(rays), wall_e2e * 1e3, float(rays_e2e)], dtype=torch.float64, device="cuda")
    if multi:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms, wall_ms = float(tmax[0]), float(tmax[2])
        rays, rays_e2e = int(tsum[1]), int(tsum[3])
    else:
        wall_ms = wall_e2e * 1e3

    if rank == 0:
        hbm, peak_src = peaks()
        value = rays / (ms * 1e-3) * 1e-6
        # roofline of the dominant kernel (traverse_kernel): algorithmic bytes = 48 B per ray it
        # processes (ray 32 B in + 16 B hit record out) over its summed launch time
        ext_ms = st.ms_extend
        n_ext = max(1, int(st.n_extend_launches) - int(st.passes))
        achieved = (int(st.rays_traversed) * 48) / (ext_ms * 1e-3) * 1e-9 if ext_ms > 0 else None
        line = {
            "metric": "Mrays/s (all bounces)", "value": value, "unit": "Mrays/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": args.workload, **{k: v for k, v in wl.items() if k != "kind"},
                       "triangles_world": int(scene.info.n_world_triangles),
                       "bvh_nodes": int(scene.info.n_bvh_nodes),
                       "scene_build_ms": round(float(scene.info.build_ms), 2),
                       "scene_upload_ms": round(float(scene.info.upload_ms), 2),
                       "l2": "no flush: wavefront state (ray/hit/throughput planes, %d MB) exceeds the 126 MB L2"
                             % (tracer_state_mb(W, H, tracer)),
                       "parallelism": f"sample-range x{world}" if multi else "single"},
            "spp_per_s": spp * world * args.steps / (ms * 1e-3),
            "rays_per_step": rays // args.steps,
            "clocks": clocks,
            "e2e": {"value": rays_e2e / (wall_ms * 1e-3) * 1e-6, "unit": "Mrays/s",
                    "h2d_bytes_per_step": C.sizeof(pt._abi.pt_camera), "d2h_bytes_per_step": W * H * 4,
                    "ms_per_step": wall_ms / args.steps},
            "gpu_launches": int(st.kernel_launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm, "unit": "GB/s",
                         "frac": (achieved / hbm) if achieved else None,
                         "traffic": ncu_traffic(args.workload) if world == 1 else None,
                         "algorithmic_bytes_per_launch": int(st.rays_traversed) * 48 / n_ext,
                         "kernel": "traverse_kernel", "launches": n_ext,
                         "rays_traversed_per_step": int(st.rays_traversed) // args.steps,
                         "avg_launch_ms": ext_ms / n_ext,
                         "extend_share_of_step": ext_ms / ms,
                         "wavefront_bytes_per_ray": 164,
                         "wavefront_frac": rays / world * 164 / (ms * 1e-3) * 1e-9 / hbm,
                         "peak_source": peak_src,
                         "observed_bound": "l1tex data-pipe wavefronts 74 % of peak, issue 48 %, 16/32 lanes "
                                           "(profiles/r1_traverse_bvh2_v6_ncu_full_summary.csv)",
                         "note": "divergent 64-byte node gathers through L1 bound this kernel, not HBM: the "
                                 "HBM fraction is reported as the contract asks, the ncu summaries explain it"},
            "kernel_ms": {"raygen_classify": st.ms_raygen_extend0, "traverse": st.ms_extend,
                          "shade_classify_compact": st.ms_shade, "accumulate": st.ms_accumulate},
        }
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline(sd, wl)
        if host_img is not None:
            line["image_mean"] = float(host_img[..., :3].mean())
        print(json.dumps(line), flush=True)
    if multi:
        dist.destroy_process_group()


def tracer_state_mb(W, H, tracer):
    spp_pass = int(os.environ.get("PT_SPP_PASS", "0")) or max(1, min(64, (1 << 27) // (W * H)))
    return int(W * H * spp_pass * 168 / 1e6)


if __name__ == "__main__":
    main()
