#!/usr/bin/env python
"""bench.py - frames/s of the full frame path at 3840x2160 (BASELINE.json metric, configs[2]) on N B200.

One "step" = one frame: the ordered pass list of RenderFrontend::prepareRenderpasses (histogram, exposure, sky LUTs,
HiZ, light matrices, SDF culling + trace, GI denoise + upscale, froxel volumetrics, G-buffer shading, TAA, bloom,
tonemap) over the synthetic Sponza-sized scene, replayed through the C-ABI of include/plain_b200.h.

  value : frames/s with the raster-pass outputs (depth, motion, normal, G-buffer, shadow maps) already resident in HBM
  e2e   : the same frame driven through plain_frontend_render_frame with HOST buffers: every step uploads that frame's
          raster-pass outputs from pinned host memory and reads the tonemapped 8-bit frame back, copies inside the
          timed region
  roofline     : the most expensive pass of the frame, algorithmic bytes (SURVEY.md 8d / DESIGN.md) / its mean duration
  cpu_baseline : the scalar C++ oracle (oracle/, kind "port") on the host cores, one real frame of the workload (rank 0, N = 1 only)

`--impl reference` times the reference's own CPU implementation of the path: its per-pixel arithmetic is GLSL that
cannot execute here (no Vulkan), so this is the oracle port on all host threads, on real frames of the named
configuration (1 warm-up + at most 3 timed frames: ~12 s each at 3840x2160).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
_REAL_STDOUT = sys.stdout

WIDTH, HEIGHT, INSTANCES = 3840, 2160, 100
CAMERA = ((-13.0, -1.7, 0.5), (1.0, 0.0, 0.0), (0.0, 0.0, 1.0), (0.0, -1.0, 0.0))  # eye height, looking down the atrium
SUN_DEG = (40.0, 35.0)
START_EXPOSURE = 2e-5
METRIC = "frames/s at 3840x2160 (full pipeline)"
HBM_FALLBACK_GBS = 6650.0  # B200_PROFILING.md fallback


def algorithmic_bytes(w, h):
    """Compulsory bytes per frame of each pass (every distinct input byte once + every output byte once), SURVEY.md 8d."""
    N, n = w * h, (w // 2) * (h // 2)
    F = ((w + 7) // 8) * ((h + 7) // 8) * 64
    b = {
        "Histogram per tile": 4 * N + 512,
        "Depth min/max pyramid creation": 4 * N + 8 * n * 4 // 3,
        "Depth downscale": 4 * n + 2 * n,
        "Indirect diffuse SDF trace": 8 * n + 12 * n,
        "Indirect diffuse spatial filter": 18 * n + 12 * n,
        "Indirect diffuse temporal filter": 32 * n + 24 * n,
        "Indirect lighting upscale": 4 * N + 14 * n + 12 * N,
        "Froxel volume material": 8 * F,
        "Froxel light scattering": 16 * F,
        "Volumetric lighting reprojection": 24 * F,
        "Volumetric light integration": 16 * F,
        "Forward shading": 32 * N + 8 * F,
        "Temporal filtering": 24 * N,
        "Apply bloom": 4 * N + 4 * N + N,
        "Tonemapping": 8 * N,
    }
    # bloom chain (Bloom.cpp:56-144, 6 mips): downsample mip k reads mip k-1 and writes mip k; upsample mip k reads the downsample
    # mip k+1 and - unless it is the lowest - the upsample mip k+1, writes mip k (R11G11B10, 4 bytes per texel)
    mip = lambda k: max(w >> k, 1) * max(h >> k, 1)
    for k in range(1, 6):
        b["Bloom downsample mip %d" % k] = 4 * mip(k - 1) + 4 * mip(k)
    for k in range(4, -1, -1):
        b["Bloom Upsample mip %d" % k] = 4 * mip(k + 1) * (1 if k == 4 else 2) + 4 * mip(k)
    return b


# passes every rank computes in full when a frame is split into row bands; all others only cover the rank's band
REPLICATED_PASSES = ("Sky ", "Pre-expose", "Histogram reset", "Histogram combine", "Compute light matrix", "SDF camera", "Bloom downsample mip 2", "Bloom downsample mip 3",
                     "Bloom downsample mip 4", "Bloom downsample mip 5", "Bloom Upsample mip 4", "Bloom Upsample mip 3", "Bloom Upsample mip 2")


NCU_KERNEL_OF_PASS = {"Indirect diffuse spatial filter": "giSpatialFilterKernel", "Indirect diffuse SDF trace": "sdfDiffuseTraceKernel",
                      "Temporal filtering": "temporalFilterKernel", "Forward shading": "gbufferShadingKernel", "Bloom Upsample mip 0": "bloomUpsampleKernel"}


def ncu_traffic(pass_name):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the pass's kernel from the newest committed `ncu --set full` capture
    of this workload (profiles/r2*_ncu_*.json, written by tools/ncu_summary.py): (bytes, file) or (None, None). A capture is a
    separate run under the profiler - the bench line says which file the figure comes from."""
    want = NCU_KERNEL_OF_PASS.get(pass_name)
    if not want:
        return None, None
    for f in sorted((ROOT / "profiles").glob("r[2-9]*_ncu_*.json"), reverse=True):
        try:
            rows = [r for r in json.loads(f.read_text()) if want in r["kernel"]]
        except Exception:
            continue
        if rows:
            return int(max(r["dram_bytes"] for r in rows)), f.name  # the largest launch = the full-resolution one
    return None, None


def measured_hbm_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return HBM_FALLBACK_GBS, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe): NVML polled every 5 ms from a
    thread (a 20-step region at 8 GPUs lasts tens of milliseconds - too short for nvidia-smi's 100 ms loop, the fallback)."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, device):
        self.device, self.sm, self.max_sm, self.reasons = device, [], None, set()
        self.stop_flag, self.thread, self.proc, self.source = False, None, None, None

    def _nvml_loop(self, nv, h):
        bits = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8), "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20), "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        while not self.stop_flag:
            try:
                self.sm.append(int(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                r = int(nv.nvmlDeviceGetCurrentClocksEventReasons(h)) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") else int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                for name, bit in bits.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.005)

    def _smi_loop(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.proc.stdout:
            f = [x.strip() for x in line.split(",")]
            if f and f[0].isdigit():
                self.sm.append(int(f[0]))
            if len(f) > 1 and f[1].isdigit():
                self.max_sm = int(f[1])
            for i in range(4):
                if len(f) >= 6 and f[2 + i].lower().startswith("active"):
                    self.reasons.add(names[i])

    def start(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            index = int(visible.split(",")[self.device]) if visible and all(x.strip().isdigit() for x in visible.split(",")) else self.device
            h = nv.nvmlDeviceGetHandleByIndex(index)
            self.max_sm = int(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            self.source = "nvml, 5 ms"
            self.thread = threading.Thread(target=self._nvml_loop, args=(nv, h), daemon=True)
            self.thread.start()
            return
        except Exception:
            pass
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.source = "nvidia-smi, 100 ms"
            threading.Thread(target=self._smi_loop, daemon=True).start()
        except Exception:
            self.proc = None

    def stop(self):
        self.stop_flag = True
        if self.thread:
            self.thread.join(timeout=1.0)
        if self.proc:
            self.proc.terminate()
        sm = sorted(self.sm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_sm, "reasons": sorted(self.reasons), "samples": len(sm), "source": self.source}


def oracle_api():
    """CPU oracle (test infrastructure): only for the cpu_baseline / --impl reference legs."""
    from plainrenderer_b200 import ffi
    lib = ROOT / "oracle" / "_build" / "liboracle.so"
    if not lib.exists():
        subprocess.run(["make", "-C", str(ROOT / "oracle")], check=True, capture_output=True)
    return ffi.Api(str(lib), "oracle_", "oracle_frontend_")


def time_oracle(width, height, frames, warm):
    """Median seconds per frame of the CPU oracle on all host threads at width x height (same scene, camera, sequence)."""
    from plainrenderer_b200 import ffi
    api = oracle_api()
    s = ffi.default_settings(api, width, height, sun_direction_deg=SUN_DEG)
    fe = ffi.Frontend(api, s)
    scene = ffi.SyntheticScene(api, n_instances=INSTANCES)
    scene.attach(fe)
    fe.set_exposure(START_EXPOSURE)
    cam = ffi.camera(*CAMERA)
    inputs = [scene.render_inputs(s, cam, f + 1, shadows=(f == 0)) for f in range(2)]
    times = []
    for f in range(warm + frames):
        i = inputs[f % 2]
        t0 = time.perf_counter()
        fe.render_frame(cam, (f + 1) / 60.0, 1 / 60.0, i["depth"], i["motion"], i["normal"], i["gbuffer"], inputs[0]["shadow_maps"] if f == 0 else None)
        fe.read_output()
        if f >= warm:
            times.append(time.perf_counter() - t0)
    fe.close()
    return float(np.median(times))


REFERENCE_MAX_TIMED_FRAMES = 3  # a 3840x2160 frame of the oracle takes ~12 s on 16 host threads: the arm stays within a few minutes


def run_reference(args, rank):
    """The reference's own CPU implementation of the path = the oracle port (its GLSL cannot execute here), timed on REAL frames of the
    named configuration: 1 warm-up frame + min(--steps, 3) timed frames at WIDTH x HEIGHT, nothing extrapolated. `steps` / `warmup` in
    the line are the counts actually run."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    frames, warm = max(min(args.steps, REFERENCE_MAX_TIMED_FRAMES), 1), 1
    sec = time_oracle(WIDTH, HEIGHT, frames, warm)
    fps = 1.0 / sec
    sample = "%d timed frame(s) after %d warm-up frame at the full %dx%d (median), same scene / camera / pass list, oracle port on all host threads" % (frames, warm, WIDTH, HEIGHT)
    line = {"impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": frames, "warmup": warm,
            "requested_steps": args.steps, "requested_warmup": args.warmup,
            "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.gpus),
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "the reference's per-pixel path is GLSL under Vulkan and cannot execute here; this arm is the scalar C++ oracle port (oracle/) on all host threads"}
    print(json.dumps(line), file=_REAL_STDOUT, flush=True)


def workload_config(n_gpus):
    name = "configs[2]" if (WIDTH, HEIGHT) == (3840, 2160) else "configs[4] (GI-trace-bound stress; not the headline metric)"
    return {"workload": "%s: %dx%d synthetic Sponza-sized scene (%d SDF instances), full pipeline (GI/TAA/sky/volumetrics/bloom), static camera with TAA jitter" % (name, WIDTH, HEIGHT, INSTANCES),
            "resolution": [WIDTH, HEIGHT], "sdf_instances": INSTANCES,
            "parallelism": "1 GPU" if n_gpus == 1 else "%d GPUs: every frame split into %d screen-space row bands (multiples of 32 rows), 5 exchanges per frame on the critical path over NVLink (histogram all-reduce, 4 row all-gathers) + 3 all-gathers of next-frame data behind the frame; sky LUTs, culling lists and bloom mips >= 2 replicated" % (n_gpus, n_gpus),
            "l2": "per-frame working set (>1.5 GB touched, G-buffer alone 133 MB) exceeds the 126 MB L2; no flush needed"}


def run_ours(args, rank, world, local_rank):
    import torch
    import plainrenderer_b200 as pr
    from plainrenderer_b200 import ffi

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the frame path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    api = pr.load(args.contract)
    from plainrenderer_b200 import sharding
    sharded = world > 1
    s = ffi.default_settings(api, WIDTH, HEIGHT, sun_direction_deg=SUN_DEG, shard_rank=rank if sharded else 0, shard_count=world if sharded else 0)
    fe = ffi.Frontend(api, s, device=local_rank)
    scene = ffi.SyntheticScene(api, n_instances=INSTANCES)
    scene.attach(fe)
    fe.set_exposure(START_EXPOSURE)
    cam = ffi.camera(*CAMERA)
    be = fe.backend
    band = sharding.full_res_band(api, HEIGHT, world, rank) if sharded else (0, HEIGHT)
    upload_rows = (max(band[0] - 16, 0), min(band[1] + 16, HEIGHT))  # band + the halo the stencils read (TAA, shading +-8, trace +-1)

    # raster-pass outputs for the TAA jitter phases, ray cast on the host into pinned memory (untimed set-up)
    def pinned(nbytes, dtype):
        t = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
        return t, t.numpy().view(dtype)
    keep, phases = [], []
    t_setup = time.time()
    n_phases = args.phases
    for ph in range(n_phases):
        bufs = {}
        for name, nbytes, dt in (("depth", WIDTH * HEIGHT * 4, np.float32), ("motion", WIDTH * HEIGHT * 4, np.int16), ("normal", WIDTH * HEIGHT * 4, np.uint8), ("gbuffer", WIDTH * HEIGHT * 16, np.uint32)):
            t, a = pinned(nbytes, dt)
            keep.append(t)
            bufs[name] = a
        if ph == 0:
            sm = []
            for _ in range(s.sun_shadow_cascade_count):
                t, a = pinned(2048 * 2048 * 2, np.uint16)
                keep.append(t)
                sm.append(a)
            bufs["shadow_maps"] = sm
        scene.render_inputs(s, cam, ph + 1, shadows=(ph == 0), out=bufs, threads=max((os.cpu_count() or 1) // world, 1))
        phases.append(bufs)
    out_t, out_host = pinned(WIDTH * HEIGHT * 4, np.uint8)
    setup_s = time.time() - t_setup

    stream_ptr = C.c_void_p()
    api.b["get_stream"](be.ctx, C.byref(stream_ptr))
    stream = torch.cuda.ExternalStream(stream_ptr.value, device=torch.device("cuda", local_rank))
    comm = sharding.DistComm(api, HEIGHT, device=torch.device("cuda", local_rank), stream=stream, frontend=fe, peer=not args.nccl_exchange) if sharded else None
    bytes_first_frame = [None]

    frame = [0]
    pass_acc, pass_order = {}, []

    def collect_timings(weight):
        for name, ms in be.pass_timings():
            if name not in pass_acc:
                pass_acc[name] = 0.0
                pass_order.append(name)
            pass_acc[name] += ms * weight

    def step(upload, readback, timing_weight=None):
        f = frame[0]
        frame[0] += 1
        ph = phases[f % n_phases]
        args = (ph["depth"], ph["motion"], ph["normal"], ph["gbuffer"], phases[0]["shadow_maps"] if f == 0 else None) if upload else (None, None, None, None, None)
        if sharded:
            fe.begin_frame(cam, (f + 1) / 60.0, 1 / 60.0, *args, async_upload=True, rows=upload_rows)
            while True:
                x = fe.run_segment()
                if x is None:
                    break
                comm.exchange(x)
            comm.check_peer_error(blocking=False)  # every frame, no synchronisation: a barrier time-out aborts the run instead of being timed
            if timing_weight is not None:
                collect_timings(timing_weight)  # the backend accumulates the timings of all segments of the frame
            if bytes_first_frame[0] is None:
                bytes_first_frame[0] = comm.bytes_sent  # the first frame runs every exchange through Python: its byte count is the per-frame figure
        else:
            fe.render_frame(cam, (f + 1) / 60.0, 1 / 60.0, *args, async_upload=True)
            if timing_weight is not None:
                collect_timings(timing_weight)
        if readback:
            fe.read_output_rows(out_host.reshape(HEIGHT, WIDTH * 4), band, async_pinned=True)

    def barrier():
        torch.cuda.synchronize()
        be._check(api.b["wait_for_gpu_idle"](be.ctx), "idle")
        if world > 1:
            torch.distributed.barrier()
            torch.cuda.synchronize()

    def timed(k, upload, readback):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(k):
            step(upload, readback)
        be._check(api.b["join_transfers"](be.ctx), "join_transfers")  # the end event also covers the copies on the upload / download streams
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    # warm-up: uploads every phase once (fills both ping-pong targets, shadow maps), lets exposure/TAA/GI histories settle
    be.set_graph_replay_enabled(not args.no_graph)
    # at least 6 untimed frames: a frame's pass list is replayed as a CUDA graph keyed by its resources, and those cycle with period 6 (two
    # presentable / ping-pong images x three motion buffers) - with fewer, a graph would be instantiated inside the timed region
    # row-sharded: the first three frames also map the exchanged images (three motion buffers) and run their exchanges through Python; the
    # submissions only take their final shape - and their graphs get instantiated - in the six frames after that
    warmup_frames = max(args.warmup, 6 if not sharded else 14, n_phases if n_phases <= 8 else 8)
    for _ in range(warmup_frames):
        step(True, True)
    barrier()

    clocks = ClockSampler(local_rank)
    clocks.start()
    ms_resident = timed(args.steps, upload=False, readback=False)
    launches = be.last_frame_launch_count()
    ms_e2e = timed(args.steps, upload=True, readback=True)
    clock_info = clocks.stop()

    # per-pass durations (CUDA events on the backend stream around every pass, mean over the same number of frames)
    be.set_graph_replay_enabled(False)
    be.set_timing_enabled(True)
    tk = max(min(args.steps, 10), 1)
    for _ in range(tk):
        step(False, False, timing_weight=1.0 / tk)
    be.set_timing_enabled(False)
    acc, order = pass_acc, pass_order
    barrier()

    if rank == 0:
        ms_step = ms_resident / args.steps
        fps = 1000.0 / ms_step      # N > 1: the SAME frame is split over the N GPUs (strong scaling)
        ms_step_e2e = ms_e2e / args.steps
        fps_e2e = 1000.0 / ms_step_e2e
        # per rank: band + halo of depth / normal / G-buffer / motion (a sharded rank all-gathers the motion vectors over NVLink); 1 GPU: whole images
        h2d = WIDTH * (upload_rows[1] - upload_rows[0]) * (4 + 4 + 16 + 4)
        d2h = WIDTH * (band[1] - band[0]) * 4
        peak, peak_src = measured_hbm_peak()
        alg_frame = algorithmic_bytes(WIDTH, HEIGHT)
        # rank 0's kernels cover rank 0's band: the bytes a launch moves are the band's share of the pass (replicated passes move all of them)
        band_share = (band[1] - band[0]) / float(HEIGHT)
        if acc.get("Indirect lighting upscale", 1.0) < 0.004:
            # pass fusion (the product's default): the upscale is folded into the shading kernel - its two full-resolution images
            # (12 bytes per pixel written, 12 read) do not exist; shading reads the upscale's inputs instead (full-res depth + half-res GI / depth)
            N_, n_ = WIDTH * HEIGHT, (WIDTH // 2) * (HEIGHT // 2)
            alg_frame["Forward shading"] += 4 * N_ + 14 * n_ - 12 * N_
            alg_frame["Indirect lighting upscale"] = 0
        froxel_chain = ("Froxel volume material", "Froxel light scattering", "Volumetric lighting reprojection")
        if all(acc.get(k, 1.0) < 0.004 for k in froxel_chain):
            # pass fusion: the four froxel passes run as ONE launch over froxel columns in the place of the integration pass - the material and scattering
            # volumes are never written, the reprojected texel goes from registers into the integration: history read 8 + history write 8 + integrated write 8
            # bytes per froxel (531 -> 199 MB at 3840x2160)
            F_ = ((WIDTH + 7) // 8) * ((HEIGHT + 7) // 8) * 64
            for k in froxel_chain:
                alg_frame[k] = 0
            alg_frame["Volumetric light integration"] = 24 * F_
        alg = {k: (v if (not sharded or k.startswith(REPLICATED_PASSES)) else int(v * band_share)) for k, v in alg_frame.items()}
        top = max(acc, key=lambda k: acc[k])
        launches_of_top = 2 if top == "Indirect diffuse spatial filter" else 1
        top_ms = acc[top] / launches_of_top
        achieved = alg.get(top, 0) / (top_ms * 1e-3) / 1e9 if top_ms > 0 else 0.0
        passes_sum = sum(acc.values())
        traffic, traffic_src = ncu_traffic(top) if not sharded else (None, None)
        pair_ms = acc.get("Forward shading", 0.0) + acc.get("Indirect diffuse SDF trace", 0.0)
        pair_bytes = alg["Forward shading"] + alg["Indirect diffuse SDF trace"]
        pair_gbs = pair_bytes / (pair_ms * 1e-3) / 1e9 if pair_ms > 0 else 0.0
        line = {"metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
                "higher_is_better": True, "scaling": "strong" if world > 1 else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(world),
                "e2e": {"value": fps_e2e, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": ms_step_e2e},
                "gpu_launches": launches * args.steps,
                "clocks": clock_info,
                "roofline": {"bound": "hbm", "kernel": top, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak if peak else None, "traffic": traffic,
                             "traffic_source": ("profiles/%s: a separate `ncu --set full` run of this workload, not this run" % traffic_src) if traffic_src else None,
                             "algorithmic_bytes_per_launch": alg.get(top, 0), "kernel_ms": top_ms, "share_of_frame": acc[top] / passes_sum if passes_sum else None, "peak_source": peak_src},
                # the north star's target pair: bytes(S1 shading) + bytes(S2 trace) over the sum of their durations
                "roofline_pair": {"kernels": ["Forward shading", "Indirect diffuse SDF trace"], "algorithmic_bytes": pair_bytes, "ms": round(pair_ms, 4), "achieved": pair_gbs, "unit": "GB/s",
                                  "frac": pair_gbs / peak if peak else None},
                "passes_ms": {k: round(acc[k], 4) for k in order},
                "frame_roofline": {"algorithmic_bytes_per_frame": int(sum(alg_frame.values()) + alg_frame["Indirect diffuse spatial filter"]),
                                   "hbm_bound_ms": (sum(alg_frame.values()) + alg_frame["Indirect diffuse spatial filter"]) / peak / 1e6},
                "setup_s": round(setup_s, 1), "graph_replay": not args.no_graph, "warmup_frames_run": warmup_frames}
        try:  # every pass against the HBM roofline (the north star asks for each kernel's achieved GB/s): algorithmic bytes / measured duration
            per_pass = {}
            for name, nbytes in alg.items():
                ms = acc.get(name, 0.0)
                if ms > 0:
                    launches_of = 2 if name == "Indirect diffuse spatial filter" else 1
                    gbps = nbytes * launches_of / (ms * 1e-3) / 1e9
                    per_pass[name] = {"ms": round(ms, 4), "GB/s": round(gbps, 1), "frac": round(gbps / peak, 4) if peak else None}
            line["passes_roofline"] = per_pass
        except Exception as e:  # never let a reporting extra break the bench line
            line["passes_roofline"] = {"error": str(e)}
        if args.contract != "exact":
            # never the headline by default: the fast contract matches the oracle within a tolerance, not bit for bit
            line["config"]["numeric_contract"] = "fast (EXPERIMENTAL, not the product's contract): SFU approximations + contraction in the floating-point passes (libplain_b200_fast.so)"
        if sharded:
            comm.check_peer_error()
            line["sharding"] = {"rows_of_rank0": list(band), "exchanges_per_frame": 8, "exchanges_on_the_critical_path": 5, "exchanges_deferred_behind_the_frame": 3, "bytes_sent_per_frame_rank0": bytes_first_frame[0],
                                "transport": "peer pushes over NVLink (CUDA IPC) + flag barriers, enqueued by run_segment" if comm.peer else "NCCL send/recv batches issued from Python",
                                "exchanges_through_python": comm.python_exchanges,
                                "note": "passes_ms are rank 0's kernels only (its band); e2e byte counts are per rank"}
        if world == 1 and not args.no_cpu_baseline:
            sec = time_oracle(WIDTH, HEIGHT, 1, 1)
            line["cpu_baseline"] = {"value": 1.0 / sec, "unit": "frames/s", "cores": os.cpu_count() or 1, "kind": "port",
                                    "sample": "1 timed frame after 1 warm-up frame at the full %dx%d, same scene / camera / pass list, oracle port on all host threads" % (WIDTH, HEIGHT)}
        print(json.dumps(line), file=_REAL_STDOUT, flush=True)
    fe.close()
    if world > 1:
        torch.distributed.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--phases", type=int, default=2, help="TAA jitter phases of raster-pass outputs generated on the host (each 232 MB at 4K)")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--nccl-exchange", action="store_true", help="N > 1: exchange over NCCL send/recv from Python instead of peer pushes over NVLink")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--contract", default="exact", choices=["exact", "fast"], help="exact (default): bit-exact against the oracle; fast: libplain_b200_fast.so, "
                    "hardware approximations in the floating-point passes, parity within a tolerance (DESIGN.md section 12)")
    ap.add_argument("--workload", default="c3", choices=["c3", "c5"], help="c3 (default, the headline metric): BASELINE configs[2], 3840x2160 / 100 instances; "
                    "c5: BASELINE configs[4], 7680x4320 / 256 instances (reported beside the headline, never instead of it)")
    args = ap.parse_args()
    global WIDTH, HEIGHT, INSTANCES, METRIC
    if args.workload == "c5":
        WIDTH, HEIGHT, INSTANCES = 7680, 4320, 256
        METRIC = "frames/s at 7680x4320 (full pipeline, 256 SDF instances)"
    # the driver parses ONE JSON line from stdout: everything else a library prints there (e.g. the NCCL version banner) goes to stderr
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
    else:
        if world != args.gpus and world == 1 and args.gpus > 1:
            print("bench.py: --gpus %d needs torchrun (one rank per GPU); running 1 rank" % args.gpus, file=sys.stderr)
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
