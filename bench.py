#!/usr/bin/env python
"""bench.py — upscaled Mpixels/s of the FSR 1.0 hot path (EASU+RCAS) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

One "step" = one synthetic frame through EASU -> RCAS.  Default workload = BASELINE.json configs[1]:
1920x1080 -> 3840x2160, RGBA16F, sharpness 0.25 stops.  With N > 1 (launched by torchrun, one rank per GPU)
the frame is N slabs tall (1920 x 1080N -> 3840 x 2160N), sharded by output row slab with the EASU input halo
exchanged between neighbouring ranks over NCCL every step: per-GPU work is fixed ("weak" scaling).

Timing rules followed: >= 3 warm-up steps; frames rotate through a ring of buffer sets larger than L2
(so no step finds its input or output in cache from the previous one); the timed region is bracketed by a
barrier + cuda synchronize, measured with CUDA events on the launching stream, max over ranks; SM clocks and
throttle reasons are sampled with nvidia-smi during the timed region.
Prints ONE JSON line (rank 0).
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

WORKLOADS = {
    # name: (in_w, in_h, out_w, out_h, dtype, description)
    "1080p-4k-fp16": (1920, 1080, 3840, 2160, "f16", "1920x1080->3840x2160 EASU+RCAS RGBA16F (BASELINE configs[1])"),
    "1440p-4k-fp16": (2560, 1440, 3840, 2160, "f16", "2560x1440->3840x2160 EASU+RCAS RGBA16F (BASELINE configs[2])"),
    "uq-4k-fp16": (2953, 1661, 3840, 2160, "f16", "2953x1661->3840x2160 (true Ultra Quality 1.3x) RGBA16F"),
    "1080p-4k-fp32": (1920, 1080, 3840, 2160, "f32", "1920x1080->3840x2160 EASU+RCAS RGBA32F (BASELINE configs[3])"),
    "2160p-8k-fp16": (3840, 2160, 7680, 4320, "f16", "3840x2160->7680x4320 EASU+RCAS RGBA16F (BASELINE configs[4])"),
}
SHARPNESS = 0.25
RING = 8
METRIC = "upscaled Mpixels/sec (EASU+RCAS)"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def load_traffic():
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p))
        except Exception:
            pass
    return {}


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region: NVML polled every ~2 ms from a thread
    (the timed region is tens of milliseconds, too short for `nvidia-smi -lms`), nvidia-smi as fallback."""
    BITS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index):
        self.index, self.samples, self.stop_flag, self.t, self.err = index, [], False, None, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index(index))
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as e:  # noqa: BLE001
            self.nv, self.err = None, repr(e)

    @staticmethod
    def _physical_index(i):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            try:
                return int(vis.split(",")[i])
            except Exception:
                pass
        return i

    def _poll(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                try:
                    reasons = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    reasons = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                pw = nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0
                self.samples.append((sm, reasons, pw))
            except Exception as e:  # noqa: BLE001
                self.err = repr(e)
                break
            time.sleep(0.002)

    def start(self):
        if self.nv is None:
            return
        self.t = threading.Thread(target=self._poll, daemon=True)
        self.t.start()

    def stop(self):
        if self.nv is None:
            return self._smi_once()
        self.stop_flag = True
        self.t.join(timeout=1)
        sm = sorted(s[0] for s in self.samples)
        reasons = set()
        for _, r, _ in self.samples:
            for bit, name in self.BITS.items():
                if r & bit:
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_sm, "sm_mhz_min": sm[0] if sm else None,
                "power_w_max": max((s[2] for s in self.samples), default=None), "samples": len(sm),
                "reasons": sorted(reasons), "how": "NVML polled every 2 ms during the timed region"}

    def _smi_once(self):
        try:
            out = subprocess.check_output(["nvidia-smi", "-i", str(self.index), "--query-gpu=clocks.sm,clocks.max.sm",
                                           "--format=csv,noheader,nounits"], text=True).strip().split(",")
            return {"sm_mhz": float(out[0]), "sm_max_mhz": float(out[1]), "samples": 1, "reasons": [],
                    "how": "nvidia-smi once after the timed region (NVML unavailable: %s)" % self.err}
        except Exception:
            return {"sm_mhz": None, "sm_max_mhz": None, "samples": 0, "reasons": ["unavailable"]}


# ------------------------------------------------------------------------------------------- reference arm
def cpu_reference_runner():
    """Returns (kind, fn(frame_f32, ow, oh, y0, y1) -> None) timing the reference's own CPU-compilable source
    (oracle/_ref) when it was built, else the oracle port."""
    import oracle_lib as ol
    R = ol.ref()
    lib, kind = (R, "reference") if R is not None else (ol.oracle(), "port")
    import numpy as np

    def run(frame, ow, oh, y0, y1, econ, rcon, tmp_holder):
        e0, e1 = max(y0 - 1, 0), min(y1 + 1, oh)
        ih, iw = frame.shape[:2]
        tmp, out = tmp_holder
        P, Z = ctypes.c_void_p, ctypes.c_size_t
        ec, rc = (ctypes.c_uint32 * 16)(*econ), (ctypes.c_uint32 * 4)(*rcon)
        if kind == "reference":
            lib.fsr1ref_easu_f(P(frame.ctypes.data), iw, ih, Z(iw * 4), P(tmp.ctypes.data), ow, oh, Z(ow * 4), ec, e0, e1)
            lib.fsr1ref_rcas_f(P(tmp.ctypes.data), ow, oh, Z(ow * 4), P(out.ctypes.data), Z(ow * 4), rc, 0, y0, y1)
        else:
            lib.fsr1o_easu_f32(P(frame.ctypes.data), iw, ih, Z(iw * 4), P(tmp.ctypes.data), ow, oh, Z(ow * 4), ec, e0, e1)
            lib.fsr1o_rcas_f32(P(tmp.ctypes.data), ow, oh, Z(ow * 4), P(out.ctypes.data), Z(ow * 4), rc, 0, y0, y1)
    return kind, run, ol


def time_cpu(wl, steps, warmup, budget_s):
    """Times the CPU reference on bands of output rows of the workload; returns dict(value, cores, kind, sample, s)."""
    import numpy as np
    import fsr1_b200 as F
    iw, ih, ow, oh = wl[:4]
    kind, run, ol = cpu_reference_runner()
    frame = F.to_half(F.uniform(iw, ih, 12345)).astype(np.float32)
    econ, rcon = ol.easu_con(iw, ih, ow, oh), ol.rcas_con(SHARPNESS)
    holder = (np.zeros((oh, ow, 4), np.float32), np.zeros((oh, ow, 4), np.float32))
    # calibrate on 32 rows
    t = time.perf_counter(); run(frame, ow, oh, 0, 32, econ, rcon, holder); per_row = (time.perf_counter() - t) / 32
    t = time.perf_counter(); run(frame, ow, oh, 64, 96, econ, rcon, holder); per_row = min(per_row, (time.perf_counter() - t) / 32)
    rows = int(max(8, min(oh, budget_s / max(per_row, 1e-9) / max(steps + warmup, 1))))
    bands = [(y, min(y + rows, oh)) for y in range(0, oh - rows + 1, rows)] or [(0, oh)]
    for i in range(warmup):
        y0, y1 = bands[i % len(bands)]
        run(frame, ow, oh, y0, y1, econ, rcon, holder)
    t0 = time.perf_counter()
    px = 0
    for i in range(steps):
        y0, y1 = bands[i % len(bands)]
        run(frame, ow, oh, y0, y1, econ, rcon, holder)
        px += (y1 - y0) * ow
    dt = time.perf_counter() - t0
    cores = int(os.environ.get("OMP_NUM_THREADS", os.cpu_count() or 1))
    return {"value": px / dt / 1e6, "unit": "Mpix/s", "cores": cores, "kind": kind,
            "sample": "%d steps x %d-row bands of the %dx%d output, fp32 F path, OpenMP over rows" % (steps, rows, ow, oh),
            "seconds": dt, "ms_per_step": dt / steps * 1e3}


# ------------------------------------------------------------------------------------------- our arm
def run_ours(args, rank, world, local_rank):
    global RING
    if args.frames > 0:
        RING = args.frames
    import numpy as np
    import torch
    import fsr1_b200 as F
    api = F.api
    wl = WORKLOADS[args.workload]
    iw, ih, ow, oh, dts = wl[:5]
    tdt = torch.float16 if dts == "f16" else torch.float32
    bpp = 8 if dts == "f16" else 16
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    K, W = args.steps, max(args.warmup, 3)

    def host_frame(seed, rows=None, first=0):
        f = F.uniform(iw, ih, seed)
        return f.astype(np.float16) if dts == "f16" else f

    econ1, rcon = api.easu_con(iw, ih, iw, ih, ow, oh), api.rcas_con(SHARPNESS)
    stream = torch.cuda.current_stream()
    out = {}
    pipe = None   # api.FramePipeline when frames are software-pipelined on two streams (default at N=1)
    if world == 1:
        def resident(a):  # rows padded to a 16-byte multiple, like any texture allocation (TMA / 128-bit access need it)
            h, w = a.shape[:2]
            buf = torch.zeros((h, (w + 1) & ~1, 4), dtype=tdt, device=dev)
            buf[:, :w] = torch.from_numpy(a).to(dev)
            return buf[:, :w]
        ins = [resident(host_frame(12345 + t)) for t in range(RING)]
        tmps = [torch.empty((oh, ow, 4), dtype=tdt, device=dev) for _ in range(RING)]
        outs = [torch.empty((oh, ow, 4), dtype=tdt, device=dev) for _ in range(RING)]
        imgs = [(api.image(ins[i]), api.image(tmps[i]), api.image(outs[i])) for i in range(RING)]
        prepared = [api.PreparedUpscale(a, t, b, econ1, rcon) for a, t, b in imgs]   # arguments marshalled once per buffer set
        pipe = None if args.no_pipeline else api.FramePipeline(imgs, econ1, rcon, device=dev)

        def step_seq(i):
            prepared[i % RING].launch(stream)

        def step(i):
            if pipe is None:
                step_seq(i)
            else:
                pipe.submit(i % RING)      # RCAS of frame i overlaps EASU of frame i+1 (two streams)

        def step_easu(i):
            a, t, _ = imgs[i % RING]
            api.easu(a, t, econ1, stream=stream)

        def step_rcas(i):
            _, t, b = imgs[i % RING]
            api.rcas(t, b, rcon, stream=stream)
        total_out_px = ow * oh
        halo = 0
    else:
        # default ("weak"): the frame is `world` slabs tall, this rank owns input rows [rank*ih, (rank+1)*ih) of it;
        # --shard-frame ("strong"): the named frame itself is cut into `world` row slabs (BASELINE configs[4] at N=8)
        H_in, H_out = (ih, oh) if args.shard_frame else (ih * world, oh * world)
        ups = [F.ShardedUpscaler(iw, H_in, ow, H_out, world, rank, SHARPNESS, dtype=tdt, device=dev) for _ in range(RING)]
        for t, u in enumerate(ups):  # each frame's slab is resident in HBM inside its halo window
            src = torch.from_numpy(host_frame(12345 + t + (0 if args.shard_frame else 1000 * rank))).to(dev)
            o0, o1 = u.plan.owned_in_rows(rank)
            u.owned.copy_(src[o0:o1] if args.shard_frame else src)
        if args.graph:
            for u in ups:            # halo exchange + EASU + RCAS recorded once; a step is one graph launch
                u.capture()

            def step(i):
                ups[i % RING].upscale()
        elif args.no_overlap:
            def step(i):
                ups[i % RING].upscale()
        elif args.halo_depth > 1 or args.pipeline_sharded or args.halo_batch > 1:
            # EXPERIMENTAL (not the default): halo exchange `halo_depth` frames ahead (more slack against rank-to-rank
            # jitter than the one-frame prefetch below) and, with --pipeline-sharded, RCAS of frame i on a second
            # stream while EASU of frame i+1 runs (api.FramePipeline over the slab windows), as at N=1
            depth = max(1, min(args.halo_depth, RING - 1))
            comm = torch.cuda.Stream(device=dev)
            ready = [torch.cuda.Event() for _ in range(RING)]
            done = [torch.cuda.Event() for _ in range(RING)]
            used = [False] * RING
            if args.pipeline_sharded:
                u0 = ups[0]
                e0, e1 = u0.plan.easu_rows(rank)
                y0, y1 = u0.plan.out_rows(rank)
                pipe = api.FramePipeline(
                    [(api.image(u.window, height=u.in_h, row0=u._win0), api.image(u.tmp, height=u.out_h, row0=e0),
                      api.image(u.out, height=u.out_h, row0=y0)) for u in ups],
                    ups[0].econ, ups[0].rcon, device=dev, easu_rows=(e0, e1), rcas_rows=(y0, y1))

            def prefetch(j):
                k = j % RING
                if used[k]:
                    comm.wait_event(done[k])     # frame j-RING, the window's previous reader
                with torch.cuda.stream(comm):
                    ups[k]._exchange()
                    ready[k].record(comm)
            batch = max(1, min(args.halo_batch, RING // 2))

            def prefetch_group(g):               # --halo-batch: the halos of `batch` consecutive frames in ONE NCCL group
                slots = [(g * batch + t) % RING for t in range(batch)]
                for k in slots:
                    if used[k]:
                        comm.wait_event(done[k])
                with torch.cuda.stream(comm):
                    F.ShardedUpscaler.exchange_many([ups[k] for k in slots])
                    for k in slots:
                        ready[k].record(comm)
            if batch > 1:
                prefetch_group(0)
            else:
                for j in range(depth):
                    prefetch(j)
            frame = [0]

            def step(_):
                i = frame[0]
                frame[0] += 1
                k = i % RING
                if batch > 1:
                    if i % batch == 0:
                        prefetch_group(i // batch + 1)
                else:
                    prefetch(i + depth)
                if pipe is not None:
                    pipe.stream_easu.wait_event(ready[k])
                    pipe.submit(k)
                    done[k].record(pipe.stream_easu)      # EASU is the only reader of the window
                else:
                    stream.wait_event(ready[k])
                    ups[k]._launch(stream)
                    done[k].record(stream)
                used[k] = True
        else:
            # frame pipeline: while frame i is upscaled, the halo rows of frame i+1 (already resident) travel on a
            # second stream; kernels of frame i+1 wait on that exchange's event only
            comm = torch.cuda.Stream(device=dev)
            ready = [torch.cuda.Event() for _ in range(RING)]

            def prefetch(j):
                u = ups[j % RING]
                comm.wait_stream(stream)         # the window's previous readers (frame j-RING) are ordered before
                with torch.cuda.stream(comm):
                    u._exchange()
                    ready[j % RING].record(comm)
            prefetch(0)
            frame = [0]

            def step(_):
                i = frame[0]
                frame[0] += 1
                prefetch(i + 1)
                stream.wait_event(ready[i % RING])
                ups[i % RING]._launch(stream)
        step_easu = step_rcas = None
        total_out_px = ow * H_out
        halo = ups[0].plan.halo_bytes(rank, iw, bpp)

    def timed(fn, n, pre=None, post=None):
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        if pre:
            pre()
        for i in range(n):
            fn(i)
        if post:
            post()
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if dist is not None:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dist.barrier()
            ms = float(t.item())
        return ms

    for i in range(W):
        step(i)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = api.launch_count()
    piped = pipe is not None
    ms = timed(step, K, pre=(lambda: pipe.begin(stream)) if piped else None, post=(lambda: pipe.end(stream)) if piped else None)
    launches = api.launch_count() - launches0
    if world > 1 and args.graph:
        launches = 2 * K  # each replayed graph holds this library's two kernels (EASU, RCAS), recorded at capture
    clocks = sampler.stop() if rank == 0 else None
    value = total_out_px * K / (ms * 1e-3) / 1e6

    peak, peak_src = load_peaks()
    kernels, roofline = {}, None
    latency_ms = None
    if world == 1:
        if piped:
            for i in range(W):
                step_seq(i)
            latency_ms = timed(step_seq, K) / K      # one frame at a time on one stream: EASU then RCAS, no overlap
        Pin, Pout = iw * ih, ow * oh
        alg = {"easu": bpp * (Pin + Pout), "rcas": bpp * 2 * Pout}
        names = {}
        for key, fn in (("easu", step_easu), ("rcas", step_rcas)):
            for i in range(W):
                fn(i)
            names[key] = api.last_kernel()
            kms = timed(fn, K) / K
            gbs = alg[key] / (kms * 1e-3) / 1e9
            kernels[key] = {"kernel": names[key], "us": kms * 1e3, "algorithmic_bytes": alg[key], "GBps": gbs,
                            "frac_of_hbm_peak": gbs / peak}
        dom = max(kernels, key=lambda k: kernels[k]["us"])
        traffic = load_traffic().get(kernels[dom]["kernel"])
        roofline = {"bound": "hbm", "kernel": kernels[dom]["kernel"], "achieved": kernels[dom]["GBps"], "peak": peak,
                    "unit": "GB/s", "frac": kernels[dom]["GBps"] / peak, "traffic": traffic, "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": alg[dom], "us_per_launch": kernels[dom]["us"],
                    "note": "EASU is instruction-issue-bound on B200, not HBM-bound: 182 instructions per output pixel at the "
                            "2.1-2.6 inst/cycle/SM this mix can issue (HFMA2, FFMA2 and scalar FFMA all sustain ~2 inst/cycle/SM: "
                            "profiles/r01_ubench_pipes.txt) is >= 58 us per 4K frame vs 12.6 us at the HBM roofline; the kernel "
                            "runs at 2.53 inst/cycle/SM (ncu: profiles/r01_ncu_summary.txt, DESIGN.md section 4)"}
        path_bytes = alg["easu"] + alg["rcas"]
        kernels["path"] = {"algorithmic_bytes": path_bytes, "us": ms / K * 1e3,
                           "GBps": path_bytes / (ms / K * 1e-3) / 1e9, "frac_of_hbm_peak": path_bytes / (ms / K * 1e-3) / 1e9 / peak}

    # ---- end to end: host (pinned) -> device -> kernels -> host, through the C ABI's host-frame entry point
    e2e = None
    if world == 1:
        NS = 3
        ctxs = [api.HostContext(iw, ih, ow, oh, api.FORMAT_RGBA16F if dts == "f16" else api.FORMAT_RGBA32F) for _ in range(NS)]
        streams = [torch.cuda.Stream(device=dev) for _ in range(NS)]
        hin = [torch.from_numpy(host_frame(777 + t)).pin_memory() for t in range(NS)]
        hout = [torch.empty((oh, ow, 4), dtype=tdt).pin_memory() for _ in range(NS)]
        Ke = max(3, min(K, 60))

        def e2e_step(i):
            j = i % NS
            ctxs[j].upscale_host(hin[j], hout[j], SHARPNESS, stream=streams[j])
        for i in range(NS * 2):
            e2e_step(i)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(Ke):
            e2e_step(i)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        e2e = {"value": ow * oh * Ke / dt / 1e6, "unit": "Mpix/s", "h2d_bytes_per_step": iw * ih * bpp,
               "d2h_bytes_per_step": ow * oh * bpp, "steps": Ke, "ms_per_step": dt / Ke * 1e3,
               "how": "fsr1_context_upscale_host on pinned host frames, %d streams round-robin" % NS}
        checks = float(hout[0][::97, ::89, :3].float().sum())  # the result is really read on the host
        e2e["host_checksum"] = checks
        for c in ctxs:
            c.close()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = time_cpu(wl, steps=12, warmup=1, budget_s=15.0)
        cpu.pop("seconds", None)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "Mpix/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "strong" if (world > 1 and args.shard_frame) else "weak", "vs_baseline": None,
            "dtype": "f16" if dts == "f16" else "f32", "data": "synthetic",
            "config": {"workload": wl[5] if world == 1 else wl[5] + (" cut into %d row slabs, NCCL halo" % world if args.shard_frame else " x%d slabs tall, row-slab sharded, NCCL halo" % world),
                       "sharpness_stops": SHARPNESS, "frame": "LCG uniform noise, seed 12345+t",
                       "l2": "ring of %d frame sets (%.0f MB per rank) > 126 MB L2" % (RING, RING * (iw * ih + 2 * ow * oh) * bpp / 1e6),
                       "parallelism": "1 GPU" if world == 1 else "row-slab x%d, %d B halo recv per rank per step, %s" % (
                           world, halo, "one CUDA graph per frame" if args.graph else ("halo exchange in line" if args.no_overlap else (
                               "EXPERIMENTAL: halo exchange %s%s" % ("of %d frames per NCCL group" % args.halo_batch if args.halo_batch > 1 else "%d frames ahead" % args.halo_depth,
                                                                      ", RCAS/EASU of consecutive frames on two streams" if args.pipeline_sharded else "")
                               if (args.halo_depth > 1 or args.pipeline_sharded or args.halo_batch > 1) else "halo exchange of frame i+1 overlapped with frame i on a second stream")))},
            "gpu_launches": int(launches), "clocks": clocks,
        }
        if world == 1:
            line["config"]["pipelining"] = ("two CUDA streams: RCAS of frame i overlaps EASU of frame i+1 (api.FramePipeline)"
                                            if piped else "none: EASU then RCAS of each frame on one stream")
            if latency_ms is not None:
                line["unpipelined_ms_per_step"] = latency_ms
                line["unpipelined_value"] = total_out_px / (latency_ms * 1e-3) / 1e6
        if roofline:
            line["roofline"] = roofline
            line["kernels"] = kernels
        if cpu:
            line["cpu_baseline"] = cpu
        if e2e:
            line["e2e"] = e2e
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def run_reference(args, rank):
    if rank != 0:
        return
    # torchrun exports OMP_NUM_THREADS=1 to every rank; the reference arm is meant to use all host threads it can,
    # so restore the default BEFORE the OpenMP runtime of the oracle libraries is loaded
    if os.environ.get("WORLD_SIZE") and os.environ.get("OMP_NUM_THREADS") == "1":
        os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
    wl = WORKLOADS[args.workload]
    r = time_cpu(wl, steps=args.steps, warmup=max(args.warmup, 1), budget_s=100.0)
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": "Mpix/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": max(args.warmup, 1), "ms_per_step": r["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl[5], "sharpness_stops": SHARPNESS,
                       "note": "the reference ships no CPU pixel path; this is its own FsrEasuF/FsrRcasF source compiled "
                               "for the host (oracle/_ref) or, if that was not built, the oracle port"},
            "cpu_baseline": {"value": r["value"], "unit": "Mpix/s", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]},
            "e2e": {"value": r["value"], "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="1080p-4k-fp16", choices=list(WORKLOADS))
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-pipeline", action="store_true", help="N=1: run EASU and RCAS of each frame back to back on one stream (no frame overlap)")
    ap.add_argument("--frames", type=int, default=0, help="distinct synthetic frames resident in HBM (default 8; BASELINE configs[2] uses 120)")
    ap.add_argument("--shard-frame", action="store_true", help="multi-GPU: shard the workload's own frame by rows (strong scaling) instead of stacking one frame per rank")
    ap.add_argument("--halo-depth", type=int, default=1, help="multi-GPU (experimental): exchange halos this many frames ahead (default 1 = the measured configuration)")
    ap.add_argument("--halo-batch", type=int, default=1, help="multi-GPU (experimental): exchange the halos of this many consecutive frames in one NCCL group (<= ring/2)")
    ap.add_argument("--pipeline-sharded", action="store_true", help="multi-GPU (experimental): overlap RCAS of frame i with EASU of frame i+1 on two streams, as at N=1")
    ap.add_argument("--graph", action="store_true", help="multi-GPU (experimental): replay one CUDA graph per frame (NCCL send/recv + kernels)")
    ap.add_argument("--no-overlap", action="store_true", help="multi-GPU: exchange halos in line with the kernels instead of one frame ahead")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, rank)
    if args.gpus > 1 and world == 1:
        # launched without torchrun: re-exec under it (one rank per GPU)
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", str(29400 + os.getpid() % 500), os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
