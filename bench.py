#!/usr/bin/env python
"""bench.py — throughput of the vox_box hot path on B200 (BASELINE.json metric: frames/sec).

Default workload = BASELINE.json configs[1] ("C2"): LPC order-12 autocorrelation + Levinson on 1 h of
synthetic 16 kHz audio per GPU, 25 ms / 10 ms frames (N=400, hop=160, symmetric Hann), 360 utterances x
10 s = 359 280 frames.  A "step" is one pass of the hot path over that batch.  One process per GPU; the
path shards by utterance with no data-path collective (weak scaling: every rank owns its own hour).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config c2|c3|c4|c5]

Prints ONE JSON line (rank 0).  `value` = device-resident throughput (CUDA events on the library's
stream, max over ranks); `e2e` = the same work through the host-pointer C-ABI call (pinned host buffers,
H2D + kernels + D2H inside the timed region); `roofline` = the dominant kernel against its bounding pipe;
`cpu_baseline` = the CPU oracle (a C++ port of the Rust reference — no Rust toolchain in this image)
timed on this box's host cores.  `--impl reference` times that CPU port alone.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "vox_box.rs_b200", "python"))
sys.path.insert(0, ROOT)

CONFIGS = {
    # name: fs, N, hop, utterance seconds, utterances per GPU, description
    "c2": dict(fs=16000, n=400, hop=160, seconds=10.0, utts=360, p=12,
               workload="C2: LPC-12 (Hann -> autocorrelate(13) fp64 -> Levinson) on 1 h synthetic 16 kHz audio, "
                        "N=400 hop=160, 360 utt x 998 frames = 359280 frames per GPU"),
}


def _peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return None


class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU while the timed regions run (NVML; falls back to
    the nvidia-smi query line of /opt/skills/guides/B200_PROFILING.md)."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x10: "sync_boost"}

    def __init__(self, index):
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop = threading.Event()
        self._t = None
        self.how = None

    def _run_nvml(self):
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
        self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
        self.how = "nvml"
        while not self._stop.is_set():
            self.samples.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
            mask = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h) if hasattr(
                pynvml, "nvmlDeviceGetCurrentClocksEventReasons") else pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
            for bit, name in self.REASONS.items():
                if mask & bit:
                    self.reasons.add(name)
            time.sleep(0.002)

    def _run_smi(self):
        import subprocess
        self.how = "nvidia-smi"
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        p = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                              "-lms", "100"], stdout=subprocess.PIPE, text=True)
        try:
            while not self._stop.is_set():
                line = p.stdout.readline()
                if not line:
                    break
                f = [x.strip() for x in line.split(",")]
                self.samples.append(int(f[0]))
                self.max_mhz = int(f[1])
                for nm, v in zip(names, f[2:]):
                    if v.lower().startswith("active"):
                        self.reasons.add(nm)
        finally:
            p.terminate()

    def _run(self):
        try:
            self._run_nvml()
        except Exception:
            try:
                self._run_smi()
            except Exception:
                self.how = "unavailable"

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=5)

    def summary(self):
        s = sorted(self.samples)
        # "under load": drop samples below half the maximum seen (idle gaps between regions)
        hot = [x for x in s if x >= 0.5 * s[-1]] if s else []
        return {"sm_mhz": hot[len(hot) // 2] if hot else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s), "how": self.how}


def make_corpus(cfg, rank):
    from voxbox_b200 import synth
    return synth.corpus(cfg["utts"], cfg["fs"], cfg["seconds"], first=rank * cfg["utts"])


def run_reference(args, cfg, rank, world):
    """CPU arm: the oracle port of the reference's per-frame loop, frame-parallel over all host threads
    (the "rayon wrapper" stand-in), on this config.  Rank 0 only."""
    if rank != 0:
        return
    import oracle
    oracle.build()
    audio = make_corpus(cfg, 0)
    n_samp = audio.shape[1]
    J = oracle.n_frames_of(n_samp, cfg["n"], cfg["hop"])
    # bounded sample: 36 utterances (1/10 of the hour) per step, all host threads
    sample_utts = min(cfg["utts"], 36)
    threads = oracle.max_threads()

    def step():
        for u in range(sample_utts):
            oracle.batch_lpc(audio[u], J, cfg["n"], cfg["hop"], oracle.WIN_HANN_SYMMETRIC, cfg["p"], n_threads=0)

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    frames = sample_utts * J * args.steps
    value = frames / dt
    line = {
        "impl": "reference", "metric": "LPC-12 frames/sec (autocorrelation + Levinson)", "value": value,
        "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": cfg["workload"], "sample": f"{sample_utts} of {cfg['utts']} utterances per step"},
        "cpu_baseline": {"value": value, "unit": "frames/s", "cores": threads, "kind": "port",
                         "sample": f"{sample_utts} utterances x {J} frames per step, {args.steps} steps, OpenMP over frames"},
        "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def cpu_baseline(cfg, audio):
    """Bounded CPU sample on rank 0 (N=1 only): oracle port, single thread as shipped + all threads."""
    import oracle
    oracle.build()
    n_samp = audio.shape[1]
    J = oracle.n_frames_of(n_samp, cfg["n"], cfg["hop"])
    threads = oracle.max_threads()
    one_utts, all_utts = 20, min(cfg["utts"], 20 * max(1, threads))
    t0 = time.perf_counter()
    for u in range(one_utts):
        oracle.batch_lpc(audio[u], J, cfg["n"], cfg["hop"], oracle.WIN_HANN_SYMMETRIC, cfg["p"], n_threads=1)
    t1 = time.perf_counter() - t0
    reps = 0
    t0 = time.perf_counter()
    while True:
        for u in range(all_utts):
            oracle.batch_lpc(audio[u], J, cfg["n"], cfg["hop"], oracle.WIN_HANN_SYMMETRIC, cfg["p"], n_threads=0)
        reps += 1
        if time.perf_counter() - t0 > 3.0 or reps >= 50:
            break
    tn = time.perf_counter() - t0
    return {"value": all_utts * J * reps / tn, "unit": "frames/s", "cores": threads, "kind": "port",
            "single_thread_value": one_utts * J / t1,
            "sample": f"C++ f64 port of the Rust reference (no cargo here): {all_utts} utterances x {J} frames x {reps} "
                      f"passes on {threads} OpenMP threads; single thread: {one_utts} utterances"}


def run_ours(args, cfg, rank, world, local_rank):
    import ctypes as C

    import voxbox_b200 as vb
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()

    def max_over_ranks(x):
        if dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    ctx = vb.Context(local_rank)
    audio = make_corpus(cfg, rank)
    U, n_samp = audio.shape
    N, hop, p = cfg["n"], cfg["hop"], cfg["p"]
    J = ctx.n_frames_of(n_samp, N, hop)
    F = U * J
    d_audio = ctx.to_device(audio)
    d_r = ctx.empty((F, p + 1), np.float32)
    d_ac = ctx.empty((F, p + 1), np.float32)
    fr = ctx.frames(d_audio.ptr, F, N, hop, vb.WINDOW_HANN_SYMMETRIC, frames_per_segment=J, segment_stride=n_samp)

    def step():
        ctx._check(ctx.lib.vbx_lpc(ctx.h, C.byref(fr), p, d_r.ptr, d_ac.ptr, None, vb.F32), "vbx_lpc")

    peaks = ctx.measure_peaks()  # FP32/FP64 FMA pipe peaks of this device (roofline denominators)
    with ClockSampler(local_rank) as clocks:
        # ---- device-resident throughput -------------------------------------------------------
        for _ in range(max(args.warmup, 3)):
            step()
        ctx.sync()
        barrier()
        l0 = ctx.kernel_launches
        ctx.timer_start()
        for _ in range(args.steps):
            step()
        ms = ctx.timer_stop_ms()
        launches = ctx.kernel_launches - l0
        barrier()
        ms = max_over_ranks(ms)

        # ---- end to end through the host-pointer C-ABI call ------------------------------------------
        h_in = C.c_void_p()
        ctx._check(ctx.lib.vbx_malloc_host(ctx.h, audio.nbytes, C.byref(h_in)), "vbx_malloc_host")
        h_audio = np.ctypeslib.as_array(C.cast(h_in, C.POINTER(C.c_float)), shape=audio.shape)
        h_audio[...] = audio
        out_bytes = F * (p + 1) * 4
        h_out = [C.c_void_p(), C.c_void_p()]
        for h in h_out:
            ctx._check(ctx.lib.vbx_malloc_host(ctx.h, out_bytes, C.byref(h)), "vbx_malloc_host")
        hfr = ctx.frames(h_in.value, F, N, hop, vb.WINDOW_HANN_SYMMETRIC, frames_per_segment=J, segment_stride=n_samp)

        def e2e_step():
            ctx._check(ctx.lib.vbx_lpc_host(ctx.h, C.byref(hfr), p, h_out[0], h_out[1], None, vb.F32), "vbx_lpc_host")

        e2e_steps = max(3, min(args.steps, 20))
        for _ in range(3):
            e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        e2e_s = max_over_ranks(time.perf_counter() - t0)
        barrier()
    h_r = np.ctypeslib.as_array(C.cast(h_out[0], C.POINTER(C.c_float)), shape=(F, p + 1))
    checksum = float(np.sum(h_r[:: max(1, F // 1000), 0], dtype=np.float64))

    if rank == 0:
        total_frames = F * world
        value = total_frames * args.steps / (ms * 1e-3)
        # roofline of the dominant (only) kernel, lpc_fused_kernel<13,float>: FP64-pipe bound.
        # Algorithmic work per frame: 2·13·400 lag MACs + 400 window + ~350 Levinson = 11 150 flop,
        # 744 B (4·hop in, 2·13·4 out).
        flop_per_frame = 2 * (p + 1) * N + N + 350
        bytes_per_frame = 4 * hop + 2 * 4 * (p + 1)
        kernel_s = ms * 1e-3 / args.steps
        ach_tf = flop_per_frame * F / kernel_s / 1e12
        ach_gb = bytes_per_frame * F / kernel_s / 1e9
        mp = _peaks()
        hbm_peak = mp["hbm_gbs"] if mp else 6650.0
        line = {
            "metric": "LPC-12 frames/sec (autocorrelation + Levinson)", "value": value, "unit": "frames/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": cfg["workload"], "frames_per_gpu": F, "order": p, "window": "hann_symmetric",
                       "outputs": "r[13], ac[13] fp32", "l2": "inputs larger than L2 (230 MB audio per step), no flush",
                       "parallelism": f"utterance-sharded x{world}, no collective"},
            "e2e": {"value": total_frames * e2e_steps / e2e_s, "unit": "frames/s", "h2d_bytes_per_step": int(audio.nbytes),
                    "d2h_bytes_per_step": int(2 * out_bytes), "steps": e2e_steps, "api": "vbx_lpc_host (pinned host buffers)"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "fp64", "achieved": ach_tf, "peak": peaks["fp64_tflops"], "unit": "TFLOP/s",
                         "frac": ach_tf / peaks["fp64_tflops"], "traffic": None,
                         "kernel": "lpc_fused_kernel<13,float>", "flop_per_frame": flop_per_frame,
                         "peak_source": "vbx_measure_peaks: DFMA loop on this device, this run"},
            "roofline_hbm": {"bound": "hbm", "achieved": ach_gb, "peak": hbm_peak, "unit": "GB/s",
                             "frac": ach_gb / hbm_peak, "bytes_per_frame": bytes_per_frame,
                             "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if mp else "fallback 6650 (of fallback)"},
            "pipe_peaks": peaks,
            "clocks": clocks.summary(),
            "checksum": checksum,
        }
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_baseline(cfg, audio)
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()
    ctx.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS))
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cfg = CONFIGS[args.config]
    if args.impl == "reference":
        run_reference(args, cfg, rank, world)
    else:
        run_ours(args, cfg, rank, world, local_rank)


if __name__ == "__main__":
    main()
