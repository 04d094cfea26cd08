#!/usr/bin/env python
"""bench.py — throughput of the vox_box hot path on B200 (BASELINE.json metric: frames/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config c2|c3|c4|c5] [--utts U]

Default workload = BASELINE.json configs[1] ("C2"): LPC order-12 autocorrelation + Levinson on 1 h of
synthetic 16 kHz audio per GPU, 25 ms / 10 ms frames (N=400, hop=160, symmetric Hann), 360 utterances x
10 s = 359 280 frames.  A "step" is one pass of the hot path over that batch.  The other configs
(c3 formants, c4 pitch, c5 MFCC) are the remaining BASELINE.json shapes, on a per-GPU batch that fits
a quick run (stated in config.workload).  One process per GPU; the path shards by utterance with no
data-path collective (weak scaling: every rank owns its own batch; torch.distributed/NCCL is used only
for the barrier and the max-over-ranks of the timing).

Prints ONE JSON line (rank 0).  `value` = device-resident throughput (CUDA events on the library's
stream, max over ranks); `e2e` = the same work through the host-pointer C-ABI call (pinned host buffers,
H2D + kernels + D2H inside the timed region); `roofline` = the dominant kernel against the pipe that
bounds it (measured on this device in this run) and `roofline_hbm` the same against MEASURED_PEAKS.json's
HBM number; `cpu_baseline` = the CPU oracle (a C++ port of the Rust reference — there is no Rust
toolchain in this image) timed on this box's host cores.  `--impl reference` times that CPU port alone.
"""
import argparse
import ctypes as C
import json
import os
import re
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "vox_box.rs_b200", "python"))
sys.path.insert(0, ROOT)

MALE = (320., 1440., 2760., 3200.)  # lib.rs:27

CONFIGS = {
    "c2": dict(kind="lpc", fs=16000, n=400, hop=160, seconds=10.0, utts=360, distinct=360, p=12,
               metric="LPC-12 frames/sec (autocorrelation + Levinson)", dtype="f64",
               workload="C2: LPC-12 (Hann -> autocorrelate(13) fp64 -> Levinson) on 1 h synthetic 16 kHz audio, "
                        "N=400 hop=160, {utts} utt x {J} frames = {F} frames per GPU"),
    "c3": dict(kind="formants", fs=44100, n=1102, hop=441, seconds=10.0, utts=4500, distinct=24, p=12,
               metric="LPC-12 + formant frames/sec (autocorrelation + Levinson -> Laguerre roots -> McCandless)", dtype="f64",
               workload="C3: formant extraction (Hann -> autocorrelate(13) -> Levinson -> Laguerre roots -> resonances -> "
                        "McCandless tracker) on synthetic 44.1 kHz speech, N=1102 hop=441, {utts} utt x {J} frames = {F} frames "
                        "per GPU per step (4500 utterances = one GPU's share of 100 h over 8 GPUs; "
                        "{distinct} distinct utterances tiled)"),
    "c4": dict(kind="pitch", fs=16000, n=640, hop=160, seconds=10.0, utts=360, distinct=48,
               metric="Boersma pitch frames/sec (75-600 Hz)", dtype="f32 lag sweep + f64 refinement",
               workload="C4: Boersma pitch 75-600 Hz (Hann -> all-lag autocorrelation -> candidates -> Brent/sinc refinement) "
                        "on synthetic 16 kHz audio, N=640 hop=160, {utts} utt x {J} frames = {F} frames per GPU per step "
                        "(1 h of the 1000 h corpus per step; {distinct} distinct utterances tiled)"),
    "c5": dict(kind="mfcc", fs=16000, n=400, hop=160, seconds=10.0, utts=3600, distinct=48,
               metric="MFCC frames/sec (13 coefficients from 40 mel bands)", dtype="f64",
               workload="C5: MFCC (Hann -> FFT -> 40 mel bands -> log10 -> DCT, 13 kept) on synthetic 16 kHz audio, N=400 "
                        "hop=160, {utts} utt x {J} frames = {F} frames per GPU per step ({distinct} distinct utterances tiled)"),
}


def _peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return None


def _ncu_traffic(profile_name):
    """dram read+write bytes per launch from a committed `ncu --set full` summary under profiles/ (or None)."""
    try:
        txt = open(os.path.join(ROOT, "profiles", profile_name)).read()
        tot = 0.0
        for key in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            m = re.search(key + r" = ([0-9.]+) (\w+)", txt)
            tot += float(m.group(1)) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[m.group(2)]
        return tot
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
    """[utts, n_samples] fp32: `distinct` synthetic utterances (seeded by rank) tiled to `utts` rows."""
    from voxbox_b200 import synth
    distinct = min(cfg["distinct"], cfg["utts"])
    base = synth.corpus(distinct, cfg["fs"], cfg["seconds"], first=rank * distinct)
    reps = -(-cfg["utts"] // distinct)
    return np.ascontiguousarray(np.tile(base, (reps, 1))[:cfg["utts"]])


# ------------------------------------------------------------------------------------------------
# CPU port (oracle) of each config: used by --impl reference and by the cpu_baseline leg
# ------------------------------------------------------------------------------------------------
def numa_bind(gpu_index):
    """Bind the calling thread to the CPUs NVML reports as local to the GPU; returns the previous mask (or None)."""
    try:
        import pynvml
        before = os.sched_getaffinity(0)
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        n_words = (os.cpu_count() + 63) // 64
        words = pynvml.nvmlDeviceGetCpuAffinity(h, n_words)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1}
        cpus &= before
        if cpus and cpus != before:
            os.sched_setaffinity(0, cpus)
            return before
    except Exception:
        pass
    return None


def host_threads():
    """Host threads for the CPU arm: every core this process may run on (torchrun exports OMP_NUM_THREADS=1, which
    would silently make the "all host threads" baseline single-threaded, so the count is passed explicitly)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def cpu_pass(cfg, audio_rows, J, n_threads):
    """One pass of the reference's per-frame loop over the given utterances on the CPU oracle (n_threads: 1 = as
    shipped, >1 = OpenMP over frames / utterances)."""
    import oracle
    N, hop, fs = cfg["n"], cfg["hop"], float(cfg["fs"])
    kind = cfg["kind"]
    for row in audio_rows:
        if kind == "lpc":
            oracle.batch_lpc(row, J, N, hop, oracle.WIN_HANN_SYMMETRIC, cfg["p"], n_threads=n_threads)
        elif kind == "pitch":
            oracle.batch_pitch(row, J, N, hop, oracle.WIN_HANN_SYMMETRIC, fs, 0.45, 75.0, 600.0, 16, n_threads=n_threads)
        elif kind == "mfcc":
            oracle.batch_mfcc(row, J, N, hop, oracle.WIN_HANN_SYMMETRIC, 40, 133.0, 6855.0, fs, n_keep=13, n_threads=n_threads)
    if kind == "formants":
        # utterances in parallel (the tracker is sequential inside an utterance); frame f of the flattened batch
        # starts at f*hop, so utterance u owns frames [u*ns/hop, u*ns/hop + J) when hop divides the utterance length
        rows = np.ascontiguousarray(np.stack(audio_rows))
        U, ns = rows.shape
        est = np.array([[f, 1.0] for f in MALE])
        if ns % hop == 0:
            # ranges [start_u, start_u + J) are the utterances; the 1-2 frames that straddle two utterances in between
            # are walked as (negligible) extra ranges because the oracle's batch loop takes contiguous frame ranges
            starts = np.arange(U, dtype=np.int64) * (ns // hop)
            offs = np.stack([starts, starts + J], axis=1).reshape(-1)
            oracle.batch_formants(rows.reshape(-1), int(offs[-1]), N, hop, oracle.WIN_HANN_SYMMETRIC, 1, fs, cfg["p"], offs, est,
                                  n_threads=n_threads)
        else:
            for u in range(U):
                oracle.batch_formants(rows[u], J, N, hop, oracle.WIN_HANN_SYMMETRIC, 1, fs, cfg["p"],
                                      np.array([0, J], dtype=np.int64), est, n_threads=n_threads)


def cpu_sample_utts(cfg, threads):
    """Utterances per bounded CPU sample (about 3-10 s of CPU work on `threads` cores)."""
    per_thread = {"lpc": 20, "formants": 2, "pitch": 1, "mfcc": 6}[cfg["kind"]]
    return max(1, min(cfg["utts"], per_thread * max(1, threads)))


def run_reference(args, cfg, rank, world):
    """CPU arm: the oracle port of the reference's per-frame loop, frame-parallel over all host threads
    (the "rayon wrapper" stand-in), on this config.  Rank 0 only."""
    if rank != 0:
        return
    import oracle
    oracle.build()
    threads = host_threads()
    sub = dict(cfg)
    sub["utts"] = sub["distinct"] = cpu_sample_utts(cfg, threads)
    audio = make_corpus(sub, 0)
    J = oracle.n_frames_of(audio.shape[1], cfg["n"], cfg["hop"])
    rows = list(audio)

    for _ in range(args.warmup):
        cpu_pass(cfg, rows, J, threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_pass(cfg, rows, J, threads)
    dt = time.perf_counter() - t0
    frames = len(rows) * J * args.steps
    value = frames / dt
    full = dict(cfg)
    Jf = J
    line = {
        "impl": "reference", "metric": cfg["metric"], "value": value,
        "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": cfg["workload"].format(utts=full["utts"], J=Jf, F=full["utts"] * Jf, distinct=full["distinct"]),
                   "sample": f"{len(rows)} of {cfg['utts']} utterances per step"},
        "cpu_baseline": {"value": value, "unit": "frames/s", "cores": threads, "kind": "port",
                         "sample": f"C++ f64 port of the Rust reference (no cargo in this image): {len(rows)} utterances x {J} "
                                   f"frames per step, {args.steps} steps, OpenMP over frames on {threads} threads"},
        "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def cpu_baseline(cfg, rank_audio):
    """Bounded CPU sample on rank 0 (N=1 only): oracle port, single thread as shipped + all threads."""
    import oracle
    oracle.build()
    J = oracle.n_frames_of(rank_audio.shape[1], cfg["n"], cfg["hop"])
    threads = host_threads()
    n_all = min(cpu_sample_utts(cfg, threads), rank_audio.shape[0])
    n_one = max(1, min(n_all, {"lpc": 20, "formants": 3, "pitch": 1, "mfcc": 6}[cfg["kind"]]))
    rows = list(rank_audio[:n_all])
    t0 = time.perf_counter()
    cpu_pass(cfg, rows[:n_one], J, 1)
    t1 = time.perf_counter() - t0
    reps = 0
    t0 = time.perf_counter()
    while True:
        cpu_pass(cfg, rows, J, threads)
        reps += 1
        if time.perf_counter() - t0 > 3.0 or reps >= 50:
            break
    tn = time.perf_counter() - t0
    return {"value": n_all * J * reps / tn, "unit": "frames/s", "cores": threads, "kind": "port",
            "single_thread_value": n_one * J / t1,
            "sample": f"C++ f64 port of the Rust reference (no cargo here): {n_all} utterances x {J} frames x {reps} "
                      f"passes on {threads} OpenMP threads; single thread: {n_one} utterances"}


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
class Workload:
    """Device buffers + the device-pointer step and the host-pointer (e2e) step of one config."""

    def __init__(self, ctx, vb, cfg, audio):
        self.ctx, self.vb, self.cfg = ctx, vb, cfg
        L, h = ctx.lib, ctx.h
        U, ns = audio.shape
        N, hop, fs = cfg["n"], cfg["hop"], float(cfg["fs"])
        J = ctx.n_frames_of(ns, N, hop)
        F = U * J
        self.U, self.J, self.F, self.ns = U, J, F, ns
        self.d_audio = ctx.to_device(audio)
        win = vb.WINDOW_HANN_SYMMETRIC
        fr = ctx.frames(self.d_audio.ptr, F, N, hop, win, frames_per_segment=J, segment_stride=ns)
        self.h_in = C.c_void_p()
        ctx._check(L.vbx_malloc_host(h, audio.nbytes, C.byref(self.h_in)), "vbx_malloc_host")
        np.ctypeslib.as_array(C.cast(self.h_in, C.POINTER(C.c_float)), shape=audio.shape)[...] = audio
        hfr = ctx.frames(self.h_in.value, F, N, hop, win, frames_per_segment=J, segment_stride=ns)
        self.h2d = int(audio.nbytes)
        # the same audio as 16-bit PCM (the format the reference's WAV drivers read, tests/lib.rs:15-19): half the H2D bytes
        pcm = np.clip(np.round(audio.astype(np.float64) * 32767.0), -32768, 32767).astype(np.int16)
        self.h_in16 = C.c_void_p()
        ctx._check(L.vbx_malloc_host(h, pcm.nbytes, C.byref(self.h_in16)), "vbx_malloc_host")
        np.ctypeslib.as_array(C.cast(self.h_in16, C.POINTER(C.c_int16)), shape=pcm.shape)[...] = pcm
        hfr16 = ctx.frames(self.h_in16.value, F, N, hop, win, dtype=vb.I16, frames_per_segment=J, segment_stride=ns)
        self.h2d16 = int(pcm.nbytes)
        kind = cfg["kind"]

        def pinned(nbytes):
            p = C.c_void_p()
            ctx._check(L.vbx_malloc_host(h, nbytes, C.byref(p)), "vbx_malloc_host")
            return p

        if kind == "lpc":
            p = cfg["p"]
            d_r, d_ac = ctx.empty((F, p + 1), np.float32), ctx.empty((F, p + 1), np.float32)
            out_bytes = F * (p + 1) * 4
            h_r, h_ac = pinned(out_bytes), pinned(out_bytes)
            self.step = lambda: ctx._check(L.vbx_lpc(h, C.byref(fr), p, d_r.ptr, d_ac.ptr, None, vb.F32), "vbx_lpc")
            self.e2e_step = lambda fr_=hfr: ctx._check(L.vbx_lpc_host(h, C.byref(fr_), p, h_r, h_ac, None, vb.F32), "vbx_lpc_host")
            self.d2h = 2 * out_bytes
            self.api = "vbx_lpc_host (pinned host buffers)"
            self.check = lambda: float(np.sum(np.ctypeslib.as_array(C.cast(h_r, C.POINTER(C.c_float)), shape=(F, p + 1))[::max(1, F // 1000), 0], dtype=np.float64))
            self.outputs = "r[13], ac[13] fp32"
            self._keep = (d_r, d_ac)
        elif kind == "formants":
            p = cfg["p"]
            est0 = np.tile(np.array([[f, 1.0] for f in MALE], dtype=np.float32), (U, 1, 1))
            d_est0 = ctx.to_device(est0)
            d_est = ctx.empty((U, 4, 2), np.float32)
            d_trk = ctx.empty((F, 4, 2), np.float32)
            trk_bytes = F * 4 * 2 * 4
            h_trk, h_est = pinned(trk_bytes), pinned(est0.nbytes)

            def step():
                # the tracker state is in/out: every step starts from the MALE estimates (tests/lib.rs:36,60)
                ctx._check(L.vbx_memcpy_d2d(h, d_est.ptr, d_est0.ptr, est0.nbytes), "vbx_memcpy_d2d")
                ctx._check(L.vbx_find_formants(h, C.byref(fr), fs, p, vb.LPC_AUTOCORR, d_est.ptr, 4, d_trk.ptr, None, None, None,
                                               vb.F32), "vbx_find_formants")

            def e2e_step(fr_=hfr):
                C.memmove(h_est, est0.ctypes.data, est0.nbytes)
                ctx._check(L.vbx_find_formants_host(h, C.byref(fr_), fs, p, vb.LPC_AUTOCORR, h_est, 4, h_trk, None, None, None,
                                                    vb.F32), "vbx_find_formants_host")

            self.step, self.e2e_step = step, e2e_step
            self.h2d += int(est0.nbytes)
            self.d2h = trk_bytes + int(est0.nbytes)
            self.api = "vbx_find_formants_host (pinned host buffers)"
            self.check = lambda: float(np.sum(np.ctypeslib.as_array(C.cast(h_trk, C.POINTER(C.c_float)), shape=(F, 8))[::max(1, F // 1000)], dtype=np.float64))
            self.outputs = "4 formant (frequency, bandwidth) tracks per frame, fp32"
            self._keep = (d_est0, d_est, d_trk, est0)
        elif kind == "pitch":
            K = 16
            d_c, d_n, d_s = ctx.empty((F, K, 2), np.float32), ctx.empty((F,), np.int32), ctx.empty((F,), np.uint8)
            h_c, h_n, h_s = pinned(F * K * 8), pinned(F * 4), pinned(F)
            self.step = lambda: ctx._check(L.vbx_pitch(h, C.byref(fr), fs, 0.45, 75.0, 600.0, K, d_c.ptr, d_n.ptr, d_s.ptr, vb.F32), "vbx_pitch")
            self.e2e_step = lambda fr_=hfr: ctx._check(L.vbx_pitch_host(h, C.byref(fr_), fs, 0.45, 75.0, 600.0, K, h_c, h_n, h_s, vb.F32), "vbx_pitch_host")
            self.d2h = F * (K * 8 + 5)
            self.api = "vbx_pitch_host (pinned host buffers)"
            self.check = lambda: float(np.sum(np.ctypeslib.as_array(C.cast(h_c, C.POINTER(C.c_float)), shape=(F, K * 2))[::max(1, F // 1000), 0], dtype=np.float64))
            self.outputs = "up to 16 (frequency, strength) candidates per frame fp32 + count + status"
            self._keep = (d_c, d_n, d_s)
        elif kind == "mfcc":
            d_o = ctx.empty((F, 13), np.float32)
            h_o = pinned(F * 13 * 4)
            self.step = lambda: ctx._check(L.vbx_mfcc(h, C.byref(fr), 40, 13, 133.0, 6855.0, fs, d_o.ptr, None, vb.F32), "vbx_mfcc")
            self.e2e_step = lambda fr_=hfr: ctx._check(L.vbx_mfcc_host(h, C.byref(fr_), 40, 13, 133.0, 6855.0, fs, h_o, vb.F32), "vbx_mfcc_host")
            self.d2h = F * 13 * 4
            self.api = "vbx_mfcc_host (pinned host buffers)"
            self.check = lambda: float(np.sum(np.ctypeslib.as_array(C.cast(h_o, C.POINTER(C.c_float)), shape=(F, 13))[::max(1, F // 1000), 0], dtype=np.float64))
            self.outputs = "13 MFCC per frame fp32"
            self._keep = (d_o,)

        self.hfr16 = hfr16

    def kernel_table(self):
        """Algorithmic work per frame of every kernel of the step (SURVEY §8d figures, split per kernel):
        name -> (pipe that bounds it, flop per frame, HBM bytes per frame, committed ncu summary or None)."""
        cfg, N, hop = self.cfg, self.cfg["n"], self.cfg["hop"]
        p = cfg.get("p", 12)
        lpc = ("fp64", 2 * (p + 1) * N + N + 350, 4 * hop + 2 * 4 * (p + 1), "r1_lpc_final_full.txt")
        return {
            # lpc_fused16_kernel: the 16-aligned-framing variant (C2, C5 shapes); lpc_fused_kernel: any other framing (C3)
            "lpc": {"lpc_fused16_kernel": lpc[:3] + ("r1_lpc16_final_full.txt",), "lpc_fused_kernel": lpc},
            "formants": {
                "lpc_fused_kernel": ("fp64", 2 * (p + 1) * N + N + 350, 4 * hop + 8 * (p + 1), "r1_lpc_final_full.txt"),
                "lpc_fused16_kernel": ("fp64", 2 * (p + 1) * N + N + 350, 4 * hop + 8 * (p + 1), "r1_lpc16_final_full.txt"),
                # 10 Laguerre solves x 20 iterations x (3*12 complex FMA*8 + ~60) + polish/resonances ~ 70 k flop (fp32 pipe)
                # (SURVEY §8d counts the REFERENCE's algorithm: one root at a time.  The kernel that runs divides conjugate
                # pairs out and executes roughly half of these flops, so this fraction is work-equivalent, not pipe utilisation:
                # ncu reads 39 % FMA-pipe / 81 % issue utilisation, profiles/r1_roots_final_full.txt.)
                "lpc_roots_rt_kernel": ("fp32", 70e3, 8 * (p + 1) + 8 * p + 5, "r1_roots_final_full.txt"),
                "tracker_idx_kernel": ("fp64", 600.0, 8 * p + 4 + 4 * 8, "r1_tracker_final_full.txt"),
            },
            "pitch": {
                "pitch_lag_kernel": ("fp32", 2.0 * N * (N + 1) / 2, 4 * hop + 8 * N, "r1_lag_final_full.txt"),
                # ~16 candidates x ~26 Brent evaluations x 2(lag+2) terms x ~30 flop
                "pitch_refine_kernel": ("fp64", 2.5e6, 8 * N, "r1_refine_final_full.txt"),
                "pitch_finalize_kernel": ("fp64", 500.0, 16 * 16 + 140, None),
            },
            "mfcc": {"mfcc_kernel": ("fp64", 19.5e3, 4 * hop + 4 * 13, "r1_mfcc_final_full.txt")},
        }[cfg["kind"]]

    def roofline(self, prof, steps, peaks, hbm_peak, hbm_src):
        """Per-kernel roofline from the per-kernel device times measured over the timed region (vbx_profile_*):
        achieved = algorithmic flop (or bytes) per launch / average launch duration.  Returns (dominant kernel's
        roofline object, its HBM view, the per-kernel table)."""
        table = self.kernel_table()
        total_ms = sum(ms for ms, _ in prof.values()) or 1e-9
        kernels = {}
        for name, (ms, n) in sorted(prof.items(), key=lambda kv: -kv[1][0]):
            entry = {"ms_per_step": ms / steps, "launches_per_step": n / steps, "share": ms / total_ms}
            if name in table:
                pipe, flop, byts, _ = table[name]
                sec = ms * 1e-3 / steps
                entry.update({"bound": pipe, "tflops": flop * self.F / sec / 1e12, "frac": flop * self.F / sec / 1e12 / peaks[pipe + "_tflops"],
                              "gbs": byts * self.F / sec / 1e9, "frac_hbm": byts * self.F / sec / 1e9 / hbm_peak})
                if name == "lpc_roots_rt_kernel":
                    entry["note"] = ("flops counted for the reference's one-root-at-a-time algorithm (SURVEY 8d); the conjugate-pair "
                                     "kernel executes about half of them: work-equivalent rate, not pipe utilisation")
            kernels[name] = entry
        dom = next((k for k in kernels if k in table), None)
        if dom is None:
            return None, None, kernels
        pipe, flop, byts, profile = table[dom]
        k = kernels[dom]
        launches = max(1.0, k["launches_per_step"])
        roof = {"bound": pipe, "achieved": k["tflops"], "peak": peaks[pipe + "_tflops"], "unit": "TFLOP/s", "frac": k["frac"],
                "traffic": _ncu_traffic(profile) if profile else None,
                "traffic_note": f"dram read+write of one ncu --set full capture (profiles/{profile}); that capture's batch may be smaller "
                                "than this run's — compare per frame" if profile else None,
                "kernel": dom, "flop_per_frame": flop,
                "flop_per_launch": flop * self.F / launches, "avg_launch_ms": k["ms_per_step"] / launches, "share_of_step": k["share"],
                "peak_source": "vbx_measure_peaks: dependent-free FMA loop on this device, this run",
                "timing": "CUDA events on the library's stream around every launch of the timed region (vbx_profile_begin/end)"}
        roof_hbm = {"bound": "hbm", "achieved": k["gbs"], "peak": hbm_peak, "unit": "GB/s", "frac": k["frac_hbm"],
                    "bytes_per_frame": byts, "kernel": dom, "peak_source": hbm_src}
        return roof, roof_hbm, kernels


def run_ours(args, cfg, rank, world, local_rank):
    import voxbox_b200 as vb
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        # NCCL may print its version banner on stdout while the communicator is created: keep stdout for the ONE JSON line
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            t = torch.zeros(1, device="cuda")
            dist.all_reduce(t)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)

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

    # Run this rank on the CPUs next to its GPU while the pinned host buffers are allocated and the end-to-end leg runs
    # (NUMA-local staging memory: with 8 ranks on a two-socket host the H2D copies otherwise cross the socket link).
    # The original mask is restored before the CPU-baseline leg, which uses every host core.
    cpus_before = numa_bind(local_rank)
    ctx = vb.Context(local_rank)  # raises without the CUDA library / a device: no CPU fallback
    audio = make_corpus(cfg, rank)
    wl = Workload(ctx, vb, cfg, audio)
    F = wl.F
    warmup = max(args.warmup, 3)

    peaks = ctx.measure_peaks()  # FP32/FP64 FMA pipe peaks of this device (roofline denominators)
    with ClockSampler(local_rank) as clocks:
        # ---- device-resident throughput -------------------------------------------------------
        for _ in range(warmup):
            wl.step()
        ctx.sync()
        barrier()
        l0 = ctx.kernel_launches
        ctx.profile_begin()  # an event after every launch: per-kernel device times for the roofline
        ctx.timer_start()
        for _ in range(args.steps):
            wl.step()
        ms = ctx.timer_stop_ms()
        prof = ctx.profile_end()
        launches = ctx.kernel_launches - l0
        barrier()
        ms = max_over_ranks(ms)

        # ---- end to end through the host-pointer C-ABI call ------------------------------------------
        e2e_steps = max(3, min(args.steps, 20 if wl.h2d < (1 << 30) else 5))
        if args.device_only:
            if rank == 0:
                _, _, kernels = wl.roofline(prof, args.steps, peaks, 1.0, "")
                print(json.dumps({"device_only": True, "value": F * world * args.steps / (ms * 1e-3), "unit": "frames/s",
                                  "ms_per_step": ms / args.steps, "gpu_launches": launches,
                                  "kernel_share": {k: v["share"] for k, v in kernels.items()}}))
            return
        for _ in range(3):
            wl.e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            wl.e2e_step()
        e2e_s = max_over_ranks(time.perf_counter() - t0)
        barrier()
        checksum = wl.check()
        # the same call fed with 16-bit PCM host samples (scaled by 1/32767 on load)
        for _ in range(2):
            wl.e2e_step(wl.hfr16)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            wl.e2e_step(wl.hfr16)
        e2e16_s = max_over_ranks(time.perf_counter() - t0)
        barrier()

    if rank == 0:
        total_frames = F * world
        value = total_frames * args.steps / (ms * 1e-3)
        mp = _peaks()
        hbm_peak = mp["hbm_gbs"] if mp else 6650.0
        hbm_src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if mp else "fallback 6650 GB/s (of fallback)"
        roof, roof_hbm, kernels = wl.roofline(prof, args.steps, peaks, hbm_peak, hbm_src)
        in_mb = audio.nbytes / 1e6
        line = {
            "metric": cfg["metric"], "value": value, "unit": "frames/s",
            "n_gpus": world, "steps": args.steps, "warmup": warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": cfg["dtype"], "data": "synthetic",
            "config": {"workload": cfg["workload"].format(utts=wl.U, J=wl.J, F=F, distinct=min(cfg["distinct"], wl.U)),
                       "frames_per_gpu": F, "window": "hann_symmetric", "outputs": wl.outputs,
                       "l2": f"inputs larger than L2 ({in_mb:.0f} MB audio per step), no flush" if in_mb > 130 else
                             f"inputs ({in_mb:.0f} MB) fit L2: compute-bound kernels, no flush",
                       "parallelism": f"utterance-sharded x{world}, no collective"},
            "e2e": {"value": total_frames * e2e_steps / e2e_s, "unit": "frames/s", "h2d_bytes_per_step": wl.h2d,
                    "d2h_bytes_per_step": int(wl.d2h), "steps": e2e_steps, "api": wl.api},
            "e2e_pcm16": {"value": total_frames * e2e_steps / e2e16_s, "unit": "frames/s", "h2d_bytes_per_step": wl.h2d16,
                          "d2h_bytes_per_step": int(wl.d2h), "steps": e2e_steps,
                          "note": "same call, host samples as int16 PCM (vbx_frames.dtype = VBX_I16)"},
            "gpu_launches": int(launches),
            "roofline": roof, "roofline_hbm": roof_hbm, "kernels": kernels, "pipe_peaks": peaks,
            "clocks": clocks.summary(),
            "checksum": checksum,
        }
        if cpus_before:
            os.sched_setaffinity(0, cpus_before)
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_baseline(cfg, audio)
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()
    ctx.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS))
    ap.add_argument("--utts", type=int, default=None, help="utterances per GPU per step (default: the config's)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--device-only", action="store_true",
                    help="device-resident steps only (no e2e / cpu legs): the command the ncu launch lists under profiles/ are taken with")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cfg = dict(CONFIGS[args.config])
    if args.utts:
        cfg["utts"] = args.utts
    if args.steps is None:
        args.steps = {"lpc": 200, "formants": 10, "pitch": 5, "mfcc": 20}[cfg["kind"]] if args.impl == "ours" else 3
    if args.impl == "reference":
        run_reference(args, cfg, rank, world)
    else:
        run_ours(args, cfg, rank, world, local_rank)


if __name__ == "__main__":
    main()
