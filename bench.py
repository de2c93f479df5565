#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 delay-Doppler hot path.

Workload (BASELINE.json configs[1]): one "step" = one 1 s CPI @ 2 MS/s through
WienerHopf (410 taps) -> Ambiguity (300 delay x 257 Doppler) -> Map::set_metrics ->
CfarDetector1D -> Centroid -> Interpolate on synthetic IQ (blah2_b200/scene.py).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

Contract (one JSON line on rank 0):
  value   IQ Msamples/s, whole job over all N GPUs, inputs resident in HBM (float2), CUDA-event
          timed on the launching stream, max over ranks;  maps_per_s = value / (N_samples/1e6).
  e2e     same metric through the public host API (Pipeline.submit_host / fetch = C ABI
          b200dd_pipeline_submit_host / _fetch, PINNED complex128 host buffers): H2D of x and y and D2H of
          the map + detections inside the timed region, every step.
  roofline  the CAF range-correlation kernel (the kernel BASELINE.json's metric names), timed live
          with CUDA events around that kernel (b200dd_caf_profile_device), algorithmic bytes per
          launch = 16 N_used + 8 nDop nDel, against MEASURED_PEAKS.json hbm_gbs.
  cpu_baseline  the reference's own src/process code (oracle/_ref, unmodified sources + our FFT /
          Armadillo shims) on ONE host thread for ONE CPI of the same workload.
  --impl reference: the same reference code, one host process per concurrent CPI (time-bounded).
Multi-GPU: independent CPIs sharded over ranks ("weak" scaling, no data-path collective); the only
collective is the final NCCL gather of every rank's last map to rank 0 (inside the timed region).
"""
from __future__ import annotations

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

FS = 2_000_000
N = 2_000_000
GEOM = dict(delayMin=0, delayMax=299, dopplerMin=-128, dopplerMax=128, fs=FS, nSamples=N, roundHamming=True)
CLUTTER = (-10, 400)
DET = dict(pfa=1e-5, nGuard=2, nTrain=6, minDelay=5, minDoppler=15.0, nCentroid=6)
WORKLOAD = ("cfg2: WienerHopf(410 taps)+Ambiguity(300 delay x 257 Doppler)+set_metrics+CFAR/Centroid/Interpolate, "
            "1 s CPI @ 2 MS/s, N=2e6 per channel")
KERNELS_PER_STEP = 4 + 2 + 2 + 3 + 1  # wh(corr,solve,wspec,apply) caf(range,doppler) metrics(2) cfar(flag,scan,emit) tail(centroid+interp)


def ncu_traffic(kernel="caf_range_", capture="cfg2"):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the range kernel (caf_range_kernel or
    caf_range_grouped_kernel, whichever the plan uses) in the config-2 capture of the committed ncu --set full
    extract (profiles/r01z_kernels.json), or None."""
    p = os.path.join(ROOT, "profiles", "r01z_kernels.json")
    try:
        ks = json.load(open(p))["kernels"]
        for k in ks:
            if kernel in k["name"] and capture in str(k.get("capture", capture)):
                return int(k["dram_bytes_read"] + k["dram_bytes_write"])
    except Exception:
        pass
    return None


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [v.strip() for v in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def _cpu_chain():
    """(run(x, y) -> (stage_ms, n_detections), kind): the compiled reference (oracle/_ref) when it was built,
    else the numpy port of the same algorithm (oracle/blah2_oracle.py)."""
    from oracle import refpath as R
    if R.available():
        ch = R.Chain(GEOM["delayMin"], GEOM["delayMax"], GEOM["dopplerMin"], GEOM["dopplerMax"], FS, N, True, clutter=CLUTTER,
                     **DET)

        def run(x, y):
            r = ch.run(x, y, want_map=False)
            return [float(v) for v in r["stage_ms"]], int(len(r["detections"][0]))
        return run, "reference"
    from oracle import blah2_oracle as O
    g = O.ambiguity_geometry(GEOM["delayMin"], GEOM["delayMax"], GEOM["dopplerMin"], GEOM["dopplerMax"], FS, N, True)

    def run(x, y):
        t0 = time.perf_counter()
        r = O.chain(x, y, g, clutter=CLUTTER, det=DET)
        return [0.0, (time.perf_counter() - t0) * 1e3, 0.0], int(len(r["detections"][0]))
    return run, "port"


def _ref_worker(conn, seed):
    """One host process = one stream of CPIs through the reference's own classes (separate address spaces:
    the reference's per-sample deque traffic makes threads contend on the allocator)."""
    from blah2_b200.scene import make_scene

    sc = make_scene(N, FS, seed=seed)
    run, kind = _cpu_chain()
    conn.send("ready:" + kind)
    while True:
        cmd = conn.recv()
        if cmd == "stop":
            break
        conn.send(run(sc.x, sc.y)[0])


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path on the host cores."""
    if rank != 0:
        return
    import multiprocessing as mp

    cores = os.cpu_count() or 1
    try:
        avail_gb = os.sysconf("SC_AVPHYS_PAGES") * os.sysconf("SC_PAGE_SIZE") / 2**30
    except Exception:
        avail_gb = 16.0
    procs_n = int(max(1, min(cores, 32, avail_gb // 3)))  # ~2 GB of FFT plans + buffers per process
    ctx = mp.get_context("spawn")
    workers = []
    for i in range(procs_n):
        a, b = ctx.Pipe()
        p = ctx.Process(target=_ref_worker, args=(b, 20260923), daemon=True)
        p.start()
        workers.append((p, a))
    kind = "reference"
    for _, c in workers:
        msg = c.recv()
        assert msg.startswith("ready:")
        kind = msg.split(":", 1)[1]

    def step():
        for _, c in workers:
            c.send("run")
        return [c.recv() for _, c in workers]

    # bounded: a CPI takes seconds on the CPU; warm-up is capped at one step and the timed steps stop
    # after ~150 s (steps actually timed are reported as "steps")
    for _ in range(min(args.warmup, 1)):
        step()
    t0 = time.perf_counter()
    stages = []
    done = 0
    for _ in range(args.steps):
        stages += step()
        done += 1
        if time.perf_counter() - t0 > 150.0:
            break
    dt = time.perf_counter() - t0
    for p, c in workers:
        c.send("stop")
    requested = args.steps
    args.steps = done
    cpis = procs_n * args.steps
    value = cpis * N / dt / 1e6
    st = np.mean(np.array(stages), axis=0)
    line = {
        "impl": "reference", "metric": "iq_msamples_per_s", "value": round(value, 4), "unit": "Msamples/s",
        "maps_per_s": round(cpis / dt, 4), "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(dt / args.steps * 1e3, 3), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "step": f"{procs_n} CPIs, one per host process", "steps_requested": requested},
        "cpu_baseline": {"value": round(value, 4), "unit": "Msamples/s", "cores": procs_n, "kind": kind,
                         "sample": f"{cpis} full CPIs ({procs_n} concurrent host processes); " +
                                   ("unmodified reference src/process sources linked to this repo's FFTW/Armadillo shims "
                                    "(stock FFTW not in the image)" if kind == "reference" else
                                    "numpy port of the reference algorithm (oracle/_ref was not built)"),
                         "stage_ms": {"clutter_filter": round(float(st[0]), 1),
                                      "ambiguity_processing": round(float(st[1]), 1),
                                      "detector": round(float(st[2]), 2)}, "host_cores": cores},
        "e2e": {"value": round(value, 4), "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--graph", action="store_true", help="replay the device chain as CUDA graphs (prepare_device)")
    ap.add_argument("--streams", type=int, default=6, help="CPIs in flight per GPU (independent pipelines on their own streams)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    from blah2_b200.process import Ambiguity, Pipeline, WienerHopf
    from blah2_b200.scene import make_scene
    from blah2_b200.shard import gather_maps

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    # stdout carries exactly ONE line, the JSON: keep the real stdout aside and point file descriptor 1 at stderr, so
    # that anything a library prints there (NCCL's "NCCL version ..." banner, which it writes at WARN level too)
    # lands on stderr instead
    sys.stdout.flush()
    json_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    NPIPE = max(1, args.streams)
    pipes = [Pipeline(**GEOM, clutter=CLUTTER, detection=DET, device=local_rank) for _ in range(NPIPE)]
    pipe = pipes[0]
    g = pipe.geometry
    cells = g.n_doppler_bins * g.n_delay_bins

    # ---- inputs: NB distinct CPIs resident in HBM (> 2x the 126 MB L2) + 2 pinned host CPIs ----
    sc = make_scene(N, FS, seed=20260923 + rank)
    NB = 10
    x0 = torch.from_numpy(sc.x.astype(np.complex64)).cuda()
    y0 = torch.from_numpy(sc.y.astype(np.complex64)).cuda()
    xs = [torch.roll(x0, 977 * b) for b in range(NB)]
    ys = [torch.roll(y0, 977 * b) for b in range(NB)]
    hx = [torch.from_numpy(np.roll(sc.x, 977 * b)).pin_memory() for b in range(2)]
    hy = [torch.from_numpy(np.roll(sc.y, 977 * b)).pin_memory() for b in range(2)]
    hmap = torch.empty((g.n_doppler_bins, g.n_delay_bins), dtype=torch.complex128).pin_memory()
    dmaps = [torch.empty((g.n_doppler_bins, g.n_delay_bins), dtype=torch.complex64, device="cuda") for _ in range(NPIPE)]
    dmap = dmaps[0]
    gathered = None
    streams = [torch.cuda.Stream() for _ in range(NPIPE)]
    stream = streams[0]
    st = stream.cuda_stream

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def submit(i):
        p = i % NPIPE
        with torch.cuda.stream(streams[p]):
            pipes[p].submit_device(xs[i % NB], ys[i % NB], dmaps[p], streams[p].cuda_stream)

    # ---- device-resident throughput: a stream of independent CPIs, NPIPE in flight ----
    # --graph: plan creation (untimed, before the warm-up steps) = the CUDA graph of the chain for every (input set,
    # pipeline) pair the loop below submits.  Off by default: with 50 CPIs enqueued ahead eager launches measured 3 %
    # faster than graph replay (profiles/r01_summary.md s6)
    if args.graph:
        for i in range(NB * NPIPE):   # submit(i) uses (i % NB, i % NPIPE): the pattern repeats after lcm(NB, NPIPE) steps
            p = i % NPIPE
            pipes[p].prepare_device(xs[i % NB], ys[i % NB], dmaps[p], streams[p].cuda_stream)
    for i in range(args.warmup):
        submit(i)
    for p in range(NPIPE):
        last = pipes[p].fetch(streams[p].cuda_stream)
    if world > 1:  # warm the NCCL communicator and the gather path outside the timed region
        with torch.cuda.stream(stream):
            for _ in range(2):
                gather_maps(dmap.unsqueeze(0), world, rank, world)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for s_ in streams[1:]:
        s_.wait_event(e0)
    for i in range(args.steps):
        submit(i)
    for p in reversed(range(NPIPE)):  # synchronises each stream; detections + metrics of the final CPIs on the host
        last = pipes[p].fetch(streams[p].cuda_stream)
    with torch.cuda.stream(stream):
        if world > 1:  # the final map gather (NCCL over NVLink), ordered after the kernels on `stream`
            gathered = gather_maps(dmap.unsqueeze(0), world, rank, world)
        e1.record(stream)
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())

    # ---- end to end through the host API (pinned complex128 in, complex128 map out) ----
    # Two pipelines alternate: submit_host(i) enqueues H2D + kernels + D2H, fetch(i-1) collects the
    # previous CPI's map/detections, so PCIe transfers of one CPI overlap the kernels of the other.
    hmaps = [torch.empty((g.n_doppler_bins, g.n_delay_bins), dtype=torch.complex128).pin_memory() for _ in range(NPIPE)]
    for i in range(2 * NPIPE):
        pipes[i % NPIPE].process(hx[i % 2], hy[i % 2], map_out=hmaps[i % NPIPE])
    barrier()
    e2e_steps = max(6, min(args.steps, 30))
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        p = i % NPIPE
        if i >= NPIPE:
            r_e2e = pipes[p].fetch()          # result of CPI i - NPIPE is now in host memory
        pipes[p].submit_host(hx[i % 2], hy[i % 2], map_out=hmaps[p])
    for i in range(e2e_steps, e2e_steps + NPIPE):
        r_e2e = pipes[i % NPIPE].fetch()
    torch.cuda.synchronize()
    e2e_s = torch.tensor([time.perf_counter() - t0], device="cuda")
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_s = float(e2e_s.item())

    # ---- the same, fed with the reference's replay layout (int16 I1 Q1 I2 Q2, 8 B per instant) ----
    def to_i16(b):
        x_, y_ = np.roll(sc.x, 977 * b), np.roll(sc.y, 977 * b)
        iq = np.empty((N, 4), dtype="<i2")
        iq[:, 0], iq[:, 1], iq[:, 2], iq[:, 3] = x_.real, x_.imag, y_.real, y_.imag
        return torch.from_numpy(iq).pin_memory()

    hq = [to_i16(b) for b in range(2)]
    for i in range(NPIPE):
        pipes[i].submit_host_rspduo(hq[i % 2], map_out=hmaps[i])
        pipes[i].fetch()
    barrier()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        p = i % NPIPE
        if i >= NPIPE:
            r_i16 = pipes[p].fetch()
        pipes[p].submit_host_rspduo(hq[i % 2], map_out=hmaps[p])
    for i in range(e2e_steps, e2e_steps + NPIPE):
        r_i16 = pipes[i % NPIPE].fetch()
    torch.cuda.synchronize()
    i16_s = torch.tensor([time.perf_counter() - t0], device="cuda")
    if world > 1:
        dist.all_reduce(i16_s, op=dist.ReduceOp.MAX)
    i16_s = float(i16_s.item())

    clocks = sampler.stop() if rank == 0 else None  # sampled over the device-resident AND the end-to-end timed regions

    # ---- per-kernel durations for the roofline (CUDA events around each kernel) ----
    amb = Ambiguity(GEOM["delayMin"], GEOM["delayMax"], GEOM["dopplerMin"], GEOM["dopplerMax"], FS, N, True,
                    device=local_rank)
    wh = WienerHopf(CLUTTER[0], CLUTTER[1], N, device=local_rank)
    yf = torch.empty_like(ys[0])
    prof = {"range": [], "doppler": [], "wh_corr": [], "wh_solve": [], "wh_apply": []}
    with torch.cuda.stream(stream):
        for i in range(3 + min(args.steps, 30)):
            c, s_, a_ = wh.profile_device(xs[i % NB], ys[i % NB], yf, st)
            r_, d_ = amb.profile_device(xs[i % NB], yf, dmap, st)
            if i >= 3:
                prof["range"].append(r_); prof["doppler"].append(d_)
                prof["wh_corr"].append(c); prof["wh_solve"].append(s_); prof["wh_apply"].append(a_)
    kms = {k: float(np.mean(v)) for k, v in prof.items()}

    # ---- SpectrumAnalyser(n, 2000) (blah2.cpp:198,263-265) on the resident reference channel: not part of the
    # step above (BASELINE's configs do not include it); its folding pass is the path's purely HBM-bound kernel ----
    spectrum = None
    if rank == 0:
        from blah2_b200.process import SpectrumAnalyser
        spectrum = {"api": "SpectrumAnalyser(n, 2000.0).process_device(float2 x)", "kernel": "spec_fold_kernel",
                    "algorithmic_bytes": "8 nfft (x read once)"}
        big = [torch.randn(20_000_000, dtype=torch.complex64, device="cuda") for _ in range(2)]
        for label, n_, bufs in (("n2e6", N, xs), ("n2e7", 20_000_000, big)):
            sa = SpectrumAnalyser(n_, 2000.0, device=local_rank)
            f_, r_ = [], []
            with torch.cuda.stream(stream):
                for i in range(3 + 20):
                    a_, b_ = sa.profile_device(bufs[i % len(bufs)], st)
                    if i >= 3:
                        f_.append(a_); r_.append(b_)
            fm, rm = float(np.mean(f_)), float(np.mean(r_))
            spectrum[label] = {"n_spectrum": sa.nSpectrum, "decimation": sa.decimation, "fold_ms": round(fm, 5),
                               "reduce_dft_ms": round(rm, 5), "achieved_gbs": round(8 * sa.nfft / (fm * 1e-3) / 1e9, 1)}
            sa.close()
        del big

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peaks()
    for label in ("n2e6", "n2e7"):
        spectrum[label]["frac"] = round(spectrum[label]["achieved_gbs"] / peak, 4)
    bytes_range = 16 * g.n_used + 8 * cells                 # x, y read once; range matrix written once
    bytes_caf = 16 * g.n_used + 8 * cells                   # SURVEY s8(d) B_caf (map written; R stays in L2)
    bytes_step = 2 * 16 * N + 8 * N + 8 * cells             # two compulsory passes over x,y + y' + map
    ach_range = bytes_range / (kms["range"] * 1e-3) / 1e9
    value = world * args.steps * N / (ms_total * 1e-3) / 1e6
    e2e_value = world * e2e_steps * N / e2e_s / 1e6
    line = {
        "metric": "iq_msamples_per_s", "value": round(value, 2), "unit": "Msamples/s",
        "maps_per_s": round(world * args.steps / (ms_total * 1e-3), 2),
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_total / args.steps, 5),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32 (CAF) / f64 (WienerHopf, detection)",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "cpis_per_step_per_gpu": 1, "parallelism": f"independent CPIs x{world}",
                   "l2": f"{NB} distinct CPI input sets rotated ({NB * 32} MB > L2)", "cpis_in_flight": NPIPE, "submission": "cuda graph replay" if args.graph else "eager launches",
                   "range_fft": f"M={g.range_fft_len} x{g.range_segments} segments, hop {g.range_hop}, "
                                f"{g.range_groups} warp group(s) x {g.range_parts} part(s) per batch",
                   "doppler_fft": f"Bluestein M2={g.doppler_fft_len}"},
        "e2e": {"value": round(e2e_value, 2), "unit": "Msamples/s", "h2d_bytes_per_step": 2 * 16 * N,
                "d2h_bytes_per_step": 16 * cells + 3 * 8 * int(r_e2e["detections"].get_nDetections()) + 32,
                "ms_per_step": round(e2e_s / e2e_steps * 1e3, 4), "steps": e2e_steps,
                "api": "Pipeline.submit_host(pinned complex128 x, y) / fetch() -> complex128 map + detections, "
                       f"{NPIPE} CPIs in flight"},
        "e2e_rspduo_int16": {"value": round(world * e2e_steps * N / i16_s / 1e6, 2), "unit": "Msamples/s",
                             "h2d_bytes_per_step": 8 * N, "ms_per_step": round(i16_s / e2e_steps * 1e3, 4),
                             "api": "Pipeline.submit_host_rspduo(pinned int16 I1 Q1 I2 Q2) / fetch()",
                             "n_detections": int(r_i16["detections"].get_nDetections())},
        "gpu_launches": KERNELS_PER_STEP * args.steps,
        "clocks": clocks,
        "roofline": {"kernel": "caf_range_grouped_kernel" if g.range_groups > 1 else "caf_range_kernel", "bound": "hbm", "achieved": round(ach_range, 1), "peak": peak,
                     "unit": "GB/s", "frac": round(ach_range / peak, 4), "traffic": ncu_traffic(), "peak_source": peak_src,
                     "algorithmic_bytes": bytes_range, "kernel_ms": round(kms["range"], 5),
                     "caf_total": {"ms": round(kms["range"] + kms["doppler"], 5),
                                   "frac": round(bytes_caf / ((kms["range"] + kms["doppler"]) * 1e-3) / 1e9 / peak, 4)},
                     "step": {"algorithmic_bytes": bytes_step,
                              "frac": round(bytes_step / (ms_total / args.steps * 1e-3) / 1e9 / peak, 4)}},
        "kernel_ms": {k: round(v, 5) for k, v in kms.items()},
        "spectrum": spectrum,
        "result": {"n_detections": int(last["detections"].get_nDetections()), "noisePower": round(last["noisePower"], 4),
                   "maxPower": round(last["maxPower"], 4), "filter_ok": not last["skipped"]},
    }
    if world == 1 and not args.no_cpu_baseline:
        run, kind = _cpu_chain()
        t0 = time.perf_counter()
        stage_ms, n_det = run(sc.x, sc.y)
        dt = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": round(N / dt / 1e6, 4), "unit": "Msamples/s", "cores": 1, "kind": kind,
                                "sample": "1 full CPI of the same workload on 1 host thread (" +
                                          ("reference src/process sources unmodified + this repo's FFTW/Armadillo shims"
                                           if kind == "reference" else "numpy port, oracle/_ref was not built") + ")",
                                "stage_ms": {"clutter_filter": round(stage_ms[0], 1),
                                             "ambiguity_processing": round(stage_ms[1], 1),
                                             "detector": round(stage_ms[2], 2)},
                                "host_cores": os.cpu_count(), "n_detections": n_det}
        try:  # the reference's SpectrumAnalyser on the same reference channel (blah2.cpp:263-265), for the `spectrum` block
            from oracle import refpath as R
            if R.available():
                t0 = time.perf_counter()
                R.spectrum_process(sc.x, N, 2000.0)
                line["cpu_baseline"]["spectrum_ms"] = round((time.perf_counter() - t0) * 1e3, 1)
        except Exception:
            pass
    print(json.dumps(line), file=json_out, flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
