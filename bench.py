#!/usr/bin/env python
"""Benchmark of the visibility hot path on BASELINE.json's headline configuration.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched under torchrun)
    python bench.py --impl reference ...                     (CPU arm: the reference's numpy path)

One "step" = one snapshot of configuration 2 (HERA-350, 61,075 baselines x 1024 channels against
the 300k-source GLEAM-shaped catalogue; 178,987 sources are above the horizon at LST 0h):
horizon cull + flux x beam amplitude table + the phase sum.  Metric: Gterms/s, one term = one
(source above horizon, baseline, channel) triple.

  value     device-timed (CUDA events), catalogue and array already resident in HBM
  e2e       the same snapshot through InterferometerArray.observe with HOST inputs: catalogue
            copied host->device from pinned memory and the visibilities read back device->host
            into pinned memory inside the timed region, every step
  roofline  the phase-sum kernel against the FP32-FMA issue roofline: 6 FMA-pipe lane-issues =
            12 flop-equivalents per term; peak = FFMA rate measured on this GPU by
            pb200_microbench in the same process (MEASURED_PEAKS.json has no FP32 entry)
  cpu_baseline  the oracle's float64 numpy restatement of interferometry.py:6332-6340 on the host
            cores (multiprocessing over baseline chunks like run_prisim's pp.key='bl'), bounded
            sample of the same workload.  kind = "port": the reference itself is Python 2 with
            un-installable dependencies.

Multi-GPU (weak scaling): every rank simulates its own snapshot (LST offset by rank) with all
baselines -- the path shards over snapshots with no data-path collective -- and the finished
visibilities are gathered to rank 0 over NCCL inside the timed region.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as NP

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "Gterms/s (src x bl x chan) HERA-350 x 300k-source catalogue x 1024 ch"
WORKLOAD = "config2: HERA-350 (61,075 bl) x 1024 ch x 97.65625 kHz x 300k-src GLEAM-shaped catalogue, Airy 14 m, 1 snapshot/GPU"


# --------------------------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0, period=0.2):
        super().__init__(daemon=True)
        self.gpu, self.period, self.samples, self._stop_evt = gpu_index, period, [], threading.Event()

    def run(self):
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([v.strip() for v in out.split(",")])
            except Exception:
                pass
            self._stop_evt.wait(self.period)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=5)
        sm, smax, reasons = [], 0.0, set()
        for s in self.samples:
            try:
                sm.append(float(s[0])); smax = max(smax, float(s[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": smax or None, "reasons": sorted(reasons),
                "samples": len(sm)}


# --------------------------------------------------------------------------------------------
# CPU arm: the reference's numpy expression on the host cores
# --------------------------------------------------------------------------------------------
def _cpu_chunk(args):
    from oracle import prisim_oracle as O
    bl, altaz, pbfluxes, channels, pc_altaz = args
    return O.skyvis_snapshot(bl, altaz, pbfluxes, channels, pc_altaz, max_slab_bytes=6.4e7)


def cpu_sample(cfg, nsrc, nbl, nproc, seed=0):
    """Time the oracle on a bounded slice of the workload: first `nsrc` above-horizon sources x
    `nbl` evenly spaced baselines x all channels, baseline chunks over `nproc` processes."""
    import multiprocessing as mp
    from oracle import prisim_oracle as O
    sky = cfg["skymodel"]
    hadec = NP.stack((0.0 - sky.location[:, 0], sky.location[:, 1]), axis=1)
    altaz = O.hadec2altaz(hadec, cfg["latitude"])
    m2 = O.roi_select(altaz)[:nsrc]
    sp = sky.spec_parms
    # ProcessPoolExecutor raises BrokenProcessPool if a worker cannot start (a bare Pool would
    # respawn forever); spawn, not fork: the GPU arm has a live CUDA context in this process
    from concurrent.futures import ProcessPoolExecutor
    pool = ProcessPoolExecutor(nproc, mp_context=mp.get_context("spawn")) if nproc > 1 else None
    if pool:
        list(pool.map(abs, range(nproc), timeout=300))       # workers up before the clock starts
    t0 = time.time()
    pb = O.primary_beam_generator(altaz[m2], cfg["channels"] / 1e9, cfg["telescope"], skyunits="altaz",
                                  pointing_center=NP.asarray([90.0, 270.0]))
    pbf = pb * O.power_law_spectrum(sp["flux-scale"][m2], sp["power-law-index"][m2], sp["freq-ref"][m2], cfg["channels"])
    bsel = NP.linspace(0, cfg["baselines"].shape[0] - 1, nbl).astype(int)
    bl = cfg["baselines"][bsel]
    chunks = NP.array_split(NP.arange(nbl), nproc)
    jobs = [(bl[c], altaz[m2], pbf, cfg["channels"], NP.asarray([90.0, 270.0])) for c in chunks if c.size]
    if pool:
        list(pool.map(_cpu_chunk, jobs, timeout=900))
    else:
        for j in jobs:
            _cpu_chunk(j)
    dt = time.time() - t0
    if pool:
        pool.shutdown()
    terms = float(m2.size) * nbl * cfg["channels"].size
    return terms / dt, dt, terms


def host_cores():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from prisim_b200 import synthetic as S
    cfg = S.config2()
    cores = min(host_cores(), 64)
    nsrc, nbl = 4000, max(cores * 16, 64)
    for _ in range(args.warmup):
        cpu_sample(cfg, max(nsrc // 4, 50), nbl, cores)
    rates, times = [], []
    for _ in range(args.steps):
        r, dt, terms = cpu_sample(cfg, nsrc, nbl, cores)
        rates.append(r); times.append(dt)
    value = statistics.median(rates) / 1e9
    sample = "{0} above-horizon sources x {1} baselines x 1024 channels per step ({2:.2e} terms), float64 numpy, {3} processes".format(nsrc, nbl, terms, cores)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "Gterms/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * statistics.median(times), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "note": "CPU arm: bounded sample of the same workload per step"},
            "cpu_baseline": {"value": value, "unit": "Gterms/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "Gterms/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--nsrc", type=int, default=300000, help="catalogue size (default = the headline 300k)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
        return

    import torch
    import torch.distributed as dist
    from prisim_b200 import _lib, engine
    from prisim_b200 import primary_beams as PB
    from prisim_b200 import synthetic as S
    from prisim_b200.interferometry import InterferometerArray, SimpleTime
    from prisim_b200.sharding import gather_baseline_shards  # noqa: F401  (baseline sharding lives there; bench shards snapshots)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = "cuda:{0}".format(local_rank)
    # NCCL prints its version banner on stdout at communicator creation; keep stdout clean for the
    # single JSON line by pointing fd 1 at stderr until the result is printed
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(dev))
    warmup = max(args.warmup, 3)

    cfg = S.config2(nsrc=args.nsrc)
    sky = cfg["skymodel"]
    sp = sky.spec_parms
    nbl, nchan = cfg["baselines"].shape[0], cfg["channels"].size
    lst_deg = 0.0 + 15.0 * rank / 8.0                      # every rank observes its own snapshot
    ctx = _lib.get_context(local_rank)

    # ---- resident inputs ----
    d_hadec = engine._f64(NP.stack((lst_deg - sky.location[:, 0], sky.location[:, 1]), axis=1), local_rank)
    spec = {"flux_scale": engine._f64(sp["flux-scale"], local_rank), "index": engine._f64(sp["power-law-index"], local_rank),
            "freq_ref": engine._f64(sp["freq-ref"], local_rank)}
    d_bl = engine._f64(cfg["baselines"], local_rank)
    pc_dircos = NP.asarray([0.0, 0.0, 1.0])
    beam = PB.beam_desc_from_telescope(cfg["telescope"], pointing_center=NP.asarray([90.0, 270.0]), skyunits="altaz", device=local_rank)
    # multi-GPU: the phase-sum kernel's epilogue stores each rank's snapshot straight into rank 0's buffer over
    # NVLink peer memory (sharding.PeerGatherBuffer); NCCL point-to-point is the fallback if mapping fails
    from prisim_b200.sharding import PeerGatherBuffer
    gbuf = PeerGatherBuffer((nbl, nchan), local_rank, dst=0) if world > 1 else None
    vis = gbuf.local if gbuf is not None else torch.empty((nbl, nchan), dtype=torch.complex128, device=dev)
    k1_events = []

    def step(timed):
        dircos, index = engine.sky_cull(d_hadec, "hadec", latitude_deg=cfg["latitude"], device=local_rank)
        nsrc = int(index.shape[0])
        amp = engine.amp_table(dircos, index, nsrc, spec, beam, cfg["channels"], device=local_rank)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        engine.skyvis(dircos, amp, nsrc, d_bl, pc_dircos, cfg["channels"], out=vis, device=local_rank)
        e1.record()
        if timed:
            k1_events.append((e0, e1))
        if gbuf is not None:                              # the single gather of the path (to the writing rank)
            gbuf.wait()
        return nsrc

    def fence():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(warmup):
        nsrc = step(False)
    fence()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    launches0 = ctx.launches
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(args.steps):
        nsrc = step(True)
    t1.record()
    fence()
    elapsed_ms = t0.elapsed_time(t1)
    launches = ctx.launches - launches0
    clocks = sampler.stop() if sampler else None
    k1_ms = statistics.mean(a.elapsed_time(b) for a, b in k1_events)

    terms_local = float(nsrc) * nbl * nchan
    stats = torch.tensor([elapsed_ms, terms_local, k1_ms], dtype=torch.float64, device=dev)
    if world > 1:
        mx = stats.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = stats.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        elapsed_ms, k1_ms = mx[0].item(), mx[2].item()
        terms_total = sm[1].item()
    else:
        terms_total = terms_local
    ms_per_step = elapsed_ms / args.steps
    value = terms_total / (ms_per_step * 1e-3) / 1e9

    # ---- end-to-end through the public API with host buffers ----
    e2e = None
    if not args.no_e2e:
        def pinned(a):
            return torch.from_numpy(NP.ascontiguousarray(a, dtype=NP.float64)).pin_memory().numpy()
        sky.location = pinned(sky.location)
        for key in ("flux-scale", "power-law-index", "freq-ref", "flux-offset"):
            sp[key] = pinned(sp[key])
        ia = InterferometerArray(cfg["labels"], cfg["baselines"], cfg["channels"], telescope=cfg["telescope"], latitude=cfg["latitude"],
                                 skycoords="radec", pointing_coords="hadec", device=local_rank)
        ia.cache_sky = False                              # force the host->device copy of the catalogue every step
        host_vis = torch.empty((nbl, nchan), dtype=torch.complex128, pin_memory=True)
        h2d = sky.location.nbytes + sum(sp[k].nbytes for k in ("flux-scale", "power-law-index", "freq-ref"))
        d2h = host_vis.numel() * 16

        def e2e_step():
            ia._skyvis, ia._bp, ia._Tsys, ia.timestamp = [], [], [], []     # keep one snapshot resident
            ia.obs_catalog_indices = []
            ia.observe(SimpleTime(2451545.0, lst_deg), {"Tnet": 300.0}, NP.ones(nchan), cfg["pointing_hadec"], sky, cfg["t_acc"])
            host_vis.copy_(ia.skyvis_freq_device(0), non_blocking=True)
            torch.cuda.synchronize()

        for _ in range(2):
            e2e_step()
        fence()
        w0 = time.perf_counter()
        for _ in range(args.steps):
            e2e_step()
        fence()
        e2e_ms = (time.perf_counter() - w0) * 1e3 / args.steps
        st = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(st, op=dist.ReduceOp.MAX)
        e2e = {"value": terms_total / (st[0].item() * 1e-3) / 1e9, "unit": "Gterms/s", "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "ms_per_step": st[0].item(),
               "api": "InterferometerArray.observe (precision='auto': fp32 kernel + fp64 recompute of cancelling baselines + "
                      "sampled fp64 audit) + device->host copy of skyvis_freq (pinned)",
               "precision_report": ia.precision_report[-1] if ia.precision_report else None}

    if rank == 0:
        mb = engine.microbench(local_rank)
        peak_tflops = mb["fp32_tflops"]
        k1_terms_per_s = terms_local / (k1_ms * 1e-3)
        achieved = 12.0 * k1_terms_per_s / 1e12
        peaks_file = os.path.join(ROOT, "MEASURED_PEAKS.json")
        hbm = json.load(open(peaks_file)).get("hbm_gbs") if os.path.exists(peaks_file) else None
        roofline = {"bound": "fp32_fma_issue", "kernel": "k_skyvis", "achieved": achieved, "peak": peak_tflops, "unit": "TFLOP/s",
                    "frac": achieved / peak_tflops, "peak_source": "FFMA rate measured on this GPU (pb200_microbench, same process)",
                    "peak_nominal": 148 * 128 * 2 * 1.965e9 / 1e12, "frac_of_nominal": achieved / (148 * 128 * 2 * 1.965e9 / 1e12),
                    "flop_equiv_per_term": 12, "kernel_ms": k1_ms, "kernel_share_of_step": k1_ms / ms_per_step,
                    "kernel_gterms_per_s": k1_terms_per_s / 1e9,
                    "traffic": TRAFFIC_BYTES_PER_LAUNCH, "algorithmic_bytes": float(nsrc) * nchan * 4 + nbl * nchan * 16.0,
                    "hbm_gbs_measured": hbm, "mufu_per_s": mb["mufu_per_s"], "dfma_per_s": mb["dfma_per_s"],
                    "microbench_sm_clock_hz": mb["sm_clock_hz"]}
        cpu = None
        if not args.no_cpu_baseline:
            cores = min(host_cores(), 64)
            rate, dt, tterms = cpu_sample(cfg, 8000, max(cores * 32, 64), cores)
            cpu = {"value": rate / 1e9, "unit": "Gterms/s", "cores": cores, "kind": "port", "seconds": dt,
                   "mterms_per_s_per_core": rate / 1e6 / cores,
                   "extrapolated_full_config_seconds": terms_local / rate,
                   "sample": "8000 above-horizon sources x {0} baselines x 1024 channels ({1:.2e} terms) of the same workload, "
                             "float64 numpy restatement of interferometry.py:6332-6340, {2} processes".format(max(cores * 32, 64), tterms, cores)}
        line = {"metric": METRIC, "value": value, "unit": "Gterms/s", "n_gpus": world, "steps": args.steps, "warmup": warmup,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic",
                "config": {"workload": WORKLOAD, "nsrc_catalogue": args.nsrc, "nsrc_above_horizon": nsrc, "nbl": nbl, "nchan": nchan,
                           "terms_per_step_per_gpu": terms_local, "sharding": "one snapshot per GPU, results land in rank 0's buffer ({0})".format(
                               "single GPU" if gbuf is None else ("kernel epilogue stores over NVLink peer memory" if gbuf.mode == "peer"
                                                                  else "NCCL point-to-point gather")),
                           "l2": "inputs larger than L2: amplitude table {0:.2f} GB + 1.0 GB output per step".format(nsrc * nchan * 4 / 1e9),
                           "phase_arith": "fp64 anchors, fp32 rotation recurrence, fp32 accumulate flushed to fp64"},
                "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu}
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(line))
        sys.stdout.flush()
    if world > 1:
        gbuf.close()
        dist.barrier()
        dist.destroy_process_group()


# dram__bytes_read.sum + dram__bytes_write.sum per k_skyvis launch at the headline size, from the
# committed ncu capture (profiles/); None until that capture exists for the current kernel.
TRAFFIC_BYTES_PER_LAUNCH = 29.70e9   # profiles/skyvis_headline_r01_ncu.txt: 22.86 GB read + 6.84 GB written (0.19 % of HBM bandwidth over 2.44 s;
                                      # the amplitude table is re-streamed once per wave of CTAs, 26 waves x 0.73 GB; the staggered fp64
                                      # flushes keep less of the 1 GB running-sum scratch resident in L2 than the synchronous ones did)

if __name__ == "__main__":
    main()
