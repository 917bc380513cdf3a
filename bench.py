#!/usr/bin/env python
"""Benchmark of the visibility hot path on BASELINE.json's configurations.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched under torchrun)
    python bench.py --config 3|4|5 ...                       (configs 3, 4 and 5; default 2 = the headline)
    python bench.py --impl reference ...                     (CPU arm: the reference's numpy path)

Metric everywhere: Gterms/s, one term = one (source above horizon, baseline, channel) triple of one
snapshot.  One "step":

  config 2 (default, the headline): ONE snapshot of HERA-350 (61,075 baselines x 1024 channels) against the
      300k-source GLEAM-shaped catalogue (178,987 above the horizon at LST 0h): horizon cull + flux x beam
      amplitude table + the phase sum.  N GPUs: STRONG scaling -- the one snapshot is sharded over contiguous
      baseline blocks (the reference's pp.key='bl' mode, scripts/run_prisim.py:1775-1791, :2165-2209), every
      rank's kernel epilogue stores its rows into rank 0's buffer over NVLink peer memory (the rank-0
      concatenation of :2233-2242), value = the snapshot's terms / max-over-ranks time.  `--scaling weak`
      keeps round 1's mode (one snapshot per rank).
  config 3: one snapshot of the nside-256 diffuse sky (393k pixels above the horizon, extended sources ->
      taper) x HERA-331 x 256 channels through the fp64 kernel with the gridded-HEALPix beam; snapshots are
      dealt round-robin to the ranks (one snapshot per rank and step).
  config 4: one snapshot of 128 MWA-like tiles (8,128 baselines) x 768 channels with the phased 4x4 tile beam (quantised delays,
      ground plane) evaluated per source and channel inside the amplitude-table kernel; snapshots round-robin over ranks.
  config 5: observe + Tsys noise + add + three windowed delay transforms of one HERA-350 snapshot, streamed
      through InterferometerArray.drain; baseline-sharded over N ranks like config 2.

  value     device-timed (CUDA events), catalogue and array already resident in HBM
  e2e       the same step through the public API (InterferometerArray / sharding.ShardedObserver) with HOST
            inputs: catalogue copied host->device from pinned memory and the products read back device->host
            into pinned memory (ONE copy, from the writing rank) inside the timed region, every step
  roofline  the dominant kernel against its issue roofline.  Phase sum: 6 FMA-pipe lane-issues = 12
            flop-equivalents per term; peak = 148 SMs x 128 lanes x 2 x the SM clock SAMPLED DURING THE TIMED
            LOOP (nvidia-smi); the FFMA rate measured by pb200_microbench (lanes/clk/SM, clock-independent) is
            reported beside it.  MEASURED_PEAKS.json has no FP32 entry.
  gather_check  after the timed loop every rank recomputes its block into private memory and the int64
            bit-pattern checksum of that block is compared with the checksum of the same rows of rank 0's
            buffer; 64 rows sampled over all shards are recomputed by the fp64 kernel (max |dV| / rms_b).
  cpu_baseline  the oracle's float64 numpy restatement of interferometry.py:6332-6340 on the host cores
            (multiprocessing over baseline chunks like run_prisim's pp.key='bl'), bounded sample of the same
            workload.  kind = "port": the reference itself is Python 2 with un-installable dependencies.
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
WORKLOADS = {
    2: "config2: HERA-350 (61,075 bl) x 1024 ch x 97.65625 kHz x 300k-src GLEAM-shaped catalogue, Airy 14 m, 1 snapshot",
    3: "config3: nside-256 diffuse sky (786,432 pixels, horizon-culled, extended sources) x HERA-331 (54,615 bl) x 256 ch x 390.625 kHz, "
       "gridded HEALPix beam (nside 128), fp64 kernel, 1 snapshot per GPU and step",
    4: "config4: MWA Phase-II-like 128 tiles (8,128 bl) x 768 ch x 40 kHz x 50k-src catalogue, phased 4x4 dipole tile beam (quantised delays) "
       "+ ground plane evaluated per source / channel / snapshot, 1 snapshot per GPU and step",
    5: "config5: HERA-350 x 1024 ch x 300k-src catalogue: visibilities + Tsys noise + three windowed delay transforms per snapshot",
}
NOMINAL_LANES = 148 * 128          # FP32 FMA lanes per clock on the whole chip


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
        sm, smax, reasons, power = [], 0.0, set(), []
        for s in self.samples:
            try:
                sm.append(float(s[0])); smax = max(smax, float(s[1])); power.append(float(s[2]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": smax or None, "reasons": sorted(reasons),
                "samples": len(sm), "power_w_median": statistics.median(power) if power else None}


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
            "warmup": args.warmup, "ms_per_step": 1e3 * statistics.median(times), "higher_is_better": True,
            "scaling": "strong" if args.scaling in ("auto", "strong") else "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOADS[2], "note": "CPU arm: bounded sample of the same workload per step; the host cores do not "
                                                         "depend on --gpus"},
            "cpu_baseline": {"value": value, "unit": "Gterms/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "Gterms/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------
# GPU arm: shared plumbing
# --------------------------------------------------------------------------------------------
class Env(object):
    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.args = torch, dist, args
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local_rank)
        self.dev = "cuda:{0}".format(self.local_rank)
        # NCCL prints its version banner on stdout at communicator creation; keep stdout clean for the
        # single JSON line by pointing fd 1 at stderr until the result is printed
        self.saved_stdout = os.dup(1)
        os.dup2(2, 1)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=torch.device(self.dev))
        self.warmup = max(args.warmup, 3)
        from prisim_b200 import _lib
        self.ctx = _lib.get_context(self.local_rank)

    def fence(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
            self.torch.cuda.synchronize()

    def reduce(self, values, op):
        t = self.torch.tensor(values, dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=getattr(self.dist.ReduceOp, op))
        return t.tolist()

    def timed_loop(self, step, steps):
        """warm-up, then exactly `steps` steps between barrier + synchronize; returns (max-over-ranks ms, launches, clocks)."""
        torch = self.torch
        for _ in range(self.warmup):
            step(False)
        self.fence()
        sampler = ClockSampler(self.local_rank) if self.rank == 0 else None
        if sampler:
            sampler.start()
        launches0 = self.ctx.launches
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(steps):
            step(True)
        t1.record()
        self.fence()
        elapsed = self.reduce([t0.elapsed_time(t1)], "MAX")[0]
        clocks = sampler.stop() if sampler else None
        return elapsed, self.ctx.launches - launches0, clocks

    def wall_loop(self, step, steps, warm=2, tail=None):
        for _ in range(warm):
            step()
        if tail:
            tail()
        self.fence()
        w0 = time.perf_counter()
        for _ in range(steps):
            step()
        if tail:
            tail()                                            # e.g. wait for the device->host copies still in flight
        self.fence()
        return self.reduce([(time.perf_counter() - w0) * 1e3 / steps], "MAX")[0]

    def emit(self, line):
        if self.rank == 0:
            sys.stdout.flush()
            os.dup2(self.saved_stdout, 1)
            print(json.dumps(line))
            sys.stdout.flush()

    def finish(self):
        if self.world > 1:
            self.dist.barrier()
            self.dist.destroy_process_group()


def fma_roofline(env, kernel, terms_per_launch, kernel_ms, step_ms, clocks, lane_issues_per_term, pipe, algorithmic_bytes, traffic):
    """Issue roofline of a phase-sum kernel: `lane_issues_per_term` lane-issues of `pipe` ('fp32': 128 lanes/clk/SM,
    'fp64': 64 lanes/clk/SM) per term; peak at the SM clock sampled during the timed loop."""
    from prisim_b200 import engine
    mb = engine.microbench(env.local_rank)
    clock_hz = 1e6 * (clocks["sm_mhz"] if clocks and clocks.get("sm_mhz") else 1965.0)
    lanes = NOMINAL_LANES if pipe == "fp32" else NOMINAL_LANES // 2
    flop_per_term = 2 * lane_issues_per_term
    terms_per_s = terms_per_launch / (kernel_ms * 1e-3)
    achieved = flop_per_term * terms_per_s / 1e12
    peak = lanes * 2 * clock_hz / 1e12
    measured_lanes = mb["ffma_lanes_per_clk_per_sm"] * 148 if pipe == "fp32" else mb["dfma_per_s"] / mb["sm_clock_hz"]
    peaks_file = os.path.join(ROOT, "MEASURED_PEAKS.json")
    hbm = json.load(open(peaks_file)).get("hbm_gbs") if os.path.exists(peaks_file) else None
    return {"bound": "{0}_fma_issue".format(pipe), "kernel": kernel, "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
            "frac": achieved / peak,
            "peak_source": "nominal {0} lanes/clk x 2 x SM clock sampled during the timed loop ({1:.0f} MHz)".format(lanes, clock_hz / 1e6),
            "frac_of_measured_pipe_rate": achieved / (measured_lanes * 2 * clock_hz / 1e12),
            "measured_lanes_per_clk": measured_lanes, "nominal_lanes_per_clk": lanes,
            "frac_at_max_clock_1965": achieved / (lanes * 2 * 1.965e9 / 1e12),
            "flop_equiv_per_term": flop_per_term, "kernel_ms": kernel_ms, "kernel_share_of_step": kernel_ms / step_ms,
            "kernel_gterms_per_s": terms_per_s / 1e9, "traffic": traffic, "algorithmic_bytes": algorithmic_bytes,
            "hbm_gbs_measured": hbm, "mufu_per_s": mb["mufu_per_s"], "dfma_per_s": mb["dfma_per_s"],
            "microbench_sm_clock_hz": mb["sm_clock_hz"]}


def cpu_baseline(cfg, terms_step):
    cores = min(host_cores(), 64)
    rate, dt, tterms = cpu_sample(cfg, 8000, max(cores * 32, 64), cores)
    return {"value": rate / 1e9, "unit": "Gterms/s", "cores": cores, "kind": "port", "seconds": dt,
            "mterms_per_s_per_core": rate / 1e6 / cores, "extrapolated_full_config_seconds": terms_step / rate,
            "sample": "8000 above-horizon sources x {0} baselines x 1024 channels ({1:.2e} terms) of the config-2 workload, "
                      "float64 numpy restatement of interferometry.py:6332-6340, {2} processes".format(max(cores * 32, 64), tterms, cores)}


def bit_checksum(t):
    """Order-independent exact checksum of a tensor's bit patterns (wrapping int64 sum)."""
    import torch
    return int(torch.view_as_real(t).contiguous().view(torch.int64).sum().item()) if t.numel() else 0


def pinned(a):
    import torch
    return torch.from_numpy(NP.ascontiguousarray(a, dtype=NP.float64)).pin_memory().numpy()


# --------------------------------------------------------------------------------------------
# config 2 (headline) and config 5 (pipeline): HERA-350 x 300k sources x 1024 channels
# --------------------------------------------------------------------------------------------
def run_config2(env, pipeline=False):
    torch, dist, args = env.torch, env.dist, env.args
    from prisim_b200 import engine
    from prisim_b200 import primary_beams as PB
    from prisim_b200 import synthetic as S
    from prisim_b200.interferometry import InterferometerArray, SimpleTime
    from prisim_b200.sharding import PeerGatherBuffer, ShardedObserver, shard_rows
    world, rank, lr, dev = env.world, env.rank, env.local_rank, env.dev
    strong = args.scaling in ("auto", "strong")

    cfg = S.config5(nsrc=args.nsrc) if pipeline else S.config2(nsrc=args.nsrc)
    sky = cfg["skymodel"]
    sp = sky.spec_parms
    nbl, nchan = cfg["baselines"].shape[0], cfg["channels"].size
    lst_deg = 0.0 if strong else 0.0 + 15.0 * rank / 8.0       # weak: every rank observes its own snapshot
    # strong scaling: rank r owns baselines r, r + world, ... (interleaved: the short, strongly cancelling baselines that the
    # precision control recomputes in fp64 sit at the front of PRISim's length-sorted list and would all land on rank 0)
    sl = shard_rows(nbl, world, rank, interleave=True) if strong else slice(0, nbl)
    nbl_local = len(range(*sl.indices(nbl)))

    # ---- resident inputs ----
    d_hadec = engine._f64(NP.stack((lst_deg - sky.location[:, 0], sky.location[:, 1]), axis=1), lr)
    spec = {"flux_scale": engine._f64(sp["flux-scale"], lr), "index": engine._f64(sp["power-law-index"], lr),
            "freq_ref": engine._f64(sp["freq-ref"], lr)}
    d_bl = engine._f64(cfg["baselines"][sl], lr)
    pc_dircos = NP.asarray([0.0, 0.0, 1.0])
    beam = PB.beam_desc_from_telescope(cfg["telescope"], pointing_center=NP.asarray([90.0, 270.0]), skyunits="altaz", device=lr)
    # multi-GPU: the phase-sum kernel's epilogue stores each rank's rows straight into rank 0's buffer over NVLink
    # peer memory (sharding.PeerGatherBuffer); NCCL point-to-point is the fallback if mapping fails
    gbuf = None
    if world > 1:
        gbuf = PeerGatherBuffer((nbl, nchan), lr, dst=0, interleave=True) if strong else PeerGatherBuffer((nbl, nchan), lr, dst=0)
    vis = gbuf.local if gbuf is not None else torch.empty((nbl, nchan), dtype=torch.complex128, device=dev)
    k1_events = []
    window = None
    if pipeline:
        from prisim_b200.delay_spectrum import windowing
        window = engine._f64(nchan * windowing(nchan, "bhw", area_normalize=True), lr)
        tsys = engine._f64(cfg["Tsysinfo"]["Trx"] + cfg["Tsysinfo"]["Tant"]["T0"] * (cfg["channels"] / cfg["Tsysinfo"]["Tant"]["f0"]) ** cfg["Tsysinfo"]["Tant"]["spindex"], lr)
        aeff, effq = engine._f64([cfg["A_eff"]], lr), engine._f64([cfg["eff_Q"]], lr)
        bp = engine._f64(NP.ones(nchan), lr)
        df = float(cfg["channels"][1] - cfg["channels"][0])
        tail_events = []

    def tables():
        """cull + brightest-first order + amplitude table, as InterferometerArray.observe does them"""
        dircos, index = engine.sky_cull(d_hadec, "hadec", latitude_deg=cfg["latitude"], device=lr)
        nsrc = int(index.shape[0])
        perm, nbright = engine.brightness_order(dircos, index, nsrc, spec, beam, cfg["channels"], device=lr)
        dircos, index = dircos.index_select(0, perm).contiguous(), index.index_select(0, perm).contiguous()
        return dircos, index, nsrc, nbright

    def step(timed):
        dircos, index, nsrc, nbright = tables()
        amp = engine.amp_table(dircos, index, nsrc, spec, beam, cfg["channels"], device=lr)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        engine.skyvis(dircos, amp, nsrc, d_bl, pc_dircos, cfg["channels"], out=vis, device=lr, nsrc_bright=nbright)
        e1.record()
        if timed:
            k1_events.append((e0, e1))
        if pipeline:       # the rank-0 tail of run_prisim.py:2278-2284 on this rank's rows: noise, add, three delay transforms
            e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e2.record()
            sky_local = vis if vis.is_contiguous() else vis.contiguous()
            _, nz, v = engine.noise(sky_local, tsys, aeff, effq, df, cfg["t_acc"], 5, nbl_local, nchan, snapshot=0, bl_offset=sl.start or 0,
                                    bl_step=sl.step or 1, nbl_total=nbl, want=("noise", "vis"), device=lr)
            for x in (sky_local, v, nz):
                engine.delay_transform(x, bp, window, df, pad=1.0, downsample=True)
            e3.record()
            if timed:
                tail_events.append((e2, e3))
        if gbuf is not None:                              # the single gather of the path (to the writing rank)
            gbuf.wait()
        return nsrc

    nsrc = step(False)
    elapsed_ms, launches, clocks = env.timed_loop(step, args.steps)
    k1_ms = statistics.mean(a.elapsed_time(b) for a, b in k1_events)
    terms_local = float(nsrc) * nbl_local * nchan
    k1_ms_max, = env.reduce([k1_ms], "MAX")
    terms_sum, = env.reduce([terms_local], "SUM")
    terms_step = terms_sum                                 # strong: the one snapshot; weak: world snapshots
    ms_per_step = elapsed_ms / args.steps
    value = terms_step / (ms_per_step * 1e-3) / 1e9

    # ---- correctness of what rank 0 holds (after the timed loop, same inputs) ----
    check = {"ranks": world}
    dircos, index, nsrc, nbright = tables()
    amp = engine.amp_table(dircos, index, nsrc, spec, beam, cfg["channels"], device=lr)
    private = engine.skyvis(dircos, amp, nsrc, d_bl, pc_dircos, cfg["channels"], device=lr, nsrc_bright=nbright)
    engine.skyvis(dircos, amp, nsrc, d_bl, pc_dircos, cfg["channels"], out=vis, device=lr, nsrc_bright=nbright)
    if gbuf is not None:
        gbuf.wait()
    mine = torch.tensor([bit_checksum(private)], dtype=torch.int64, device=dev)
    sums = [torch.zeros_like(mine) for _ in range(world)]
    if world > 1:
        dist.all_gather(sums, mine)
    else:
        sums = [mine]
    if rank == 0:
        full = gbuf.full if gbuf is not None else vis
        if strong:
            got = [bit_checksum(full[r::world]) for r in range(world)]
        else:
            got = [bit_checksum(full[r]) for r in range(world)] if world > 1 else [bit_checksum(full)]
        check["checksums_equal"] = [int(s.item()) == g for s, g in zip(sums, got)]
        check["ok"] = all(check["checksums_equal"])
        check["method"] = "int64 bit-pattern checksum of every rank's block recomputed into private memory == checksum of the same rows in rank 0's buffer"
        if strong:      # 64 rows spread over all shards against the fp64 kernel
            rows = torch.linspace(0, nbl - 1, 64, device=dev).long()
            amp64 = engine.amp_table(dircos, index, nsrc, spec, beam, cfg["channels"], device=lr, dtype=torch.float64)
            ref = engine.skyvis(dircos, amp64, nsrc, engine._f64(cfg["baselines"], lr).index_select(0, rows).contiguous(), pc_dircos,
                                cfg["channels"], method="fp64", device=lr)
            got_rows = full.index_select(0, rows)
            rms = torch.sqrt(ref.real.square().mean(dim=1) + ref.imag.square().mean(dim=1))
            check["fp64_rows"] = 64
            check["fp64_max_err_over_rms_b"] = float(((got_rows - ref).abs().amax(dim=1) / rms).max().item())
            check["ok"] = check["ok"] and check["fp64_max_err_over_rms_b"] <= 1e-5
            del amp64, ref
    del private, amp

    # ---- end-to-end through the public API with host buffers ----
    e2e = None
    if not args.no_e2e:
        sky.location = pinned(sky.location)
        for key in ("flux-scale", "power-law-index", "freq-ref", "flux-offset"):
            sp[key] = pinned(sp[key])
        kw = dict(telescope=cfg["telescope"], latitude=cfg["latitude"], skycoords="radec", pointing_coords="hadec", device=lr, noise_seed=5)
        if pipeline:
            kw.update(A_eff=cfg["A_eff"], eff_Q=cfg["eff_Q"])
        h2d = sky.location.nbytes + sum(sp[k].nbytes for k in ("flux-scale", "power-law-index", "freq-ref"))
        tsysinfo = cfg["Tsysinfo"] if pipeline else {"Tnet": 300.0}
        if strong:
            so = ShardedObserver(InterferometerArray, cfg["labels"], cfg["baselines"], cfg["channels"], **kw)
            ia = so.ia
        else:
            so, ia = None, InterferometerArray(cfg["labels"], cfg["baselines"], cfg["channels"], **kw)
        ia.cache_sky = False                              # force the host->device copy of the catalogue every step
        # pinned host destinations: the gathered skyvis on the writing rank (every rank in weak mode); pipeline: this rank's
        # rows of vis, noise and the three delay spectra as well
        host = [torch.empty((nbl if (rank == 0 or not strong) else 0, nchan), dtype=torch.complex128, pin_memory=True)]
        if pipeline:
            host += [torch.empty((nbl_local, nchan), dtype=torch.complex128, pin_memory=True) for _ in range(5)]
        d2h = sum(h.numel() * 16 for h in host)
        win_host = None if window is None else window.cpu().numpy()

        # device->host copies run on their own stream into two sets of pinned buffers used in turn, so the copy of snapshot j
        # overlaps the kernels of snapshot j + 1 (a streamed multi-snapshot run); every copy happens inside the timed region
        # and the clock stops only after the last one has landed
        host2 = [host, [torch.empty(h.shape, dtype=h.dtype, pin_memory=h.numel() > 0) for h in host]]
        copy_stream = torch.cuda.Stream(device=lr)
        copy_done = [None, None]
        turn = [0]

        def copy_out(dst, src, torch_owned=True):
            if torch_owned:                                   # allocator-owned tensors may be freed before the copy ran; the gather
                src.record_stream(copy_stream)                # buffers are library-owned and protected by copy_done instead
            dst.copy_(src, non_blocking=True)

        def e2e_step():
            k = turn[0] % 2
            turn[0] += 1
            if copy_done[k] is not None:
                copy_done[k].synchronize()                    # pinned set / gather buffer / product ring slot k (used two snapshots ago) are free again
            hk = host2[k]
            targs = (SimpleTime(2451545.0, lst_deg), tsysinfo, NP.ones(nchan), cfg["pointing_hadec"], sky, cfg["t_acc"])
            full = so.observe(*targs) if so is not None else (ia.observe(*targs), ia.skyvis_freq_device(-1))[1]
            prods = []
            # stream the snapshot out of the array (bounded memory); pipeline: noise + add + three delay transforms on this rank's
            # rows into the array's preallocated product ring (slot = snapshot number % 2 = k)
            ia.drain(lambda j, prod: prods.extend(prod[key] for key in ("vis_freq", "vis_noise_freq", "skyvis_lag", "vis_lag", "vis_noise_lag") if key in prod),
                     noise=pipeline, delay_transform={"pad": 1.0, "freq_wts": win_host} if pipeline else None, ring=2)
            ready = torch.cuda.Event()
            ready.record()
            copy_stream.wait_event(ready)
            with torch.cuda.stream(copy_stream):
                if full is not None and hk[0].numel():
                    copy_out(hk[0], full, torch_owned=so is None or world == 1)   # ONE device->host copy of the snapshot, from the writing rank
                for i, t in enumerate(prods):
                    copy_out(hk[1 + i], t, torch_owned=False)
                copy_done[k] = torch.cuda.Event()
                copy_done[k].record()

        def e2e_fence():
            for ev in copy_done:
                if ev is not None:
                    ev.synchronize()

        e2e_ms = env.wall_loop(e2e_step, args.steps, tail=e2e_fence)
        d2h_all, = env.reduce([float(d2h)], "SUM")
        e2e = {"value": terms_step / (e2e_ms * 1e-3) / 1e9, "unit": "Gterms/s", "h2d_bytes_per_step": int(h2d) * (world if strong else 1),
               "d2h_bytes_per_step": int(d2h_all if strong else d2h), "ms_per_step": e2e_ms,
               "api": ("sharding.ShardedObserver.observe (interleaved baseline shards, kernel epilogue stores into rank 0's buffer) + ONE device->host copy of "
                       "the gathered skyvis_freq from rank 0 (pinned)" if strong and world > 1 else
                       "InterferometerArray.observe + device->host copy of skyvis_freq (pinned)") +
                      "; precision='auto': fp32 kernel + fp64 recompute of cancelling baselines + sampled fp64 audit" +
                      ("; then InterferometerArray.drain (noise + add + three delay transforms) and device->host copies of those five products" if pipeline else ""),
               "precision_report": ia.precision_report[-1] if ia.precision_report else None}
        if so is not None:
            so.close()

    if rank == 0:
        roofline = fma_roofline(env, "k_skyvis", terms_local, k1_ms, ms_per_step, clocks, 6, "fp32",
                                float(nsrc) * nchan * 4 + nbl_local * nchan * 16.0, TRAFFIC_BYTES_PER_LAUNCH if world == 1 else None)
        roofline["kernel_ms_max_over_ranks"] = k1_ms_max
        if pipeline:
            tail_ms = statistics.mean(a.elapsed_time(b) for a, b in tail_events)
            tail_bytes = nbl_local * nchan * (16.0 * 3 + 8.0 + 3 * 32.0)      # noise: read skyvis, write noise + vis (+ Tsys row); 3 transforms: 16 B in + 16 B out
            roofline["tail"] = {"kernels": "k_noise + 3 x k_delay_fft_w32", "ms": tail_ms, "algorithmic_bytes": tail_bytes,
                                "achieved_gbs": tail_bytes / (tail_ms * 1e-3) / 1e9, "peak_gbs": roofline["hbm_gbs_measured"],
                                "frac": tail_bytes / (tail_ms * 1e-3) / 1e9 / roofline["hbm_gbs_measured"] if roofline["hbm_gbs_measured"] else None}
        cpu = None if (args.no_cpu_baseline or world > 1) else cpu_baseline(cfg, terms_step)     # the CPU arm is timed at N=1 only
        mode = ("single GPU" if world == 1 else
                ("interleaved baseline shards of ONE snapshot, " if strong else "one snapshot per GPU, ") +
                ("kernel epilogue stores over NVLink peer memory into rank 0's buffer" if gbuf.mode == "peer" else "NCCL point-to-point gather to rank 0"))
        line = {"metric": METRIC, "value": value, "unit": "Gterms/s", "n_gpus": world, "steps": args.steps, "warmup": env.warmup,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": WORKLOADS[5 if pipeline else 2], "nsrc_catalogue": args.nsrc, "nsrc_above_horizon": nsrc, "nbl": nbl,
                           "nchan": nchan, "terms_per_step": terms_step, "baselines_on_rank0": nbl_local, "sharding": mode,
                           "schedule": "persistent CTAs (1 per SM): whole output tiles in lock-step waves + stream-K split of the tail tiles along the source axis",
                           "l2": "inputs larger than L2: amplitude table {0:.2f} GB + {1:.2f} GB output per step and GPU".format(
                               nsrc * nchan * 4 / 1e9, nbl_local * nchan * 16 / 1e9),
                           "phase_arith": "fp64-reduced MUFU anchors for both channels of a pair, fp32 rotation recurrence by r^2, fp32 accumulate flushed to fp64 "
                                          "every 256 sources (every 32 for the brightest, which are summed first)"},
                "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
                "gather_check": check}
        env.emit(line)
    if gbuf is not None:
        gbuf.close()


# --------------------------------------------------------------------------------------------
# config 3: diffuse sky, fp64 + taper kernel, gridded HEALPix beam
# --------------------------------------------------------------------------------------------
def run_config3(env):
    torch, args = env.torch, env.args
    from prisim_b200 import engine
    from prisim_b200 import primary_beams as PB
    from prisim_b200 import synthetic as S
    from prisim_b200.interferometry import InterferometerArray, SimpleTime
    world, rank, lr, dev = env.world, env.rank, env.local_rank, env.dev
    cfg = S.config3(nside=args.nside, nsnap=1)
    sky = cfg["skymodel"]
    sp = sky.spec_parms
    nbl, nchan = cfg["baselines"].shape[0], cfg["channels"].size
    lst_deg = 20.0 + cfg["t_acc"] / 240.0 * rank               # snapshots dealt round-robin: rank r takes snapshot r of the drift scan
    d_hadec = engine._f64(NP.stack((lst_deg - sky.location[:, 0], sky.location[:, 1]), axis=1), lr)
    spec = {"flux_scale": engine._f64(sp["flux-scale"], lr), "index": engine._f64(sp["power-law-index"], lr),
            "freq_ref": engine._f64(sp["freq-ref"], lr)}
    d_fwhm = engine._f64(NP.sqrt(sky.src_shape[:, 0] * sky.src_shape[:, 1]), lr)
    d_bl = engine._f64(cfg["baselines"], lr)
    pc_dircos = NP.asarray([0.0, 0.0, 1.0])
    # gridded beam: the Airy pattern sampled on an nside-128 HEALPix grid, log10, at 16 frequencies (host, once)
    bf = NP.linspace(cfg["channels"][0], cfg["channels"][-1], 16)
    az, alt = S.healpix_ring_centers(128)                   # beam frame: pole = zenith, longitude = azimuth
    gridded = NP.full((az.size, bf.size), 1e-12)
    up = alt > 0.0
    gridded[up] = NP.maximum(PB.primary_beam_generator(NP.stack((alt[up], az[up]), axis=1), bf, cfg["telescope"], freq_scale="Hz", skyunits="altaz",
                                                       pointing_center=NP.asarray([90.0, 270.0]), device=lr), 1e-12)
    hb = PB.HealpixBeam(gridded, bf, cfg["channels"], spec_interp="cubic", device=lr)
    vis = torch.empty((nbl, nchan), dtype=torch.complex128, device=dev)
    ev = {"k1": [], "hp": []}

    def step(timed):
        dircos, index = engine.sky_cull(d_hadec, "hadec", latitude_deg=cfg["latitude"], device=lr)
        nsrc = int(index.shape[0])
        e0, e1, e2, e3 = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        e0.record()
        logbeam, logmax = hb.table(dircos, nsrc)
        e1.record()
        beam = engine.make_beam_desc(element=engine._lib.BEAM_LOGTABLE, d_logmax=logmax)
        amp = engine.amp_table(dircos, index, nsrc, spec, beam, cfg["channels"], pbeam=logbeam, device=lr, dtype=torch.float64)
        fw = d_fwhm.index_select(0, index.long()).contiguous()
        e2.record()
        engine.skyvis(dircos, amp, nsrc, d_bl, pc_dircos, cfg["channels"], src_fwhm_deg=fw, method="fp64", out=vis, device=lr)
        e3.record()
        if timed:
            ev["k1"].append((e2, e3)); ev["hp"].append((e0, e1))
        return nsrc

    nsrc = step(False)
    elapsed_ms, launches, clocks = env.timed_loop(step, args.steps)
    k1_ms = statistics.mean(a.elapsed_time(b) for a, b in ev["k1"])
    hp_ms = statistics.mean(a.elapsed_time(b) for a, b in ev["hp"])
    terms_local = float(nsrc) * nbl * nchan
    terms_step, = env.reduce([terms_local], "SUM")
    ms_per_step = elapsed_ms / args.steps
    value = terms_step / (ms_per_step * 1e-3) / 1e9

    e2e = None
    if not args.no_e2e:
        sky.location = pinned(sky.location)
        for key in ("flux-scale", "power-law-index", "freq-ref", "flux-offset"):
            sp[key] = pinned(sp[key])
        ia = InterferometerArray(cfg["labels"], cfg["baselines"], cfg["channels"], telescope=cfg["telescope"], latitude=cfg["latitude"],
                                 skycoords="radec", pointing_coords="hadec", device=lr, noise_seed=5)
        ia.cache_sky = False
        ia.precision = "fp64"
        host_vis = torch.empty((nbl, nchan), dtype=torch.complex128, pin_memory=True)
        h2d = sky.location.nbytes + sum(sp[k].nbytes for k in ("flux-scale", "power-law-index", "freq-ref")) + sky.src_shape.nbytes

        def e2e_step():
            ia._skyvis, ia._bp, ia._Tsys, ia.timestamp, ia.t_acc, ia.lst = [], [], [], [], [], []
            ia.obs_catalog_indices, ia.n_acc = [], 0
            ia.observe(SimpleTime(2451545.0, lst_deg), {"Tnet": 300.0}, NP.ones(nchan), cfg["pointing_hadec"], sky, cfg["t_acc"],
                       pb_info={"external_beam": hb})
            host_vis.copy_(ia.skyvis_freq_device(0), non_blocking=True)
            torch.cuda.synchronize()

        e2e_ms = env.wall_loop(e2e_step, max(2, args.steps // 2), warm=1)
        e2e = {"value": terms_step / (e2e_ms * 1e-3) / 1e9, "unit": "Gterms/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(host_vis.numel() * 16),
               "ms_per_step": e2e_ms, "api": "InterferometerArray.observe(pb_info={'external_beam': HealpixBeam}), precision='fp64', + device->host copy of skyvis_freq (pinned)"}

    if rank == 0:
        roofline = fma_roofline(env, "k_skyvis_fp64<double, taper>", terms_local, k1_ms, ms_per_step, clocks, 6, "fp64",
                                float(nsrc) * nchan * 8 + nbl * nchan * 16.0, None)
        gather_bytes = float(nsrc) * nchan * (4 * hb.map.element_size() + 8.0)         # four map rows read + one fp64 row written per (source, channel)
        roofline["healpix_gather"] = {"kernel": "k_healpix_gather + k_colmax", "ms": hp_ms, "algorithmic_bytes": gather_bytes,
                                      "achieved_gbs": gather_bytes / (hp_ms * 1e-3) / 1e9, "peak_gbs": roofline["hbm_gbs_measured"],
                                      "frac": gather_bytes / (hp_ms * 1e-3) / 1e9 / roofline["hbm_gbs_measured"] if roofline["hbm_gbs_measured"] else None}
        line = {"metric": "Gterms/s (src x bl x chan), config 3", "value": value, "unit": "Gterms/s", "n_gpus": world, "steps": args.steps,
                "warmup": env.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": WORKLOADS[3], "nsrc_above_horizon": nsrc, "nbl": nbl, "nchan": nchan, "terms_per_step": terms_step,
                           "sharding": "snapshots round-robin over ranks (each rank its own cull + beam gather + amplitude table)",
                           "l2": "inputs larger than L2: fp64 amplitude table {0:.2f} GB per step".format(nsrc * nchan * 8 / 1e9),
                           "phase_arith": "fp64 throughout: block anchors by complex powering, three-term phasor recurrence, taper by second-order recurrence (7 DFMA per term)"},
                "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": None}
        env.emit(line)


# --------------------------------------------------------------------------------------------
# config 4: MWA-like tiles, phased-array tile beam fused into the amplitude table
# --------------------------------------------------------------------------------------------
def run_config4(env):
    torch, args = env.torch, env.args
    from prisim_b200 import engine
    from prisim_b200 import geometry as GEOM
    from prisim_b200 import primary_beams as PB
    from prisim_b200 import synthetic as S
    from prisim_b200.interferometry import InterferometerArray, SimpleTime
    world, rank, lr, dev = env.world, env.rank, env.local_rank, env.dev
    cfg = S.config4()
    sky, sp = cfg["skymodel"], cfg["skymodel"].spec_parms
    nbl, nchan = cfg["baselines"].shape[0], cfg["channels"].size
    lst_deg = 50.0 + 2.0 * rank                                # snapshots dealt round-robin: every rank its own LST
    d_hadec = engine._f64(NP.stack((lst_deg - sky.location[:, 0], sky.location[:, 1]), axis=1), lr)
    spec = {"flux_scale": engine._f64(sp["flux-scale"], lr), "index": engine._f64(sp["power-law-index"], lr),
            "freq_ref": engine._f64(sp["freq-ref"], lr)}
    d_bl = engine._f64(cfg["baselines"], lr)
    pc_altaz = cfg["pointing_altaz"]
    pc_dircos = GEOM.altaz2dircos(pc_altaz, "degrees")[0]
    beam = PB.beam_desc_from_telescope(cfg["telescope"], pointing_info=cfg["pb_info"], pointing_center=pc_altaz, skyunits="altaz", device=lr)
    vis = torch.empty((nbl, nchan), dtype=torch.complex128, device=dev)
    ev = {"k1": [], "amp": []}

    def step(timed):
        dircos, index = engine.sky_cull(d_hadec, "hadec", latitude_deg=cfg["latitude"], device=lr)
        nsrc = int(index.shape[0])
        perm, nbright = engine.brightness_order(dircos, index, nsrc, spec, beam, cfg["channels"], device=lr)
        dircos, index = dircos.index_select(0, perm).contiguous(), index.index_select(0, perm).contiguous()
        e0, e1, e2 = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e0.record()
        amp = engine.amp_table(dircos, index, nsrc, spec, beam, cfg["channels"], device=lr)
        e1.record()
        engine.skyvis(dircos, amp, nsrc, d_bl, pc_dircos, cfg["channels"], out=vis, device=lr, nsrc_bright=nbright)
        e2.record()
        if timed:
            ev["amp"].append((e0, e1)); ev["k1"].append((e1, e2))
        return nsrc

    nsrc = step(False)
    elapsed_ms, launches, clocks = env.timed_loop(step, args.steps)
    k1_ms = statistics.mean(a.elapsed_time(b) for a, b in ev["k1"])
    amp_ms = statistics.mean(a.elapsed_time(b) for a, b in ev["amp"])
    terms_local = float(nsrc) * nbl * nchan
    terms_step, = env.reduce([terms_local], "SUM")
    ms_per_step = elapsed_ms / args.steps
    value = terms_step / (ms_per_step * 1e-3) / 1e9
    e2e = None
    if not args.no_e2e:
        sky.location = pinned(sky.location)
        for key in ("flux-scale", "power-law-index", "freq-ref", "flux-offset"):
            sp[key] = pinned(sp[key])
        ia = InterferometerArray(cfg["labels"], cfg["baselines"], cfg["channels"], telescope=cfg["telescope"], latitude=cfg["latitude"],
                                 skycoords="radec", pointing_coords="altaz", device=lr, noise_seed=5)
        ia.cache_sky = False
        host_vis = torch.empty((nbl, nchan), dtype=torch.complex128, pin_memory=True)
        h2d = sky.location.nbytes + sum(sp[k].nbytes for k in ("flux-scale", "power-law-index", "freq-ref"))

        def e2e_step():
            ia.observe(SimpleTime(2451545.0, lst_deg), {"Tnet": 200.0}, NP.ones(nchan), pc_altaz, sky, cfg["t_acc"], pb_info=cfg["pb_info"])
            ia.drain(lambda j, prod: host_vis.copy_(prod["skyvis_freq"], non_blocking=True), noise=False)
            torch.cuda.synchronize()

        e2e_ms = env.wall_loop(e2e_step, args.steps)
        e2e = {"value": terms_step / (e2e_ms * 1e-3) / 1e9, "unit": "Gterms/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(host_vis.numel() * 16),
               "ms_per_step": e2e_ms, "api": "InterferometerArray.observe(pb_info={delays, ...}) + device->host copy of skyvis_freq (pinned); precision='auto'",
               "precision_report": ia.precision_report[-1] if ia.precision_report else None}
    if rank == 0:
        roofline = fma_roofline(env, "k_skyvis", terms_local, k1_ms, ms_per_step, clocks, 6, "fp32",
                                float(nsrc) * nchan * 4 + nbl * nchan * 16.0, None)
        n_el = 16
        roofline["tile_beam"] = {"kernel": "k_amp_table<float> (dipole x 16-element phased array x ground plane, fp64 evaluation)", "ms": amp_ms,
                                 "source_channel_elements_per_s": float(nsrc) * nchan / (amp_ms * 1e-3),
                                 "element_phasors_per_s": float(nsrc) * nchan * n_el / (amp_ms * 1e-3),
                                 "share_of_step": amp_ms / ms_per_step}
        line = {"metric": "Gterms/s (src x bl x chan), config 4", "value": value, "unit": "Gterms/s", "n_gpus": world, "steps": args.steps,
                "warmup": env.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": WORKLOADS[4], "nsrc_above_horizon": nsrc, "nbl": nbl, "nchan": nchan, "terms_per_step": terms_step,
                           "sharding": "snapshots round-robin over ranks",
                           "l2": "L2 flushed between steps by the step's own streams: amplitude table {0:.2f} GB + {1:.2f} GB running sums".format(
                               nsrc * nchan * 4 / 1e9, nbl * nchan * 16 / 1e9)},
                "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": None}
        env.emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=[2, 3, 4, 5])
    ap.add_argument("--scaling", default="auto", choices=["auto", "strong", "weak"],
                    help="configs 2/5 at N>1: strong (default) = baseline blocks of one snapshot; weak = one snapshot per rank")
    ap.add_argument("--nsrc", type=int, default=300000, help="catalogue size (default = the headline 300k)")
    ap.add_argument("--nside", type=int, default=256, help="config 3: HEALPix nside of the diffuse sky (default = the configuration's 256)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
        return
    env = Env(args)
    if args.config == 3:
        run_config3(env)
    elif args.config == 4:
        run_config4(env)
    else:
        run_config2(env, pipeline=(args.config == 5))
    env.finish()


# dram__bytes_read.sum + dram__bytes_write.sum per k_skyvis launch at the headline size (1 GPU), from the
# committed ncu capture (profiles/); None until that capture exists for the current kernel.
TRAFFIC_BYTES_PER_LAUNCH = 142.8e9     # profiles/skyvis_dram_r02.txt: 128.8 GB read + 14.0 GB written (0.9 % of HBM bandwidth over 2.37 s; final kernel)

if __name__ == "__main__":
    main()
