#!/usr/bin/env python
"""bench.py -- reaches*timesteps/s of the routing hot path (KWT+IRF) on synthetic river networks.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload C4|C2|C3|C5] [--tsteps T] [--impl reference]

One bench "step" = one pass of the hot path over one batch of synthetic forcing: T consecutive routing time
steps (`--tsteps`) of the whole network through mr_route_resident / mr_step_batch (the reference's
main_route called T times).  value = nRch * T * K / time.

Workloads (BASELINE.json configs):
    C4 (default)  3M-reach CONUS-like forest, IRF+KWT (route_opt 12), hourly, hillslope UH on
    C2            100k-reach binary tree, IRF, hourly
    C3            3M-reach CONUS-like forest, KWT, daily
    C5            C3 + 10k lakes (Doll / endorheic)

N > 1 (torchrun, one rank per GPU): the SAME network is decomposed by the reference's rule (partition.decompose:
reaches with more than nRch/N upstream reaches are mainstem, the subtrees hanging off it and the smaller basins
are tributary domains bin-packed onto the ranks); every rank routes its tributaries, the tributary outlets are
handed to rank 0 by NCCL send/recv (multi.exchange_rows), rank 0 routes the mainstem; time = max over ranks.

`--impl reference` times the CPU restatement of the reference algorithm (oracle/, OpenMP over all host cores)
on a bounded sample of the same workload -- the Fortran reference cannot be built in this image (no Fortran
compiler / MPI / netCDF), see DESIGN.md.
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

from mizuroute_b200 import partition, synth  # noqa: E402
from mizuroute_b200.network import RouteOptions, RouteParams  # noqa: E402

METRIC = "reaches*timesteps/sec (KWT+IRF)"
UNIT = "reach-steps/s"

WORKLOADS = {
    #        kind      n          route  dt       lakes  default T
    "C2": ("binary", 100_000, "1", 3600.0, 0, 240),
    "C3": ("conus", 3_000_000, "2", 86400.0, 0, 256),
    "C4": ("conus", 3_000_000, "12", 3600.0, 0, 768),
    "C5": ("conus", 3_000_000, "2", 86400.0, 10_000, 256),
}


def make_workload(name: str, n_override: int | None):
    kind, n, route, dt, lakes, T = WORKLOADS[name]
    if n_override:
        n = n_override
    net = synth.binary_tree(n, seed=2) if kind == "binary" else synth.conus_like(n, seed=3)
    opts = RouteOptions(dt=dt, route_opt=route, runoffMin=1e-15)
    if lakes:
        synth.add_lakes(net, lakes, np.random.default_rng(103))
        opts.is_lake_sim = True
        opts.LakeInputOption = 1
    return net, RouteParams(), opts, T


def runoff_for(net, T, dt):
    return synth.runoff_series(net, T, seed=11, dt=dt)


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md)."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index: int):
        self.idx, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._pump, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            v = d.get("hbm_gbs") or d.get("hbm_gb_s")
            if v:
                return float(v), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(workload: str, kernel: str):
    """dram bytes per launch of the dominant kernel from the committed ncu capture (profiles/traffic.json)."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get(workload, {}).get(kernel)
        except Exception:
            return None
    return None


# ------------------------------------------------------------------------------------------------
def run_reference(args, net, params, opts, T):
    """CPU arm: the oracle restatement with all host threads, each step a bounded sample of the workload."""
    from oracle.oracle import Oracle
    sample_T = max(1, min(T, args.ref_tsteps))
    ro = runoff_for(net, sample_T, opts.dt)
    o = Oracle(net, params, opts, n_threads=host_cores())
    cores = pick_threads(o, ro)
    for _ in range(max(args.warmup - 1, 0)):
        o.run(ro, want_q=False)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        o.run(ro, want_q=False)
    dt_s = time.perf_counter() - t0
    units = net.nRch * sample_T * args.steps
    val = units / dt_s
    sample = f"{sample_T} routing time steps of the full {net.nRch}-reach network per bench step, from a cold start + {args.warmup} warm-up passes"
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt_s / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": args.workload, "nRch": net.nRch, "route_opt": opts.route_opt, "dt_qsim": opts.dt,
                   "timesteps_per_step": sample_T},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "note": "CPU restatement of the reference algorithm (reference Fortran not buildable in this image)"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def pick_threads(o, ro) -> int:
    """All host threads unless fewer are measurably faster (OpenMP barriers on an oversubscribed/SMT box)."""
    best, best_t = host_cores(), None
    for nt in sorted({host_cores(), max(1, host_cores() // 2), 1}, reverse=True):
        o.set_threads(nt)
        t0 = time.perf_counter()
        o.run(ro[:1], want_q=False)
        dt_s = time.perf_counter() - t0
        if best_t is None or dt_s < best_t:
            best, best_t = nt, dt_s
    o.set_threads(best)
    return best


def algorithmic_bytes(r, net, opts, T, kwt_touched_per_batch):
    """Algorithmic bytes per batch for each kernel family (DESIGN.md 'algorithmic bytes')."""
    from mizuroute_b200 import capi
    N = net.nRch
    nb = r.info(capi.INFO_NTDH_BAS)
    sum_ntdh = r.info(capi.INFO_SUM_NTDH)
    sum_nups = r.info(capi.INFO_SUM_NUPS)
    n_contrib = net.nHRU                                   # every HRU drains to one reach
    out = {"k_basin": T * (20 * n_contrib + 16 * N + ((16 * nb + 24) * N if opts.doesBasinRoute == 1 else 16 * N))}
    if "1" in opts.route_opt:
        out["k_route<IRF>"] = T * (24 * sum_ntdh + 12 * sum_nups + 76 * N)
    if "2" in opts.route_opt:
        out["k_route_kwt"] = 28 * kwt_touched_per_batch + T * (28 * sum_nups + 56 * N)
    if "0" in opts.route_opt:
        out["k_route<SUM>"] = T * (12 * sum_nups + 16 * N)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="C4", choices=sorted(WORKLOADS))
    ap.add_argument("--tsteps", type=int, default=0, help="routing time steps per bench step (0 = workload default)")
    ap.add_argument("--nrch", type=int, default=0, help="override the network size (development only)")
    ap.add_argument("--ref-tsteps", type=int, default=2, help="time steps per step of the CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    net, params, opts, T = make_workload(args.workload, args.nrch or None)
    if args.tsteps:
        T = args.tsteps
    elif args.workload == "C4" and not args.nrch:
        # the 768-step default holds 134 GB on the device and, with the pinned e2e buffers, 112 GB of host memory at the peak
        # (measured on a 196 GB host); a smaller host gets the 384-step batches (74 GB / 60 GB)
        try:
            import psutil
            if psutil.virtual_memory().available < 150 * 2 ** 30:
                T = 384
        except Exception:
            pass

    if args.impl == "reference":
        if rank == 0:
            run_reference(args, net, params, opts, T)
        return

    import torch
    import torch.distributed as dist
    from mizuroute_b200 import capi
    from mizuroute_b200.route import Router

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the routing path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- this rank's domains: tributary subtrees (+ the mainstem on rank 0), see mizuroute_b200/multi.py
    full_n = net.nRch
    dom = rm = None
    n_main = n_outlets = 0
    if world > 1:
        from mizuroute_b200.multi import DomainSet
        dom = DomainSet(net, params, opts, T, rank, world, device=local_rank)
        r, net_local, rm = dom.trib, dom.trib_net, dom.main
        n_main, n_outlets = int(dom.dec.mainstem.size), int(dom.dec.outlets.size)
        if r is None:
            raise SystemExit("bench.py: this rank holds no tributary domain (network too small for this many GPUs)")
    else:
        net_local = net
        r = Router(net_local, params, opts, device=local_rank, max_batch=T)
    # the forcing lives once on the host, in pinned memory (the e2e leg copies from it; `ro` is a numpy view of it)
    ro_pin = torch.from_numpy(runoff_for(net_local, T, opts.dt)).pin_memory()      # [T, nHRU_local]
    ro = ro_pin.numpy()
    ro_main = runoff_for(dom.main_net, T, opts.dt) if rm is not None else None
    stream = torch.cuda.Stream()
    stream_main = torch.cuda.Stream()
    r.set_stream(stream.cuda_stream)
    if rm is not None:
        rm.set_stream(stream_main.cuda_stream)
    has_kwt = "2" in opts.route_opt

    def route_all():
        """one bench step on this rank: tributaries, hand-off of the outlets to rank 0, mainstem (N > 1: enqueued
        without host waits, the mainstem on its own stream so that it overlaps the next batch's tributaries)"""
        if dom is None:
            r.route_resident(T)
        else:
            dom.route_resident_pipelined(T, stream, stream_main)

    def finish_all():
        if dom is not None:
            stream.wait_stream(stream_main)
            dom.wait()

    # ---- device-resident throughput (`value`): forcing already in HBM; [T x nHRU] doubles >> L2
    r.upload_runoff(ro)
    if rm is not None:
        rm.upload_runoff(ro_main)
    with torch.cuda.stream(stream):
        for _ in range(args.warmup):
            route_all()
        finish_all()
    if has_kwt:
        r.set_counting(True)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    phase = {}
    launches = 0
    barrier()
    with torch.cuda.stream(stream):
        ev0.record(stream)
        for _ in range(args.steps):
            route_all()
            if dom is None:
                for k, v in r.timing().items():
                    phase[k] = phase.get(k, 0.0) + v
            launches += r.info(capi.INFO_LAUNCHES_LAST)
            if rm is not None:
                launches += rm.info(capi.INFO_LAUNCHES_LAST)
        if dom is not None:
            stream.wait_stream(stream_main)            # the last mainstem batch is inside the timed region
        ev1.record(stream)
        finish_all()
        if dom is not None:                            # phase times of the last batch stand for all (events are reused)
            for k, v in r.timing().items():
                phase[k] = v * args.steps
            if rm is not None:
                phase["mainstem"] = rm.timing()["total"] * args.steps
    barrier()
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if rank == 0 else None
    touched = r.info(capi.INFO_KWT_TOUCHED) / max(args.steps, 1) if has_kwt else 0
    if has_kwt:
        r.set_counting(False)
    particles = r.info(capi.INFO_KWT_PARTICLES) / net_local.nRch if has_kwt else None
    t_ms = torch.tensor([ms], dtype=torch.float64, device="cuda")
    # per rank: timed region [ms per step], routing time of its tributaries [ms per step], reaches, stages -- who is busy how long
    mine = torch.tensor([ms / args.steps, phase.get("total", 0.0) / args.steps, float(net_local.nRch), float(r.info(capi.INFO_NSTAGE))],
                        dtype=torch.float64, device="cuda")
    per_rank = [mine.clone() for _ in range(world)]
    if world > 1:
        dist.all_gather(per_rank, mine)
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    per_rank = [[round(float(v), 3) for v in t.tolist()] for t in per_rank]
    ms_max = float(t_ms.item())

    # ---- end to end through the public API: pinned host forcing -> H2D -> route -> D2H of REACH_Q series
    e2e = None
    if not args.no_e2e:
        nm = len(opts.route_opt)
        out_pin = torch.empty((nm, T, net_local.nRch), dtype=torch.float64).pin_memory()
        if rm is not None:
            rom_pin = torch.from_numpy(ro_main).pin_memory()
            outm_pin = torch.empty((nm, T, dom.main_net.nRch), dtype=torch.float64).pin_memory()

        def e2e_all():
            if dom is None:                                # software pipeline: H2D(b+1) | route(b) | D2H(b-1) overlap
                r.route_batch_async(ro_pin, out_pin)
                return
            dom.route_batch_pipelined(ro_pin, out_pin, rom_pin if rm is not None else None, outm_pin if rm is not None else None,
                                      stream, stream_main)

        with torch.cuda.stream(stream):
            e2e_all()                                      # warm the pinned path once
            if dom is None:
                r.wait()
            else:
                finish_all()
            barrier()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                e2e_all()
            if dom is None:
                r.wait()
            else:
                finish_all()
            torch.cuda.synchronize()
            e_s = time.perf_counter() - t0
        t_e = torch.tensor([e_s], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
        h2d = torch.tensor([float(ro.nbytes + (ro_main.nbytes if rm is not None else 0))], dtype=torch.float64, device="cuda")
        d2h = torch.tensor([float((out_pin.numel() + (outm_pin.numel() if rm is not None else 0)) * 8)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(h2d); dist.all_reduce(d2h)
        e2e = {"value": full_n * T * args.steps / float(t_e.item()), "unit": UNIT,
               "h2d_bytes_per_step": int(h2d.item()), "d2h_bytes_per_step": int(d2h.item()),
               "api": ("Router.route_batch_async (mr_step_batch_async): pinned host forcing in, REACH_Q series out, copies overlapped with routing"
                       if dom is None else "DomainSet.route_batch_pipelined: mr_step_batch_async per domain with pinned host buffers + NCCL hand-off")}
        del out_pin

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel family on rank 0's domain
    alg = algorithmic_bytes(r, net_local, opts, T, touched)
    fam_ms = {"k_basin": phase.get("basin", 0.0) / args.steps}
    if "1" in opts.route_opt:
        fam_ms["k_route<IRF>"] = phase.get("route_irf", 0.0) / args.steps
    if "2" in opts.route_opt:
        fam_ms["k_route_kwt"] = phase.get("route_kwt", 0.0) / args.steps
    if "0" in opts.route_opt:
        fam_ms["k_route<SUM>"] = phase.get("route_sum", 0.0) / args.steps
    dom = max(fam_ms, key=fam_ms.get)
    peak, peak_src = hbm_peak()
    n_launch_dom = 1 if dom == "k_basin" else (r.info(capi.INFO_NSTAGE) + T)      # wavefronts of one family per batch (a KWT wavefront is 1-4 launches)
    kernels = {k: {"ms_per_step": fam_ms[k], "alg_bytes_per_step": int(alg[k]),
                   "achieved_gbs": alg[k] / (fam_ms[k] * 1e-3) / 1e9 if fam_ms[k] > 0 else None} for k in fam_ms}
    if opts.doesBasinRoute == 1 and fam_ms["k_basin"] > 0:
        # SURVEY 8(d) counts the hillslope-UH window once per STEP (16*ntdh_bas B); k_basin walks it once per 64-step
        # chunk, so its real traffic is far below that figure and "achieved" can exceed the HBM peak.  The batch-amortised
        # minimum (window once per batch) and the DP work are reported beside it: the kernel is FP64-issue-bound.
        nb = r.info(capi.INFO_NTDH_BAS)
        amort = T * (20 * net_local.nHRU + 24 * net_local.nRch) + 16 * nb * net_local.nRch
        kernels["k_basin"].update({"bound": "fp64 (mul+add per UH ordinate, --fmad=false)", "amortised_min_bytes_per_step": int(amort),
                                   "achieved_gbs_amortised": amort / (fam_ms["k_basin"] * 1e-3) / 1e9,
                                   "dp_tflops": 2.0 * nb * net_local.nRch * T / (fam_ms["k_basin"] * 1e-3) / 1e12})
    # DRAM traffic per launch from the committed ncu capture: measured there per (reach, step) task of the family's launches
    # (profiles/traffic.json, <kernel>_dram_bytes_per_task) and scaled to this run's average launch
    per_task = ncu_traffic(args.workload, dom + "_dram_bytes_per_task")
    n_tasks_step = (net_local.nRch - r.info(capi.INFO_NHEAD)) * T
    traffic = per_task * n_tasks_step / n_launch_dom if per_task else ncu_traffic(args.workload, dom)
    roof = {"bound": "hbm", "kernel": dom, "achieved": kernels[dom]["achieved_gbs"], "peak": peak, "unit": "GB/s",
            "frac": (kernels[dom]["achieved_gbs"] or 0.0) / peak, "traffic": traffic,
            "peak_source": peak_src, "launches_per_step": n_launch_dom,
            "alg_bytes_per_launch": alg[dom] / n_launch_dom,
            "avg_launch_us": 1e3 * fam_ms[dom] / n_launch_dom,
            "share_of_step": fam_ms[dom] / max(ms / args.steps, 1e-12),
            "note": "routing methods run concurrently on separate streams, so the kernel families' times overlap; a KWT wavefront is "
                    "k_route_kwt_range (small) or k_route_kwt_light + k_route_kwt_heavy + k_route_kwt_team: launches_per_step counts wavefronts",
            "kernels_of_family": (["k_route_kwt_range", "k_route_kwt_light", "k_route_kwt_heavy", "k_route_kwt_team"] if dom == "k_route_kwt" else [dom])}

    # ---- CPU baseline: the oracle, seeded with the GPU's spun-up state, routes the next steps; the GPU routes
    #      the same steps, which doubles as a full-size parity sample
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        from oracle.oracle import Oracle, seed_oracle_from_router
        cores = host_cores()
        nS = max(1, min(T, args.ref_tsteps))
        o = Oracle(net_local, params, opts, n_threads=cores)
        seed_oracle_from_router(o, r)
        t0 = time.perf_counter()
        qo = o.run(ro[:nS])
        c_s = time.perf_counter() - t0
        qg = r.route_batch(np.ascontiguousarray(ro[:nS]))
        by_threads = {str(cores): net_local.nRch * nS / c_s}
        for nt in sorted({max(1, cores // 2), 1}, reverse=True):       # fewer threads, if that is faster on this host
            o.set_threads(nt)
            t0 = time.perf_counter()
            o.run(ro[nS:nS + 1] if T > nS else ro[:1], want_q=False)
            alt = (time.perf_counter() - t0) * nS
            by_threads[str(nt)] = net_local.nRch * nS / alt          # the restatement on 1 and on all host cores
            if alt < c_s:
                c_s, cores = alt, nt
        errs = {}
        for i, c in enumerate(opts.route_opt):
            errs[{"0": "SUM", "1": "IRF", "2": "KWT"}[c]] = float(np.max(np.abs(qg[i] - qo[i]) / np.maximum(np.abs(qo[i]), 1e-300)))
        cpu = {"value": net_local.nRch * nS / c_s, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{nS} routing time steps of the full {net_local.nRch}-reach network, continuing from the GPU's spun-up state",
               "note": "CPU restatement of the reference algorithm, OpenMP level sweep (reference Fortran not buildable in this image)",
               "by_threads": by_threads, "max_rel_err_gpu_vs_cpu": errs}

    # ---- the per-step seam (mr_step = the reference's `call main_route`, one step per call, host forcing in): not the metric,
    #      reported beside it (2448 dependent stages per method on C4: latency-bound)
    k1 = None
    if world == 1 and not args.no_e2e:
        for k in range(2):
            r.main_route(ro[k % T])
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for k in range(5):
            r.main_route(ro[(2 + k) % T])
        torch.cuda.synchronize()
        k1 = 1e3 * (time.perf_counter() - t0) / 5

    line = {
        "k1_ms_per_step": k1,
        "metric": METRIC, "value": full_n * T * args.steps / (ms_max * 1e-3), "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": args.workload, "nRch": full_n, "route_opt": opts.route_opt, "dt_qsim": opts.dt,
                   "timesteps_per_step": T, "nStage": r.info(capi.INFO_NSTAGE), "l2": "inputs larger than L2 (forcing + state per step >> 126 MB)",
                   "parallelism": ("single domain" if world == 1 else f"tributary domains bin-packed over {world} GPUs, {n_outlets} tributary outlets handed to "
                                   f"the {n_main}-reach mainstem on rank 0 by NCCL send/recv"),
                   "per_rank": {"ms_per_step": [p_[0] for p_ in per_rank], "tributary_ms_per_step": [p_[1] for p_ in per_rank],
                                "nRch": [int(p_[2]) for p_ in per_rank], "nStage": [int(p_[3]) for p_ in per_rank]},
                   "rank0_nRch": net_local.nRch, "mainstem_ms_per_step": (phase.get("mainstem", 0.0) / args.steps if world > 1 else None),
                   "kwt_particles_per_reach": particles},
        "roofline": roof, "kernels": kernels, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
