#!/usr/bin/env python3
"""bench.py -- PTM logp+grad throughput on N B200s (obs-chain evaluations per second).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the host cores

One *step* = one fused logp+grad evaluation of all n observations for the rank's
batch of chains (BASELINE.json `metric`, configs[2]: n = 1M, 5+5 P-splines nbases 20,
1024 chains per GPU).  Chains shard across ranks with no collective ("weak": per-GPU
work is fixed, the job evaluates 1024*N chains).  Timing: CUDA events on the
launching stream around every step, L2 flushed (256 MiB write) between steps, a
barrier + synchronize on both sides of the K steps, max over ranks.
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "PTM logp+grad obs*chain evals/s"
UNIT = "obs*chain evals/s"
DEFAULT_CONFIG = "c3_n1M_5loc_5scale_nb20_1024chains"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default=DEFAULT_CONFIG)
    ap.add_argument("--dtype", default="f32", choices=["f64", "f32"],
                    help="arithmetic type of the sweep; f32 is the reference's default (LocScalePTM to_float32=True, "
                         "model/model.py:279-294); the other precision is measured in the same run under extra")
    ap.add_argument("--chains", type=int, default=0, help="chains per GPU (default: the config's)")
    ap.add_argument("--n", type=int, default=0, help="override the number of observations")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample-n", type=int, default=100_000)
    ap.add_argument("--no-extra", action="store_true", help="skip the secondary measurements under `extra`")
    return ap.parse_args()


# ------------------------------------------------------------------------------------
# CPU arm: the oracle (NumPy restatement of the reference algorithm) on the host cores
# ------------------------------------------------------------------------------------
_ORACLE = {}


def _oracle_worker(chains):
    om, st_all = _ORACLE["model"], _ORACLE["state"]
    from oracle import ptm_oracle as o

    sub = o.ChainState(
        beta=[b[chains] for b in st_all.beta], tau=st_all.tau[chains], delta_z=st_all.delta_z[chains],
        tau_delta=st_all.tau_delta[chains], beta0=st_all.beta0[chains], gamma0=st_all.gamma0[chains],
    )
    lp_, _ = om.logp_and_grad(sub)
    return float(lp_[0])


def cpu_oracle_throughput(config: str, dtype, sample_n: int, steps: int, warmup: int):
    """evals/s of the oracle on a bounded sample (first `sample_n` observations of the same
    seeded workload, one chain per worker per step), chains spread over all host cores."""
    import multiprocessing as mp

    try:
        from threadpoolctl import threadpool_limits
    except Exception:  # pragma: no cover
        threadpool_limits = None
    import bench_workload as bw
    from oracle import ptm_oracle as o

    cores = os.cpu_count() or 1
    workers = max(1, min(cores, 32))
    n_cfg = bw.CONFIGS[config][0]
    ns = min(sample_n, n_cfg)
    wl = bw.build(config, chains=workers, dtype=dtype, device="cpu", n_override=ns)
    terms = [
        o.term_from_reparam(wl.X[i], t.basis.knots, t.basis.Z, t.penalty, t.penalty_rank,
                            "loc" if k == 0 else "scale")
        for i, (t, k) in enumerate(wl.model.all_terms())
    ]
    om = o.PTMModel(wl.y, terms, nparam=wl.model.nparam, dtype=dtype)
    L, _ = wl.model.param_layout()
    p = wl.params
    T = len(terms)
    st = o.ChainState(
        beta=[p[:, L["beta"][t]: L["beta"][t] + terms[t].p] for t in range(T)],
        tau=p[:, L["tau"][0]: L["tau"][0] + T] if T else np.zeros((p.shape[0], 0)),
        delta_z=p[:, L["delta_z"]: L["delta_z"] + wl.model.nparam - 1],
        tau_delta=p[:, L["tau_delta"]], beta0=p[:, 0], gamma0=p[:, 1],
    )
    _ORACLE["model"], _ORACLE["state"] = om, st
    ctx = mp.get_context("fork")
    limiter = threadpool_limits(limits=1) if threadpool_limits else None
    times = []
    with ctx.Pool(workers) as pool:
        jobs = [[c] for c in range(workers)]
        for it in range(warmup + steps):
            t0 = time.perf_counter()
            pool.map(_oracle_worker, jobs)
            dt = time.perf_counter() - t0
            if it >= warmup:
                times.append(dt)
    if limiter is not None:
        limiter.restore_original_limits() if hasattr(limiter, "restore_original_limits") else None
    sec = float(np.mean(times))
    value = ns * workers / sec
    return dict(value=value, unit=UNIT, cores=workers, kind="port",
                sample=f"first {ns} obs of {config} x {workers} chains per step, NumPy f64 oracle "
                       f"(dense designs + grid-lerp + analytic grad), 1 chain per worker process, "
                       f"{steps} steps after {warmup} warm-up",
                ms_per_step=sec * 1e3, steps=steps)


def _config_dict(name, n, C, world, n_loc, n_scale, nbases, P):
    return {
        "workload": name, "n": n, "chains_per_gpu": C, "chains_total": C * world,
        "terms": f"{n_loc} loc + {n_scale} scale P-splines, nbases {nbases}, nparam 20",
        "params_per_chain": P, "parallelism": f"chain-sharded x{world}, no collective",
    }


def run_reference(args):
    """The reference's algorithm for the path on the host cores.  The reference itself cannot
    be installed here, so this is the oracle port; every step is a bounded
    sample of the workload (first --cpu-sample-n observations x one chain per core), the
    driver's --steps / --warmup are honoured as given."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    dtype = np.float64
    steps, warm = max(1, args.steps), max(0, args.warmup)
    cb = cpu_oracle_throughput(args.config, dtype, args.cpu_sample_n, steps, warm)
    import bench_workload as bw

    n, n_loc, n_scale, nbases, C = bw.CONFIGS[args.config]
    P = 2 + (n_loc + n_scale) * (nbases - 1) + 19 + (n_loc + n_scale) + 1
    cfg = _config_dict(args.config, n, args.chains or C, max(args.gpus, 1), n_loc, n_scale, nbases, P)
    cfg["reference_sample"] = cb["sample"]
    cfg["note"] = ("reference arm = CPU restatement (oracle) of the reference algorithm on the host cores; "
                   "the reference itself (jax 0.8.2 / liesel 0.4.3 / Python 3.13) cannot be installed in "
                   "this image; throughput is per obs*chain evaluation, so the bounded sample extrapolates "
                   "linearly to the configuration")
    line = {
        "impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT,
        "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": cb["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": cfg,
        "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.f.read().splitlines():
            parts = [x.strip() for x in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if sm:
            out["sm_mhz"] = float(np.median(sm))
            out["sm_max_mhz"] = float(max(mx))
            out["samples"] = len(sm)
        out["reasons"] = sorted(reasons)
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        return out


def run_ours(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dtype = np.float64 if args.dtype == "f64" else np.float32

    # CPU baseline first (forks worker processes before CUDA is initialised)
    cpu_baseline = None
    if rank == 0 and args.gpus == 1 and not args.no_cpu_baseline:
        cpu_baseline = cpu_oracle_throughput(args.config, np.float64, args.cpu_sample_n, 3, 1)
        cpu_baseline = {k: cpu_baseline[k] for k in ("value", "unit", "cores", "kind", "sample")}

    import torch
    import torch.distributed as dist

    import bench_workload as bw

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    wl = bw.build(args.config, chains=args.chains or None, dtype=dtype, device=dev, seed=0,
                  n_override=args.n or None)
    # same data on every rank, independent chain states per rank (chain-axis sharding)
    if rank > 0:
        wl.params = bw.make_states(wl.model, wl.params.shape[0], seed=1 + 1000 * rank).astype(dtype)
    model = wl.model.build()
    C = wl.params.shape[0]
    n = wl.n
    tdt = torch.float64 if dtype == np.float64 else torch.float32
    params = torch.from_numpy(wl.params).to(dev)
    logp = torch.empty(C, dtype=tdt, device=dev)
    grad = torch.empty_like(params)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        model.log_prob_and_grad(params, logp, grad)

    for _ in range(max(args.warmup, 3)):
        flush.fill_(1)
        step()
    barrier()
    K = args.steps
    ev0 = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    ev1 = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    sampler = ClockSampler(local_rank) if rank == 0 else None
    barrier()
    for k in range(K):
        flush.fill_(k & 0xFF)  # L2 flush: 256 MiB write, outside the per-step events
        ev0[k].record()
        step()
        ev1[k].record()
    barrier()
    clocks = sampler.stop() if sampler else None
    total_ms = sum(a.elapsed_time(b) for a, b in zip(ev0, ev1))
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    launches_per_step = model.last_launch_count()
    route = model.last_main_route()

    # dominant kernel (k_main) duration, CUDA events on the launching stream
    model.profile(True)
    km = []
    for _ in range(min(K, 10)):
        flush.fill_(3)
        step()
        km.append(model.last_kernel_ms())
    model.profile(False)
    pre_ms, main_ms, post_ms = (float(np.mean([x[i] for x in km])) for i in range(3))

    # end to end through the public call with HOST buffers (pinned H2D of the chain states,
    # kernels, D2H of logp + grad), wall clock around a synchronised step
    host_params = wl.params.copy()
    for _ in range(2):
        model.log_prob_and_grad_host(host_params)
    barrier()
    e2e_steps = min(K, 10)
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        lp_h, g_h = model.log_prob_and_grad_host(host_params)
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps
    t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t.item())
    finite = bool(np.isfinite(lp_h).all() and np.isfinite(g_h).all())

    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak = float(json.load(open(peaks_path))["hbm_gbs"])
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"

    # secondary measurements of the same run (bench_extra.py); the headline numbers above are
    # complete before any of this starts
    extra = {}
    if not args.no_extra and args.config == DEFAULT_CONFIG and not args.n and not args.chains:
        import bench_extra as bx
        import liesel_ptm_b200 as lp_mod

        del model, params, grad, logp
        wl.model = None
        torch.cuda.empty_cache()
        reps = max(3, min(K, 10))
        if world == 1:
            other = np.float64 if args.dtype == "f32" else np.float32
            bx._guard(extra, "c3_f64" if args.dtype == "f32" else "c3_f32",
                      lambda: bx.c3_other_dtype(torch, bw, dev, peak, reps, other))
            bx._guard(extra, "c4_information", lambda: bx.c4_information(torch, bw, dev, 5))
            bx._guard(extra, "ess", lambda: bx.ess(torch, bw, dev))
        else:
            bx._guard(extra, "c3_strong_scaling", lambda: bx.strong_scaling_c3(torch, dist, bw, dev, rank, world, dtype, reps))
            bx._guard(extra, "c5_obs_sharded", lambda: bx.c5_obs_sharded(torch, dist, bw, lp_mod, dev, rank, world, dtype, 5))

    if rank == 0:
        ms_per_step = total_ms / K
        value = n * C * world / (ms_per_step * 1e-3)
        achieved = n * C * wl.bytes_alg / (main_ms * 1e-3) / 1e9
        traffic, traffic_src = None, None
        tp = os.path.join(ROOT, "profiles", "r2_traffic.json")
        if os.path.exists(tp):
            try:
                rec = json.load(open(tp))[f"{args.dtype}_{route}"]
                traffic, traffic_src = rec["dram_bytes_per_step"], rec["source"]
            except Exception:
                traffic = None
        itemsize = wl.params.dtype.itemsize
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": dict(
                _config_dict(wl.name, n, C, world, wl.n_loc, wl.n_scale, wl.nbases, int(wl.params.shape[1])),
                l2="flushed between timed steps (256 MiB write); inputs are 88 MB < L2",
                finite_outputs=finite),
            "roofline": {
                "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic,
                "traffic_source": f"static, not re-measured in this run: {traffic_src}",
                "peak_source": peak_src,
                "kernel": "ptm::k_main_tc + ptm::k_grad_tc (tcgen05 route)" if route == "tcgen05" else "ptm::k_main",
                "route": route,
                "kernel_ms": main_ms, "pre_ms": pre_ms, "post_ms": post_ms,
                "bytes_alg_per_eval": wl.bytes_alg,
                "note": "algorithmic bytes = w*(1+T_x) per obs-chain eval; kernel_ms = CUDA events around the sweep "
                        "kernel(s) of one step on the launching stream.  The observations are re-used across the "
                        "chains of a CTA (32 in k_main, 128 in the tcgen05 route); the tcgen05 route moves "
                        "dl/deta (8 B per eval) through HBM between its two kernels (see traffic)",
            },
            "e2e": {
                "value": n * C * world / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": int(C * wl.params.shape[1] * itemsize),
                "d2h_bytes_per_step": int(C * (wl.params.shape[1] + 1) * itemsize),
                "note": "chain states from pinned host memory each step; observations are resident "
                        "(copied once at build, as in MCMC)",
            },
            "gpu_launches": int(launches_per_step * K),
            "clocks": clocks,
        }
        if cpu_baseline is not None:
            line["cpu_baseline"] = cpu_baseline
        if extra:
            line["extra"] = extra
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
