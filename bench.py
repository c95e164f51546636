#!/usr/bin/env python
"""bench.py -- NB IRLS gene-fits/sec on BASELINE.json config 2 (ruvIII.nb NB, k=5, 20k genes x 100k cells,
50 replicate groups), 1..8 B200, beside the CPU restatement of the reference.

A "step" is one inner IRLS iteration of ruvIII.nb (R/ruvIIInb.R:303-527): Huber certificate sweep, alpha
(get.coef2/get.adev), W (getW), gmean, Mb, candidate log-likelihood, accept.  gene-fits/sec = G x steps / time.

  value  counts resident in HBM, K calls of ruvnb_irls_iteration, CUDA events on the library's stream,
         max over ranks.
  e2e    the same metric through the public host API with HOST buffers: upload of the count matrix from
         pinned host memory, initial coefficients, log-likelihood, K iterations (each returns its
         log-likelihood to the host) and the read-back of W, alpha, gmean, Mb -- wall clock, max over ranks.
  N > 1  genes sharded over the ranks (strong scaling: the 20k x 100k problem is fixed), control-gene rows
         replicated, cells sharded for the W update, NCCL all-reduces for the exchange steps.

`--impl reference` times the reference's CPU algorithm (oracle/: C/OpenMP restatement when built, else the
NumPy restatement) on a bounded sample of the same workload, on rank 0 only.
"""
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

WORKLOADS = {
    # name: G, N, k, m, n_ctl
    "cfg2_nb_20kx100k": (20000, 100000, 5, 50, 1000),
    "small_2kx10k": (2000, 10000, 5, 50, 200),
    "prof_2kx100k": (2368, 100000, 5, 50, 256),  # ncu captures: same row length as cfg2, 16 rows per SM
    "tiny_256x2k": (256, 2048, 2, 3, 64),
}
SEED = 20261019 + 2


def fp64_instr_table():
    """FP64-pipe instructions per element of each sweep (ncu smsp__inst_executed_pipe_fp64 / elements of the captured launch);
    measured under ncu and committed as profiles/fp64_instr.json with the capture it came from -- nothing hard-coded here."""
    p = os.path.join(ROOT, "profiles", "fp64_instr.json")
    if not os.path.exists(p):
        return {}, None
    d = json.load(open(p))
    return d.get("instr_per_element", {}), d.get("source")


def host_threads():
    """Threads the CPU arm may use: the cores this process is allowed on (torchrun exports OMP_NUM_THREADS=1 for N > 1, which
    must not decide it)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


# ------------------------------------------------------------------------------------------------
# synthetic data, generated on the GPU by torch in gene chunks: plumbing only
# ------------------------------------------------------------------------------------------------
def truth_params(G, N, k, m, n_ctl, seed=SEED):
    from ruviiinb_b200 import synth
    rng = np.random.Generator(np.random.PCG64(seed))
    grp = synth.draw_groups(rng, N, m)
    W = rng.standard_normal((N, k))
    alpha = 0.3 * rng.standard_normal((G, k))
    gmean = np.clip(-0.5 + 1.5 * rng.standard_normal(G), -4.0, 5.0)
    Mb = 0.5 * rng.standard_normal((G, m))
    Mb[:n_ctl] = 0.0
    psi = np.exp(np.log(0.3) + 0.5 * rng.standard_normal(G))
    ctl = np.zeros(G, dtype=bool)
    ctl[:n_ctl] = True
    # bench start state: a perturbed truth (the fit itself starts from irlba + WLS; the throughput of an
    # iteration does not depend on where in the trajectory it is taken)
    W0 = W + 0.3 * rng.standard_normal((N, k))
    W0 /= np.sqrt((W0 ** 2).sum(axis=0))[None, :]
    return dict(grp=grp, W=W, alpha=alpha, gmean=gmean, Mb=Mb, psi=psi, ctl=ctl, W0=W0)


def gen_rows_gpu(tp, rows, out, device, chunk=500, seed=SEED):
    """Y[rows] ~ Poisson(Gamma(1/psi, mu psi)), log mu = gmean + Mb[,grp] + alpha W'; one torch generator
    seed per 500-gene chunk so the matrix does not depend on how genes are sharded."""
    import torch
    Wt = torch.as_tensor(tp["W"], device=device)
    grp = torch.as_tensor(tp["grp"].astype(np.int64), device=device)
    rows = np.asarray(rows)
    i = 0
    while i < len(rows):
        c = rows[i] // chunk
        j = i
        while j < len(rows) and rows[j] // chunk == c:
            j += 1
        g0, g1 = c * chunk, min((c + 1) * chunk, len(tp["gmean"]))
        a = torch.as_tensor(tp["alpha"][g0:g1], device=device)
        lmu = torch.as_tensor(tp["gmean"][g0:g1], device=device)[:, None] + torch.as_tensor(tp["Mb"][g0:g1], device=device)[:, grp] + a @ Wt.T
        psi = torch.as_tensor(tp["psi"][g0:g1], device=device)[:, None]
        shape = (1.0 / psi).expand_as(lmu).contiguous()
        torch.manual_seed(seed * 1000 + int(c))  # default CUDA generator: one seed per gene chunk
        lam = torch._standard_gamma(shape) * torch.exp(lmu) * psi
        y = torch.poisson(lam).to(torch.int32)
        sel = torch.as_tensor(rows[i:j] - g0, device=device)
        out[i:j] = y[sel].cpu().numpy()
        i = j
    return out


def gen_rows_cpu(tp, rows, seed=SEED, chunk=500):
    rows = np.asarray(rows)
    out = np.empty((len(rows), tp["W"].shape[0]), dtype=np.int32)
    for i, g in enumerate(rows):
        rng = np.random.Generator(np.random.PCG64(seed * 100003 + int(g)))
        lmu = tp["gmean"][g] + tp["Mb"][g][tp["grp"]] + tp["W"] @ tp["alpha"][g]
        lam = rng.gamma(1.0 / tp["psi"][g], np.exp(lmu) * tp["psi"][g])
        out[i] = rng.poisson(lam)
    return out


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
E2E_REPS = 3


class ClockSampler:
    """nvidia-smi polled every 20 ms in the background; only the samples taken between begin() and end() count.  start() returns
    once the first sample has arrived (nvidia-smi needs ~0.5 s to come up), so that even a sub-second region is sampled."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        self.index = index
        self.samples = []      # (perf_counter, line)
        self.proc = None
        self.t0 = self.t1 = None

    def start(self, wait=5.0):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
            t = time.perf_counter()
            while not self.samples and time.perf_counter() - t < wait:
                time.sleep(0.01)
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append((time.perf_counter(), line.strip()))

    def begin(self):
        self.t0 = time.perf_counter()

    def end(self):
        self.t1 = time.perf_counter()

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        if self.t1 is None:
            self.end()
        time.sleep(0.03)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        t0 = self.t0 if self.t0 is not None else -1e300
        inside = [l for (t, l) in self.samples if t0 <= t <= self.t1 + 0.03]
        note = None
        if not inside and self.samples:   # region shorter than one polling period: the sample nearest to it
            mid = 0.5 * (t0 + self.t1)
            inside = [min(self.samples, key=lambda x: abs(x[0] - mid))[1]]
            note = "region shorter than the polling period: nearest sample"
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in inside:
            f = [x.strip() for x in s.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        out = {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(max(mx)) if mx else None,
               "reasons": sorted(reasons), "samples": len(sm)}
        if note:
            out["note"] = note
        return out


# ------------------------------------------------------------------------------------------------
# CPU arm: the restated reference algorithm on a bounded sample
# ------------------------------------------------------------------------------------------------
def cpu_sample_run(workload, steps, warmup, target_seconds=20.0):
    """One inner IRLS iteration of the oracle on a bounded sample of the workload: all N cells, the control
    genes needed for W cut to n_ctl_s, G_s genes in total.  Returns (gene-fits/sec, description dict)."""
    G, N, k, m, n_ctl = WORKLOADS[workload]
    from oracle import cport
    tp = truth_params(G, N, k, m, n_ctl)
    cores = host_threads()
    have_c = cport.available()
    Gs = min(G, 2048, 64 * cores if have_c else 96)
    ncs = min(n_ctl, max(8, Gs // 4))
    rows = np.concatenate([np.arange(ncs), np.arange(n_ctl, n_ctl + Gs - ncs)])
    Y = gen_rows_cpu(tp, rows)
    ctl = np.zeros(Gs, dtype=bool)
    ctl[:ncs] = True
    state = dict(gmean=tp["gmean"][rows].copy(), Mb=tp["Mb"][rows].copy(), alpha=tp["alpha"][rows].copy(),
                 alpha_dev=tp["alpha"][rows] - tp["alpha"][rows].mean(axis=0), W=tp["W0"].copy(), psi=tp["psi"][rows].copy(),
                 step=np.ones(Gs))
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        state, ll = cport.inner_iteration(Y, state, tp["grp"], ctl, k_huber=100.0, lambda_a=0.01, lambda_b=16.0, nthreads=cores)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    tot = sum(times)
    kind = "port"
    sample = (f"{Gs} genes ({ncs} controls) x all {N} cells of {workload}, {steps} inner iterations, "
              f"{'C/OpenMP' if have_c else 'NumPy'} restatement of R/ruvIIInb.R:303-484")
    used = cport.threads() if have_c else 1    # omp_get_max_threads() after the call: what the C side really ran with
    return Gs * steps / tot, dict(kind=kind, cores=used, sample=sample), 1e3 * tot / steps


# ------------------------------------------------------------------------------------------------
def _claim_stdout():
    """Keep the real stdout for the ONE JSON line: native libraries (NCCL prints its version banner on stdout) and
    anything else writing to fd 1 go to stderr instead."""
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    return os.fdopen(saved, "w")


def main():
    json_out = _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    import bench_workloads as bw
    ap.add_argument("--workload", default="cfg2_nb_20kx100k", choices=sorted(WORKLOADS) + sorted(bw.CELL_WORKLOADS) + sorted(bw.FIT_WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-full-fit", action="store_true",
                    help="skip the WHOLE ruvIII.nb fit (ruvnb_fit_nb: Winsorize, init, dispersion estimation, alternating IRLS to "
                         "convergence) whose wall time is the second half of BASELINE.json's metric (\"full_fit\" in the line)")
    ap.add_argument("--full-fit", action="store_true", help="(default; kept for older command lines)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.workload in bw.FIT_WORKLOADS:     # BASELINE configs 3 and 1 on one GPU
        if args.impl == "reference":
            args.workload = "cfg2_nb_20kx100k"     # the CPU arm has the NB inner iteration only
        else:
            import torch
            torch.cuda.set_device(local)
            line = bw.run_fit(args, rank, world, local, None, json_out, ClockSampler)
            print(json.dumps(line), file=json_out, flush=True)
            return
    if args.workload in bw.CELL_WORKLOADS:    # BASELINE configs 4 and 5: the 10^6-cell passes, cells sharded over the ranks
        if args.impl == "reference":
            if rank == 0:
                print(json.dumps(bw.run_reference(args)), file=json_out, flush=True)
            return
        import torch
        import torch.distributed as dist
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device -- the B200 path has no CPU fallback")
        torch.cuda.set_device(local)
        if world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        fn = bw.run_fast if "fast" in args.workload or "cells_small" in args.workload and os.environ.get("BENCH_FAST") else bw.run_getres
        line = fn(args, rank, world, local, dist if world > 1 else None, json_out, ClockSampler)
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        if line is not None:
            print(json.dumps(line), file=json_out, flush=True)
        return
    G, N, k, m, n_ctl = WORKLOADS[args.workload]
    config = {"workload": f"ruvIII.nb NB inner IRLS iteration, k={k}, {G} genes x {N} cells, {m} replicate groups, "
                          f"{n_ctl} control genes ({args.workload})",
              "l2": "inputs larger than L2 (int32 counts %.1f GB per pass)" % (4.0 * G * N / 1e9),
              "sharding": "genes" if world > 1 else "none"}

    if args.impl == "reference":
        if rank != 0:
            return
        v, desc, ms = cpu_sample_run(args.workload, args.steps, max(1, min(args.warmup, 1)))
        line = {"impl": "reference", "metric": "NB IRLS gene-fits/sec", "value": v, "unit": "gene-fits/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                "cpu_baseline": dict(value=v, unit="gene-fits/s", **desc),
                "e2e": {"value": v, "unit": "gene-fits/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line), file=json_out, flush=True)
        return

    import torch
    import torch.distributed as dist
    from ruviiinb_b200 import _lib
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the B200 path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    tp = truth_params(G, N, k, m, n_ctl)
    # gene shard of this rank (strong scaling)
    per = (G + world - 1) // world
    g0, g1 = min(G, rank * per), min(G, (rank + 1) * per)
    Gl = g1 - g0
    t_gen = time.perf_counter()
    Yh_t = torch.empty((Gl, N), dtype=torch.int32, pin_memory=True)
    Yh = Yh_t.numpy()
    gen_rows_gpu(tp, np.arange(g0, g1), Yh, dev)
    # gene shards: the replicated control-gene slab is fetched from the owning ranks over NCCL by the library (Yctl = None);
    # RUVNB_BENCH_YCTL_HOST=1 uploads a host copy from every rank instead (the round-1 path)
    Yctl = None
    if world > 1 and os.environ.get("RUVNB_BENCH_YCTL_HOST"):
        Yctl_t = torch.empty((n_ctl, N), dtype=torch.int32, pin_memory=True)
        Yctl = Yctl_t.numpy()
        gen_rows_gpu(tp, np.arange(n_ctl), Yctl, dev)
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    t_gen = time.perf_counter() - t_gen

    opts = _lib.default_opts(k=k)
    opts.verbose = int(os.environ.get("RUVNB_VERBOSE", "0")) if rank == 0 else 0   # stage timings on stderr
    from ruviiinb_b200 import shard

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def make_ctx():
        ctx = _lib.Context(local)
        if world > 1:  # a fresh NCCL unique id per communicator (an id cannot be reused after its communicator is destroyed)
            ctx.comm_init(shard.broadcast_unique_id(dist, _lib.load(), device=dev), rank, world)
        ctx.set_design(G, N, tp["grp"], tp["ctl"], rows=(g0, g1))
        return ctx

    def set_start_state(ctx):
        ctx.set_state("W", tp["W0"])
        ctx.init_coef(opts)
        ctx.set_state("psi", tp["psi"][g0:g1])

    # ---- device-resident leg ------------------------------------------------------------------
    ctx = make_ctx()
    ctx.upload_counts(Yh, Yctl=Yctl)
    set_start_state(ctx)
    ctx.loglik(opts)
    for _ in range(args.warmup):
        ctx.irls_iteration(opts)
        ctx.accept()
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ctx.kernel_time_reset()
    l0 = ctx.kernel_launches
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    barrier()
    clocks.begin()
    e0.record(stream)
    logls, exact_rows_steps, exact_cells_steps = [], [], []
    ctx.cert_diag()
    for _ in range(args.steps):
        ll, bad = ctx.irls_iteration(opts)
        ctx.accept()
        logls.append(ll)
        exact_rows_steps.append(ctx.last_exact_rows)
        exact_cells_steps.append(ctx.last_exact_cells)
    e1.record(stream)
    barrier()
    clocks.end()
    clk = clocks.stop() if rank == 0 else None
    ms = e0.elapsed_time(e1)
    launches = ctx.kernel_launches - l0
    exact_rows = ctx.last_exact_rows
    ktimes = ctx.kernel_times()
    cert_diag = ctx.cert_diag()
    fp64_peak = ctx.measure_fp64() if rank == 0 else None
    tms = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms_max = float(tms.item())
    value = G * args.steps / (ms_max * 1e-3)
    ctx.close()

    # ---- end-to-end leg: host buffers through the public API ------------------------------------
    e2e = None
    if not args.no_e2e:
        # three repetitions, the MEDIAN one is reported (all three are listed): the region is a second of host-driven work on a shared
        # box, and a burst from another tenant (host memory bandwidth, driver calls) has been seen to stretch a single run several-fold
        runs = []
        for _rep in range(E2E_REPS):
            ctx = make_ctx()   # context, NCCL communicator and design: one-time set-up, outside the timed region
            barrier()
            t0 = time.perf_counter()
            ctx.upload_counts(Yh, Yctl=Yctl)
            t_e_up = time.perf_counter() - t0
            set_start_state(ctx)
            ctx.loglik(opts)
            t_e_ll = time.perf_counter() - t0 - t_e_up
            for _ in range(args.steps):
                ctx.irls_iteration(opts)
                ctx.accept()
            t_e_loop = time.perf_counter() - t0 - t_e_up - t_e_ll
            out = {n: ctx.get_state(n) for n in ("W", "alpha", "gmean", "Mb")}
            barrier()
            dt = time.perf_counter() - t0
            ctx.close()
            tdt = torch.tensor([dt], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(tdt, op=dist.ReduceOp.MAX)
            runs.append({"upload": t_e_up, "start_state_and_first_loglik": t_e_ll, "iterations": t_e_loop, "total": float(tdt.item())})
        med = sorted(runs, key=lambda r: r["total"])[len(runs) // 2]
        dt = med["total"]
        h2d = Yh.nbytes + (Yctl.nbytes if Yctl is not None else 0) + tp["W0"].nbytes + 8 * Gl + 4 * N + G
        d2h = sum(v.nbytes for v in out.values()) + 8 * args.steps
        e2e = {"value": G * args.steps / dt, "unit": "gene-fits/s", "h2d_bytes_per_step": int(h2d * world / args.steps),
               "d2h_bytes_per_step": int(d2h * world / args.steps),
               "seconds": med, "repetitions": [G * args.steps / r["total"] for r in runs],
               "what": "on an existing context: upload_counts (pinned host int32 -> HBM, permuted) + W0/psi upload + init_coef + loglik + K irls_iteration (each returns its log-likelihood) + get_state(W, alpha, gmean, Mb); wall clock, max over ranks; median of the repetitions listed"}

    # ---- optional: the whole fit, host buffers in, converged parameters out ---------------------------------------
    full_fit = None
    if not args.no_full_fit:
        # What the warm-up steps are to the iteration line: CUDA loads a kernel's code at its first launch, and Winsorize, init_W and
        # the dispersion step have not run in this process yet.  A small whole fit (same k, same options, one GPU, no communicator)
        # pays those one-time costs outside the timed region; its own wall time is reported beside the timed fit.
        t_w = time.perf_counter()
        if not os.environ.get("RUVNB_BENCH_NO_WARM_FIT"):
            Gw, Nw, mw, ncw = 1024, 4096, 8, 128
            tpw = truth_params(Gw, Nw, k, mw, ncw, seed=SEED + 99)
            Yw = np.empty((Gw, Nw), dtype=np.int32)
            gen_rows_gpu(tpw, np.arange(Gw), Yw, dev)
            with _lib.Context(local) as cw:
                cw.set_design(Gw, Nw, tpw["grp"], tpw["ctl"])
                cw.upload_counts(Yw)
                cw.winsorize()
                cw.fit_nb(opts, W0=cw.init_W(k)[0])
                cw.get_state("psi")
        t_warm_fit = time.perf_counter() - t_w
        ctx = make_ctx()
        fclocks = ClockSampler(local)
        if rank == 0:
            fclocks.start()
        barrier()
        fclocks.begin()
        t0 = time.perf_counter()
        ctx.upload_counts(Yh, Yctl=Yctl)
        t_up = time.perf_counter() - t0
        ctx.winsorize()
        t_win = time.perf_counter() - t0 - t_up
        W0c, sv, svd_iters = ctx.init_W(k)                      # H1 on the device: a COLD start, as ruvIII.nb does it
        t_init = time.perf_counter() - t0
        ctx.kernel_time_reset()
        logl_outer, recs, st = ctx.fit_nb(opts, W0=W0c)
        fit_kernels = {n: {"total_ms": round(tot, 3), "launches": cnt} for n, (tot, cnt) in ctx.kernel_times().items() if cnt}
        res = {n: ctx.get_state(n) for n in ("W", "alpha", "gmean", "Mb", "psi")}
        barrier()
        dt = time.perf_counter() - t0
        fclocks.end()
        fclk = fclocks.stop() if rank == 0 else None
        ctx.close()
        tdt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tdt, op=dist.ReduceOp.MAX)
        full_fit = {"wall_seconds": float(tdt.item()), "clocks": fclk, "outer_iterations": st["n_outer"], "inner_iterations": st["n_inner"],
                    "inner_loop_seconds": st["inner_seconds"], "dispersion_seconds": st["disp_seconds"],
                    "upload_winsorize_initW_seconds": t_init, "upload_seconds": t_up, "winsorize_seconds": t_win,
                    "initW_seconds": t_init - t_up - t_win, "initW_iterations": svd_iters,
                    "warmup_fit_seconds": t_warm_fit,
                    "gene_fits_per_s_inner_loops": G * st["n_inner"] / st["inner_seconds"], "logl_outer": [float(x) for x in logl_outer],
                    "exact_mad_rows": st.get("n_exact_mad_rows"), "inner_loop_brackets": fit_kernels,
                    "decisions": [[r["outer"], r["iter"], r["halving"], r["accepted"]] for r in recs],
                    "canonical_corr_W_truth": [float(x) for x in np.linalg.svd(np.linalg.qr(res["W"] - res["W"].mean(0))[0].T @ np.linalg.qr(tp["W"] - tp["W"].mean(0))[0], compute_uv=False)] if rank == 0 else None,
                    "what": "upload (pinned host int32) + Winsorize + ruvnb_init_W (truncated SVD of the control rows) + ruvnb_fit_nb with estimateDisp at every psi refresh + read-back; wall clock, max over ranks; preceded, untimed, by a 1024 x 4096 warm-up fit (warmup_fit_seconds) that triggers the first-launch kernel loads"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel ---------------------------------------------------------
    peaks = {}
    pk_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk_path):
        peaks = json.load(open(pk_path))
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    fused = not os.environ.get("RUVNB_NO_FUSE") and ktimes.get("gram", (0.0, 0))[1] > 0
    wcache = fused or not os.environ.get("RUVNB_NO_WCACHE")
    # algorithmic bytes per element of each sweep: counts are int32 (4 B); the fused sweep k_ll_w also WRITES
    # the two FP64 slabs (w, w Z: 16 B) that k_gram reads back (16 B) and k_pass2w reads half of (8 B) -- W / alpha / Mb are L2-resident
    per_elt = {"devstats": 4.0, "pass1": 4.0 + (8.0 if wcache else 0.0), "pass2": 8.0 if wcache else 4.0, "loglik": 20.0 if fused else 4.0,
               "gram": 16.0, "exact_mad": 0.0}
    sweep_bytes = {n: v * Gl * N for n, v in per_elt.items()}
    sweep_bytes["w_update"] = 4.0 * n_ctl * N / world
    for n in ktimes:                     # stream segments between the sweeps (small kernels + collectives): time only
        sweep_bytes.setdefault(n, 0.0)
    kern = {}
    for name, (tot, cnt) in ktimes.items():
        if cnt:
            kern[name] = {"ms": tot / cnt, "launches_per_step": cnt / args.steps,
                          "GBps": sweep_bytes[name] / (tot / cnt * 1e-3) / 1e9 if sweep_bytes[name] else None}
    dom = max((n for n in kern if sweep_bytes[n] > 0), key=lambda n: kern[n]["ms"] * kern[n]["launches_per_step"])
    label = {"loglik": "k_ll_w (log-likelihood + certificate + the next iteration's element-wise working quantities -> two slabs)" if fused else "k_loglik",
             "gram": "k_gram (Gram / score / group sums from the slabs)", "pass2": "k_pass2w (group sums of w W_new from the weight slab)" if wcache else "k_pass2",
             "pass1": "k_pass1", "w_update": "k_w_update_fast", "exact_mad": "k_signdev + k_row_sigma",
             "seg_alpha": "segment: alpha phase (normalise, all-reduce, solve, all-gather, clamp)", "seg_w": "segment: W exchange (all-reduce, column norms)",
             "seg_mb": "segment: gmean / Mb finish", "seg_tail": "segment: certificate finish, log-likelihood total, all-reduce"}
    key = {"loglik": "ll_w" if fused else "loglik", "pass2": "pass2w" if wcache else "pass2"}.get(dom, dom)
    # measured DRAM traffic of the dominant kernel (ncu --set full: dram__bytes_read.sum + dram__bytes_write.sum per element of the
    # captured launch, profiles/traffic.json), scaled to the elements this rank's launch walks
    traffic, traffic_src = None, None
    tr_path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tr_path):
        tr = json.load(open(tr_path))
        bpe = tr.get("bytes_per_element", {}).get(key)
        if bpe is not None:
            traffic = bpe * Gl * N
            traffic_src = tr.get("source")
    instr, instr_src = fp64_instr_table()
    ipe = instr.get(key)
    step_bytes = sum(sweep_bytes[n] * kern[n]["launches_per_step"] for n in kern if sweep_bytes[n])
    roofline = {"bound": "hbm", "kernel": label.get(dom, dom), "achieved": kern[dom]["GBps"], "peak": peak, "unit": "GB/s",
                "frac": kern[dom]["GBps"] / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": sweep_bytes[dom], "algorithmic_bytes_per_element": per_elt.get(dom),
                "note": "the dominant sweep is bound by the FP64 pipe, not by HBM (DESIGN.md section 5): see the fp64 block; k_gram and k_pass2w are the HBM-bound sweeps",
                "fp64": {"peak_tflops_measured": fp64_peak, "instr_per_element": ipe, "instr_source": instr_src,
                         "achieved_tflops": (2.0 * ipe * Gl * N / (kern[dom]["ms"] * 1e-3) / 1e12) if ipe else None,
                         "frac": (2.0 * ipe * Gl * N / (kern[dom]["ms"] * 1e-3) / 1e12 / fp64_peak) if ipe and fp64_peak else None,
                         "what": "FP64-pipe instructions per element counted from ncu (profiles/fp64_instr.json), 2 flops each, against the DFMA-chain peak measured in this run"},
                "whole_step": {"algorithmic_bytes": step_bytes, "GBps": step_bytes / (ms_max / args.steps * 1e-3) / 1e9,
                               "frac": step_bytes / (ms_max / args.steps * 1e-3) / 1e9 / peak},
                "kernels": {label.get(n, n): dict(v, frac_hbm=(v["GBps"] / peak if v["GBps"] else None)) for n, v in kern.items()}}

    line = {"metric": "NB IRLS gene-fits/sec", "value": value, "unit": "gene-fits/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config, "clocks": clk,
            "gpu_launches": int(launches), "exact_mad_rows_per_step": exact_rows_steps, "exact_mad_cells_per_step": exact_cells_steps, "cert_fail_reasons": cert_diag, "roofline": roofline, "logl_trace": logls, "datagen_seconds": t_gen}
    if e2e:
        line["e2e"] = e2e
    if full_fit:
        line["full_fit"] = full_fit
    if not args.no_cpu_baseline:   # rank 0, every N (the other ranks have left)
        v, desc, _ = cpu_sample_run(args.workload, 2, 1)
        line["cpu_baseline"] = dict(value=v, unit="gene-fits/s", **desc)
    print(json.dumps(line), file=json_out, flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
