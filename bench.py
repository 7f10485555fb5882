#!/usr/bin/env python
"""bench.py -- Koopman samples/s per ISOKANN iteration (BASELINE.json metric).

One "step" is one body of the reference's run! loop (src/iso.jl:72-94): Koopman target
(featurize -> chi over all K*N samples -> K-mean -> isotarget) followed by one training epoch
(minibatched fwd/bwd + optimiser steps over the N start points).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config c5] [--impl reference]

N > 1 is launched by torchrun, one rank per GPU; start points are sharded over ranks
(strong scaling: the workload is fixed).  Prints ONE JSON line on rank 0.  The headline (`value`, `e2e`,
`roofline`) is BASELINE config 5; `configs` carries configs 2, 3 and 4 (ISA and PseudoInv) measured the same way.
"""
from __future__ import annotations

import os
import sys


def _affinity_threads() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


if "--impl" in sys.argv and "reference" in sys.argv:
    # the CPU arm uses every host core it may run on, whatever the launcher exported (torchrun sets
    # OMP_NUM_THREADS=1): BLAS thread pools read these variables when numpy is imported
    for _v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[_v] = str(_affinity_threads())

import argparse
import json
import subprocess
import threading
import time
import zlib
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "koopman_samples_per_s_per_isokann_iteration"
UNIT = "samples/s"
def dtype_note():
    fwd = os.environ.get("ISOKANN_TC_FWD", "bf16x3")
    return ("fp32 storage and accumulation; wide Dense layers multiply split 16-bit operands on tcgen05: training "
            "steps bf16 pairs x 3 MMAs (hi*hi + hi*lo + lo*hi), inference forward (Koopman pass, chis) "
            + ("fp16 pairs x 2 MMAs (hi*hi + lo*hi, weights rounded once to fp16)" if fwd == "fp16x2"
               else "bf16 pairs x 3 MMAs") + "; narrow layers run FP32 FFMA")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", default="c5", choices=["c1", "c2", "c3", "c4", "c5"])
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--N", type=int, default=None)
    ap.add_argument("--K", type=int, default=None)
    ap.add_argument("--minibatch", type=int, default=None)
    ap.add_argument("--target", default=None, choices=["shiftscale", "isa", "pinv"])
    ap.add_argument("--gemm", default="auto", choices=["auto", "fp32", "tc"])
    ap.add_argument("--tc-fwd", default=None, choices=["bf16x3", "fp16x2"],
                    help="operand format of the inference forward of the wide layers (default: the library's)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the c2/c3/c4 entries of `configs`")
    ap.add_argument("--profile", action="store_true",
                    help="short run for ncu: no e2e / CPU baseline / extra configs, warm-up not forced to 3; "
                         "never a bench value")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample", type=int, default=0, help="start points in the CPU baseline sample")
    return ap.parse_args()


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return {"hbm_gbs": d["hbm_gbs"], "tflops": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "src": "measured"}
    return {"hbm_gbs": 6650.0, "tflops": 1400.0, "src": "fallback"}


def workload_config(w, N, K, B, target, extra=None):
    cfg = {"workload": f"{w.name}: {w.n_atoms}-atom pairdist featurizer F={w.F}, pairnet {w.widths}, N={N}, K={K}, "
                       f"{target} target, {w.opt}, minibatch={B}",
           "N": N, "K": K, "minibatch": B, "widths": list(w.widths), "target": target, "optimiser": w.opt,
           "arithmetic": dtype_note(),
           "l2": "inputs (coords of K*N samples) larger than the 126 MB L2" if N * K * w.D * 4 > 126e6
                 else "inputs fit in L2; steady-state iteration"}
    if extra:
        cfg.update(extra)
    return cfg


# ---------------------------------------------------------------------------------------------
# CPU baseline: the oracle port of the reference's run! on the host cores, bounded sample
# ---------------------------------------------------------------------------------------------
def blas_threads():
    try:
        from threadpoolctl import threadpool_info
        return max([int(p.get("num_threads", 1)) for p in threadpool_info()] or [1])
    except Exception:
        return None


def cpu_iteration_sample(pkg, w, Ns, K, B, steps, warmup, target):
    """time `steps` oracle iterations on Ns start points (cached Float32 features exactly as the
    reference does, src/simulation.jl:112); returns (samples/s, seconds per step)"""
    import oracle
    xs, ys = pkg.synthetic.make_data(w, Ns, K)
    rec = lambda a: np.ascontiguousarray(np.asarray(a).T)
    if w.featurizer == "identity":
        xsf, ysf = rec(xs).astype(np.float32), rec(ys).astype(np.float32)
    else:
        xsf, ysf = oracle.flatpairdists(rec(xs)), oracle.flatpairdists(rec(ys))
    m = oracle.init_params(oracle.Model(list(w.widths), w.layernorm), np.random.default_rng(w.seed + 1))
    cfg = oracle.OptConfig(kind=w.opt)
    st = oracle.opt_init(cfg, oracle.num_params(m))
    perms = pkg.synthetic.make_perms(w, Ns, warmup + steps)
    for i in range(warmup):
        oracle.run(m, xsf, ysf, cfg, st, 1, B, [perms[i]], target)
    t0 = time.perf_counter()
    for i in range(steps):
        oracle.run(m, xsf, ysf, cfg, st, 1, B, [perms[warmup + i]], target)
    dt = (time.perf_counter() - t0) / steps
    return Ns * K / dt, dt


def cpu_sample_size(w, N, K, requested):
    """start points of the CPU sample: all of N when one iteration is below ~2e12 flop, else N/64"""
    if requested > 0:
        return min(N, requested)
    flop_per_start = 2.0 * w.macs() * (K + 3)
    if flop_per_start * N <= 2.0e12:
        return N
    return max(256, N // 64)


def run_reference(args, pkg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = pkg.synthetic.WORKLOADS[args.config]
    N = args.N or w.N
    K = args.K or w.K
    B = args.minibatch if args.minibatch is not None else w.minibatch
    target = args.target or w.target
    Ns = cpu_sample_size(w, N, K, args.cpu_sample)
    steps, warm = max(1, args.steps), max(0, args.warmup)
    if args.cpu_sample <= 0 and Ns >= 4096:
        # keep the whole --steps/--warmup run within ~4 minutes: calibrate on a 1/8 sample of the sample and halve
        # Ns (1/64 -> 1/128 -> ...) while the projected run time exceeds the budget
        n0 = max(256, Ns // 8)
        _, dt0 = cpu_iteration_sample(pkg, w, n0, K, min(B, n0) if B else 0, 1, 1, target)
        while Ns > 1024 and dt0 * (Ns / n0) * (steps + warm) > 240.0:
            Ns //= 2
    Bs = min(B, Ns) if B else 0
    val, dt = cpu_iteration_sample(pkg, w, Ns, K, Bs, steps, warm, target)
    threads = blas_threads()
    sample = (f"{Ns} of {N} start points (x{K} Koopman samples) = 1/{N / Ns:.0f} of the workload, one full iteration "
              f"each step, minibatch min(B, Ns) = {Bs}; per-sample throughput, i.e. scaled linearly to the full N")
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warm, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": workload_config(w, N, K, B, target,
                                  {"note": "CPU restatement of reference run! (Flux semantics) in numpy/OpenBLAS: "
                                           "multi-threaded BLAS, single-threaded broadcasts, like Flux on the CPU; "
                                           "Julia is not installable here, so the reference itself cannot be timed",
                                   "arithmetic": "fp32 numpy (OpenBLAS sgemm)"}),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": _affinity_threads(), "threads": threads, "kind": "port",
                         "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "power_w_max": float(max(pw)), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------------------------
BLOCK_PTS = 4096   # start points per noise block of the synthetic ys


def device_data(pkg, w, N, K, off, n_loc, device):
    """synthetic coordinates generated on the device: xs (N, D) on every rank, ys (n_loc, K, D) shard.  The noise of
    ys is drawn per block of 4096 start points from a seed that depends on the block only, so every world size
    sees the same data set and the `check` blocks of the 1/2/4/8-GPU lines are comparable."""
    import torch
    rng = np.random.default_rng(w.seed)
    if w.featurizer == "identity":
        xs_h, ys_h = pkg.synthetic.make_data(w, N, K)
        xs = torch.from_numpy(np.ascontiguousarray(xs_h.T)).to(device)
        ys = torch.from_numpy(np.ascontiguousarray(ys_h.T[off:off + n_loc])).to(device)
        return xs, ys
    states = pkg.synthetic.adp_states(w.states) if w.n_atoms == 22 else pkg.synthetic.villin_states(rng, w.n_atoms)
    base = torch.tensor(np.stack([s.reshape(-1) for s in states]), dtype=torch.float32, device=device)
    g = torch.Generator(device=device)
    g.manual_seed(1234 + w.seed)
    which = torch.randint(0, len(states), (N,), generator=g, device=device)
    xs = base[which] + 0.05 * torch.randn((N, w.D), generator=g, device=device)
    ys = torch.empty((n_loc, K, w.D), dtype=torch.float32, device=device)
    g2 = torch.Generator(device=device)
    for b in range(off // BLOCK_PTS, (off + n_loc + BLOCK_PTS - 1) // BLOCK_PTS if n_loc > 0 else 0):
        g2.manual_seed(1_000_003 * (w.seed + 1) + b)
        noise = 0.03 * torch.randn((BLOCK_PTS, K, w.D), generator=g2, device=device)
        s, e = max(off, b * BLOCK_PTS), min(off + n_loc, (b + 1) * BLOCK_PTS, N)
        ys[s - off:e - off] = xs[s:e, None, :] + noise[s - b * BLOCK_PTS:e - b * BLOCK_PTS]
    return xs.contiguous(), ys


class Runner:
    """one workload on this rank's GPU: engine + resident synthetic data + timing helpers"""

    def __init__(self, pkg, args, w, N, K, B, target, world, rank, local, uid):
        import torch
        self.torch = torch
        self.pkg, self.w, self.N, self.K, self.B, self.target = pkg, w, N, K, B, target
        self.world, self.rank, self.local = world, rank, local
        self.device = torch.device("cuda", local)
        self.off, self.n_loc = pkg.parallel.shard_range(N, world, rank)
        rngp = np.random.default_rng(w.seed + 1)
        model = pkg.densenet(w.widths, layernorm=w.layernorm, rng=rngp)
        rule = pkg.AdamRegularized() if w.opt == "adam" else pkg.NesterovRegularized()
        self.eng = pkg.Engine(model, rule, "allpairs" if w.featurizer == "allpairs" else "identity", w.n_atoms, None,
                              device=local, gemm=args.gemm)
        if world > 1:
            self.eng.comm_init(world, rank, uid)
        self.xs, self.ys = device_data(pkg, w, N, K, self.off, self.n_loc, self.device)
        torch.cuda.synchronize()
        self.eng.set_data_dev(self.xs, self.ys, w.D, K, N, self.off, self.n_loc)
        self.ext = torch.cuda.ExternalStream(self.eng.stream(), device=self.device)

    def barrier(self):
        import torch.distributed as dist
        self.eng.synchronize()
        self.torch.cuda.synchronize()
        if self.world > 1:
            dist.barrier()
            self.torch.cuda.synchronize()

    def timed(self, fn, steps):
        """device time of `steps` calls of fn, bracketed by barrier+sync, CUDA events on the library's stream,
        max over ranks"""
        import torch.distributed as dist
        torch = self.torch
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.barrier()
        with torch.cuda.stream(self.ext):
            e0.record()
        for i in range(steps):
            fn(i)
        with torch.cuda.stream(self.ext):
            e1.record()
        self.barrier()
        ms = e0.elapsed_time(e1)
        if self.world > 1:
            t = torch.tensor([ms], device=self.device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    def iterate(self, perm):
        return self.eng.iterate(self.target, 1, 1, self.B, perm)

    def check(self, losses):
        """numbers that let the 1/2/4/8-GPU lines be compared: the data set, the initial weights and the
        permutations are identical for every world size, so these agree up to fp32 summation order"""
        import torch.distributed as dist
        flat = self.eng.download_params()
        k = self.eng.koopman()
        crc = zlib.crc32(flat.tobytes())
        same = True
        if self.world > 1:
            t = self.torch.tensor([crc], dtype=self.torch.int64, device=self.device)
            all_t = [self.torch.zeros_like(t) for _ in range(self.world)]
            dist.all_gather(all_t, t)
            same = all(int(x.item()) == crc for x in all_t)
        return {"first_loss": float(losses[0]), "last_loss": float(losses[-1]),
                "param_l2": float(np.linalg.norm(flat.astype(np.float64))),
                "param_crc32": int(crc), "ranks_hold_identical_params": bool(same),
                "kchi_min": float(k.min()), "kchi_max": float(k.max()), "kchi_mean": float(k.astype(np.float64).mean()),
                "after": f"{len(losses)} warm-up+timed iterations from the seeded initial weights"}

    def close(self):
        self.eng.close()
        del self.xs, self.ys
        self.torch.cuda.empty_cache()


def measure(r: Runner, nsteps, nwarm, perms, with_clocks):
    """resident-data throughput, then the same steps again with CUDA events around every kernel launch"""
    losses = []
    for i in range(nwarm):
        losses.extend(r.iterate(perms[i]))
    r.eng.reset_stats()
    r.eng.enable_timing(2)          # one event pair per phase: the timed pass itself yields the phase split
    sampler = ClockSampler(r.local) if with_clocks and r.rank == 0 else None
    if sampler:
        sampler.start()
    ms = r.timed(lambda i: losses.extend(r.iterate(perms[nwarm + i])), nsteps)
    clocks = sampler.stop() if sampler else None
    st0 = r.eng.stats()
    r.eng.enable_timing(False)
    launches = st0["kernel_launches"]
    chk = r.check(losses)
    r.eng.reset_stats()
    r.eng.enable_timing(True)
    r.timed(lambda i: r.iterate(perms[nwarm + i]), nsteps)
    st = r.eng.stats()
    r.eng.enable_timing(False)
    # phases from the timed pass (graph replay, no per-kernel events); kernel classes from the instrumented pass
    for k in ("ms_koopman_total", "ms_target_total", "ms_train_total"):
        st[k + "_instrumented"] = st[k]
        st[k] = st0[k]
    st["graph_launches"] = st0.get("graph_launches", 0)
    return ms, clocks, launches, chk, st


def phase_dict(st, nsteps):
    return ({"koopman": st["ms_koopman_total"] / nsteps, "target": st["ms_target_total"] / nsteps,
             "train": st["ms_train_total"] / nsteps, "train_instrumented": st["ms_train_total_instrumented"] / nsteps,
             "epochs_replayed_as_graph": st["graph_launches"]},
            {"featurize": st["ms_featurize"] / nsteps, "gemm": st["ms_gemm"] / nsteps,
             "reduce": st["ms_reduce"] / nsteps, "train_elementwise": st["ms_train_elementwise"] / nsteps,
             "optimiser": st["ms_optimiser"] / nsteps, "nccl": st["ms_nccl"] / nsteps})


def hbm_roofline(w, N, K, ms_iter, pk, nd):
    """whole-iteration HBM roofline of the narrow-net configs: BASELINE.md section 4 minimum bytes
    4*D*N*(K+1) (+ 4*D*N for the extra chi(xs) pass of the N-D targets), coordinates read once per pass"""
    byts = 4.0 * w.D * N * (K + 1 + (1 if nd else 0))
    ach = byts / (ms_iter * 1e-3) / 1e9
    return {"bound": "hbm", "kernel": "whole iteration (featurize+MLP fused passes over the coordinates)",
            "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": ach / pk["hbm_gbs"], "traffic": None,
            "algorithmic_bytes": byts, "peak_source": pk["src"]}


EXTRA = [("c2", "c2", None), ("c3", "c3", None), ("c4_isa", "c4", "isa"), ("c4_pinv", "c4", "pinv")]


def run_b200(args, pkg):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the ISOKANN hot path has no CPU fallback "
                         "(use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)
    if args.tc_fwd:
        os.environ["ISOKANN_TC_FWD"] = args.tc_fwd
    w = pkg.synthetic.WORKLOADS[args.config]
    N = args.N or w.N
    K = args.K or w.K
    B = args.minibatch if args.minibatch is not None else w.minibatch
    target = args.target or w.target
    nsteps, nwarm = args.steps, (args.warmup if args.profile else max(3, args.warmup))
    if args.profile:
        args.no_e2e = args.no_cpu_baseline = args.no_extra = True
    pk = peaks()

    def new_uid():
        return pkg.parallel.broadcast_unique_id(rank) if world > 1 else None

    r = Runner(pkg, args, w, N, K, B, target, world, rank, local, new_uid())
    eng, xs, ys, off, n_loc = r.eng, r.xs, r.ys, r.off, r.n_loc
    perms = pkg.synthetic.make_perms(w, N, nwarm + nsteps)
    ms, clocks, launches, chk, st = measure(r, nsteps, nwarm, perms, True)
    value = N * K * nsteps / (ms * 1e-3)

    # ---- end to end through the public API with HOST buffers ----
    e2e = None
    if not args.no_e2e:
        # plain (pageable) numpy arrays, as a Julia caller would hand over: the library page-locks them itself
        xs_j = np.asfortranarray(xs.cpu().numpy().T)          # Julia-shaped (D, N)
        ys_j = np.asfortranarray(ys.cpu().numpy().T)          # (D, K, n_loc)
        nparams = eng.P

        def e2e_step(i):
            # SimulationData upload + run!(iso, 1) + the loss and cpu(iso) (the updated model) back on the host
            eng.set_data_async(xs_j, ys_j, n_offset=off, n_local=n_loc)
            eng.iterate(target, 1, 1, B, perms[i % len(perms)])
            eng.download_params()
        for i in range(2):
            e2e_step(i)
        ms_e = r.timed(e2e_step, nsteps)
        # whole job, all ranks together: every rank uploads its shard of ys, its own rows of xs (the other rows come
        # from the peers over NVLink) and the permutation; every rank reads the loss and the parameters back
        h2d = 4 * (xs.numel() + N * K * xs.shape[1]) + 8 * N * world
        d2h = (8 + 4 * nparams) * world
        e2e = {"value": N * K * nsteps / (ms_e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "ms_per_step": ms_e / nsteps,
               "call": "isokann_set_data_async (pageable host xs/ys, page-locked by the library, ys streamed in "
                       "behind the Koopman pass) + run!(iso,1) + loss and cpu(iso) parameters to host"}
        eng.release_host_buffers()
        del xs_j, ys_j
    r.close()

    # ---- the other BASELINE configs, measured the same way (device-timed, resident data) ----
    extra = {}
    if not args.no_extra and args.config == "c5" and args.N is None:
        for key, cname, tgt in EXTRA:
            we = pkg.synthetic.WORKLOADS[cname]
            te = tgt or we.target
            re_ = Runner(pkg, args, we, we.N, we.K, we.minibatch, te, world, rank, local, new_uid())
            pe = pkg.synthetic.make_perms(we, we.N, 3 + 5)
            ms_x, _, launches_x, chk_x, st_x = measure(re_, 5, 3, pe, False)
            ph, kk = phase_dict(st_x, 5)
            extra[key] = {"config": workload_config(we, we.N, we.K, we.minibatch, te),
                          "ms_per_iteration": ms_x / 5, "value": we.N * we.K * 5 / (ms_x * 1e-3), "unit": UNIT,
                          "steps": 5, "warmup": 3, "gpu_launches": int(launches_x),
                          "roofline": hbm_roofline(we, we.N, we.K, ms_x / 5, pk, te != "shiftscale"),
                          "phase_ms_per_step": ph, "kernel_ms_per_step": kk, "check": chk_x}
            re_.close()

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    traffic = None
    tf = ROOT / "profiles" / "traffic.json"
    if tf.exists() and N == w.N and K == w.K:
        traffic = json.loads(tf.read_text()).get(args.config, {}).get("traffic_bytes_per_launch")
    tensor_bound = max(w.widths[1:-1] or [0]) >= 256
    if tensor_bound:
        ach = st["gemm_flops"] / (st["ms_gemm"] * 1e-3) / 1e12 if st["ms_gemm"] > 0 else 0.0
        roof = {"bound": "tensor", "kernel": "dense-layer GEMM (fused bias+activation)", "achieved": ach,
                "peak": pk["tflops"], "unit": "TFLOP/s", "frac": ach / pk["tflops"], "traffic": traffic,
                "executed_tflops": (st["gemm_mma_flops"] / (st["ms_gemm"] * 1e-3) / 1e12) if st["ms_gemm"] > 0 else 0.0,
                "mmas_per_product": st["gemm_mma_flops"] / st["gemm_flops"] if st["gemm_flops"] > 0 else None,
                "peak_source": f"{pk['src']} bf16 sustained", "launches": st["n_gemm_launches"],
                "avg_launch_ms": st["ms_gemm"] / max(1, st["n_gemm_launches"]),
                "note": "algorithmic 2*M*N*K flops of all GEMM launches / their CUDA-event time; every product costs "
                        "mmas_per_product 16-bit MMAs (3 = hi*hi + hi*lo + lo*hi on bf16 pairs; 2 = hi*hi + lo*hi on "
                        "fp16 pairs with the weights rounded once, inference forward only) to stay inside the fp32 "
                        "tolerance, so frac <= 1/mmas_per_product by construction and executed_tflops is what the "
                        "tensor pipe runs; traffic = mean dram bytes per launch from profiles/traffic.json (ncu)"}
    else:
        roof = hbm_roofline(w, N, K, ms / nsteps, pk, target != "shiftscale")
    # the other kernel north_star asks a roofline figure for: the featurizer against measured HBM bandwidth
    fz = None
    if st["ms_featurize"] > 0 and st["featurize_bytes"] > 0:
        fa = st["featurize_bytes"] / (st["ms_featurize"] * 1e-3) / 1e9
        fz = {"bound": "hbm", "achieved": fa, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": fa / pk["hbm_gbs"],
              "launches": st["n_featurize_launches"],
              "note": "algorithmic 4*(D+F) bytes per record / CUDA-event time of the featurizer launches, timed inside "
                      "the step (power-capped SM clock); split bf16 output moves 4*(D+ld) bytes, ld = F padded to 64"}
    ph, kk = phase_dict(st, nsteps)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": nsteps, "warmup": nwarm,
        "ms_per_step": ms / nsteps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": workload_config(w, N, K, B, target, {"parallelism": f"start points sharded over {world} GPU(s)",
                                                       "gemm": args.gemm}),
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
        "roofline": roof, "roofline_featurizer": fz, "check": chk,
        "phase_ms_per_step": ph, "kernel_ms_per_step": kk, "configs": extra or None,
    }
    if not args.no_cpu_baseline and world == 1:
        Ns = cpu_sample_size(w, N, K, args.cpu_sample)
        Bs = min(B, Ns) if B else 0
        val, dt = cpu_iteration_sample(pkg, w, Ns, K, Bs, 1, 1, target)
        line["cpu_baseline"] = {"value": val, "unit": UNIT, "cores": _affinity_threads(), "threads": blas_threads(),
                                "kind": "port",
                                "sample": f"{Ns} of {N} start points (x{K} Koopman samples), one iteration after one "
                                          f"warm-up, minibatch {Bs}, {dt:.1f} s; per-sample throughput"}
    else:
        line["cpu_baseline"] = None
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse()
    import __graft_entry__ as g
    pkg = g.load_package()
    if args.impl == "reference":
        run_reference(args, pkg)
    else:
        run_b200(args, pkg)


if __name__ == "__main__":
    main()
