#!/usr/bin/env python
"""bench.py -- Koopman samples/s per ISOKANN iteration (BASELINE.json metric).

One "step" is one body of the reference's run! loop (src/iso.jl:72-94): Koopman target
(featurize -> chi over all K*N samples -> K-mean -> isotarget) followed by one training epoch
(minibatched fwd/bwd + optimiser steps over the N start points).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config c5] [--impl reference]

N > 1 is launched by torchrun, one rank per GPU; start points are sharded over ranks
(strong scaling: the workload is fixed).  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "koopman_samples_per_s_per_isokann_iteration"
UNIT = "samples/s"


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
    ap.add_argument("--gemm", default="auto", choices=["auto", "fp32", "tc"])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--profile", action="store_true",
                    help="short run for ncu: no e2e / CPU baseline, warm-up not forced to 3; never a bench value")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample", type=int, default=0, help="start points in the CPU baseline sample")
    return ap.parse_args()


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return {"hbm_gbs": d["hbm_gbs"], "tflops": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "src": "measured"}
    return {"hbm_gbs": 6650.0, "tflops": 1400.0, "src": "fallback"}


def workload_config(w, N, K, B, extra=None):
    cfg = {"workload": f"{w.name}: {w.n_atoms}-atom pairdist featurizer F={w.F}, pairnet {w.widths}, N={N}, K={K}, "
                       f"{w.target} target, {w.opt}, minibatch={B}",
           "N": N, "K": K, "minibatch": B, "widths": list(w.widths), "target": w.target, "optimiser": w.opt,
           "l2": "inputs (coords of K*N samples) larger than the 126 MB L2" if N * K * w.D * 4 > 126e6
                 else "inputs fit in L2; steady-state iteration"}
    if extra:
        cfg.update(extra)
    return cfg


# ---------------------------------------------------------------------------------------------
# CPU baseline: the oracle port of the reference's run! on the host cores, bounded sample
# ---------------------------------------------------------------------------------------------
def cpu_iteration_sample(pkg, w, Ns, K, B, steps, warmup):
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
    kw = {}
    for i in range(warmup):
        oracle.run(m, xsf, ysf, cfg, st, 1, B, [perms[i]], w.target, **kw)
    t0 = time.perf_counter()
    for i in range(steps):
        oracle.run(m, xsf, ysf, cfg, st, 1, B, [perms[warmup + i]], w.target, **kw)
    dt = (time.perf_counter() - t0) / steps
    return Ns * K / dt, dt


def cpu_sample_size(w, N, K, requested):
    if requested > 0:
        return min(N, requested)
    # aim at ~10-30 s of CPU work for warmup+steps iterations: ~2e12 flop per iteration
    flop_per_start = 2.0 * w.macs() * (K + 3)
    ns = int(2.0e12 / max(flop_per_start, 1.0))
    ns = max(256, min(N, ns))
    return ns


def n_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def run_reference(args, pkg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = pkg.synthetic.WORKLOADS[args.config]
    N = args.N or w.N
    K = args.K or w.K
    B = args.minibatch if args.minibatch is not None else w.minibatch
    Ns = cpu_sample_size(w, N, K, args.cpu_sample)
    Bs = min(B, Ns) if B else 0
    steps, warm = max(1, args.steps), max(1, min(args.warmup, 1))
    val, dt = cpu_iteration_sample(pkg, w, Ns, K, Bs, steps, warm)
    sample = (f"{Ns} of {N} start points (x{K} Koopman samples), one full iteration each step, minibatch {Bs}; "
              f"throughput is per-sample so it carries to the full N at fixed minibatch")
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warm, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": workload_config(w, N, K, B, {"note": "CPU restatement of reference run! (Flux semantics) in numpy/"
                                               "OpenBLAS; Julia is not installable here, so the reference itself "
                                               "cannot be timed"}),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": n_cores(), "kind": "port", "sample": sample},
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
def device_data(pkg, w, N, K, off, n_loc, device):
    """synthetic coordinates generated on the device: xs (N, D) on every rank, ys (n_loc, K, D) shard"""
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
    g2 = torch.Generator(device=device)
    g2.manual_seed(99 + w.seed + off)
    ys = torch.empty((n_loc, K, w.D), dtype=torch.float32, device=device)
    step = max(1, (1 << 26) // (K * w.D))
    for s in range(0, n_loc, step):
        e = min(n_loc, s + step)
        ys[s:e] = xs[off + s:off + e, None, :] + 0.03 * torch.randn((e - s, K, w.D), generator=g2, device=device)
    return xs.contiguous(), ys


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
    w = pkg.synthetic.WORKLOADS[args.config]
    N = args.N or w.N
    K = args.K or w.K
    B = args.minibatch if args.minibatch is not None else w.minibatch
    off, n_loc = pkg.parallel.shard_range(N, world, rank)

    rngp = np.random.default_rng(w.seed + 1)
    model = pkg.densenet(w.widths, layernorm=w.layernorm, rng=rngp)
    rule = pkg.AdamRegularized() if w.opt == "adam" else pkg.NesterovRegularized()
    eng = pkg.Engine(model, rule, "allpairs" if w.featurizer == "allpairs" else "identity", w.n_atoms, None,
                     device=local, gemm=args.gemm)
    if world > 1:
        uid = pkg.parallel.broadcast_unique_id(rank)
        eng.comm_init(world, rank, uid)
    xs, ys = device_data(pkg, w, N, K, off, n_loc, device)
    torch.cuda.synchronize()
    eng.set_data_dev(xs, ys, w.D, K, N, off, n_loc)
    nsteps, nwarm = args.steps, (args.warmup if args.profile else max(3, args.warmup))
    if args.profile:
        args.no_e2e = args.no_cpu_baseline = True
    perms = pkg.synthetic.make_perms(w, N, nwarm + nsteps)
    opts = {}

    def barrier():
        eng.synchronize()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    ext = torch.cuda.ExternalStream(eng.stream(), device=device)

    def timed(fn, steps):
        """device time of `steps` calls of fn, bracketed by barrier+sync, CUDA events on the library's stream"""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        with torch.cuda.stream(ext):
            e0.record()
        for i in range(steps):
            fn(i)
        with torch.cuda.stream(ext):
            e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    # ---- resident-data throughput (`value`): no per-kernel timers in this pass ----
    for i in range(nwarm):
        eng.iterate(w.target, 1, 1, B, perms[i], **opts)
    eng.reset_stats()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms = timed(lambda i: eng.iterate(w.target, 1, 1, B, perms[nwarm + i], **opts), nsteps)
    clocks = sampler.stop() if rank == 0 else None
    launches = eng.stats()["kernel_launches"]
    value = N * K * nsteps / (ms * 1e-3)
    # ---- same steps again with CUDA events around every kernel launch (roofline, phase split) ----
    eng.reset_stats()
    eng.enable_timing(True)
    timed(lambda i: eng.iterate(w.target, 1, 1, B, perms[nwarm + i], **opts), nsteps)
    st = eng.stats()
    eng.enable_timing(False)

    # ---- end to end through the public API with HOST buffers ----
    e2e = None
    if not args.no_e2e:
        xs_h = torch.empty(xs.shape, dtype=torch.float32, pin_memory=True)
        ys_h = torch.empty(ys.shape, dtype=torch.float32, pin_memory=True)
        xs_h.copy_(xs)
        ys_h.copy_(ys)
        torch.cuda.synchronize()
        xs_j, ys_j = xs_h.numpy().T, ys_h.numpy().T          # Julia-shaped (D, N), (D, K, n_loc) views
        nparams = eng.P

        def e2e_step(i):
            # SimulationData upload + run!(iso, 1) + the loss and cpu(iso) (the updated model) back on the host
            eng.set_data_async(xs_j, ys_j, n_offset=off, n_local=n_loc)
            eng.iterate(w.target, 1, 1, B, perms[i % len(perms)], **opts)
            eng.download_params()
        for i in range(2):
            e2e_step(i)
        ms_e = timed(e2e_step, nsteps)
        # whole job, all ranks together: every rank uploads its shard of ys, its own rows of xs (the other rows come
        # from the peers over NVLink) and the permutation; every rank reads the loss and the parameters back
        h2d = 4 * (xs.numel() + N * K * xs.shape[1]) + 8 * N * world
        d2h = (8 + 4 * nparams) * world
        e2e = {"value": N * K * nsteps / (ms_e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "ms_per_step": ms_e / nsteps,
               "call": "isokann_set_data_async (pinned host xs, ys streamed in behind the Koopman pass) + run!(iso,1) + loss and cpu(iso) parameters to host"}
        eng.set_data_dev(xs, ys, w.D, K, N, off, n_loc)

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    pk = peaks()
    traffic = None
    tf = ROOT / "profiles" / "traffic.json"
    if tf.exists() and N == w.N and K == w.K:
        traffic = json.loads(tf.read_text()).get(args.config, {}).get("traffic_bytes_per_launch")
    tensor_bound = max(w.widths[1:-1] or [0]) >= 256
    if tensor_bound:
        ach = st["gemm_flops"] / (st["ms_gemm"] * 1e-3) / 1e12 if st["ms_gemm"] > 0 else 0.0
        roof = {"bound": "tensor", "kernel": "dense-layer GEMM (fused bias+activation)", "achieved": ach,
                "peak": pk["tflops"], "unit": "TFLOP/s", "frac": ach / pk["tflops"], "traffic": traffic,
                "executed_tflops": 3 * ach,
                "peak_source": f"{pk['src']} bf16 sustained", "launches": st["n_gemm_launches"],
                "avg_launch_ms": st["ms_gemm"] / max(1, st["n_gemm_launches"]),
                "note": "algorithmic 2*M*N*K flops of all GEMM launches / their CUDA-event time; every k-slice is "
                        "3 bf16 MMAs (hi*hi, hi*lo, lo*hi) to keep fp32 accuracy, so frac <= 1/3 by construction and "
                        "executed_tflops = 3*achieved is what the tensor pipe runs; traffic = mean dram bytes per "
                        "launch from profiles/traffic.json (ncu)"}
    else:
        ach = st["featurize_bytes"] / (st["ms_featurize"] * 1e-3) / 1e9 if st["ms_featurize"] > 0 else 0.0
        roof = {"bound": "hbm", "kernel": "featurizer (pair distances + LayerNorm)", "achieved": ach,
                "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": ach / pk["hbm_gbs"], "traffic": traffic,
                "peak_source": pk["src"], "launches": st["n_featurize_launches"],
                "avg_launch_ms": st["ms_featurize"] / max(1, st["n_featurize_launches"]),
                "note": "algorithmic 4*(D+F) bytes per record / CUDA-event time of the featurizer launches"}
    # the other kernel north_star asks a roofline figure for: the featurizer against measured HBM bandwidth
    fz = None
    if st["ms_featurize"] > 0 and st["featurize_bytes"] > 0:
        fa = st["featurize_bytes"] / (st["ms_featurize"] * 1e-3) / 1e9
        fz = {"bound": "hbm", "achieved": fa, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": fa / pk["hbm_gbs"],
              "launches": st["n_featurize_launches"],
              "note": "algorithmic 4*(D+F) bytes per record / CUDA-event time of the featurizer launches, timed inside "
                      "the step (power-capped SM clock); split bf16 output moves 4*(D+ld) bytes, ld = F padded to 64"}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": nsteps, "warmup": nwarm,
        "ms_per_step": ms / nsteps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": workload_config(w, N, K, B, {"parallelism": f"start points sharded over {world} GPU(s)",
                                               "gemm": args.gemm}),
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
        "roofline": roof, "roofline_featurizer": fz,
        "phase_ms_per_step": {"koopman": st["ms_koopman_total"] / nsteps, "target": st["ms_target_total"] / nsteps,
                              "train": st["ms_train_total"] / nsteps},
        "kernel_ms_per_step": {"featurize": st["ms_featurize"] / nsteps, "gemm": st["ms_gemm"] / nsteps,
                               "reduce": st["ms_reduce"] / nsteps, "train_elementwise": st["ms_train_elementwise"] / nsteps,
                               "optimiser": st["ms_optimiser"] / nsteps, "nccl_allreduce": st["ms_nccl"] / nsteps},
    }
    if not args.no_cpu_baseline and world == 1:
        Ns = cpu_sample_size(w, N, K, args.cpu_sample)
        Bs = min(B, Ns) if B else 0
        val, dt = cpu_iteration_sample(pkg, w, Ns, K, Bs, 1, 1)
        line["cpu_baseline"] = {"value": val, "unit": UNIT, "cores": n_cores(), "kind": "port",
                                "sample": f"{Ns} of {N} start points (x{K} Koopman samples), one iteration, "
                                          f"minibatch {Bs}, {dt:.1f} s; per-sample throughput"}
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
