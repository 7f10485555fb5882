"""Operand formats for the wide Dense GEMMs, measured ON THE GPU and on TRAINED weights.

1. trains the c5 network [595, 2048, 2048, 1] with the library for a number of iterations (reduced N so it takes
   seconds), downloads the weights;
2. evaluates chi on a sample of rows with every candidate operand format, emulated with exact fp64 products of the
   operands rounded exactly as the tensor core would see them (torch on the GPU), against the fp64 result;
3. also compares the LIBRARY's own chi (which includes the TMEM accumulation error) with the fp64 result;
4. times cuBLASLt fp8 (e4m3) against bf16 matmuls of the same shape, sustained, to see what an fp8 correction
   term would cost under this GPU's power cap.

    python profiles/scripts/split_accuracy_gpu.py [--iters 100] > gpurun_out/split_accuracy_gpu.json
"""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))


def q_bf16(x):
    return x.to(torch.bfloat16).to(torch.float64)


def q_fp16(x):
    return x.to(torch.float16).to(torch.float64)


def q_e4m3(x):
    """e4m3 with a per-tensor power-of-two scale so that max |x| lands just below 448 (what a scaled operand
    buffer would hold)"""
    m = float(x.abs().max())
    if m == 0.0:
        return x.to(torch.float64)
    s = 2.0 ** np.floor(np.log2(448.0 / m))
    return (x.to(torch.float32) * s).to(torch.float8_e4m3fn).to(torch.float64) / s


def split(x, q):
    hi = q(x.to(torch.float32))
    lo = q((x.to(torch.float64) - hi).to(torch.float32))
    return hi, lo


def gemm(a, w, mode):
    """a (M, K), w (K, N) fp32 tensors -> fp64 product under the operand format `mode`"""
    a64, w64 = a.to(torch.float64), w.to(torch.float64)
    if mode == "fp32":
        return a64 @ w64
    if mode in ("bf16x3", "bf16x2a", "bf16x2w", "bf16x1", "fp16x3", "fp16x2a", "fp16x2w", "fp16x1"):
        q = q_bf16 if mode.startswith("bf16") else q_fp16
        ah, al = split(a, q)
        wh, wl = split(w, q)
        out = ah @ wh
        if mode.endswith("x3"):
            out = out + al @ wh + ah @ wl
        elif mode.endswith("x2a"):     # activations split, weights rounded once
            out = out + al @ wh
        elif mode.endswith("x2w"):     # weights split, activations rounded once
            out = out + ah @ wl
        return out
    if mode in ("fp16+fp8corr", "bf16+fp8corr"):
        q = q_fp16 if mode.startswith("fp16") else q_bf16
        ah, al = split(a, q)
        wh, wl = split(w, q)
        # main term on the 16-bit pipe; both correction terms as ONE fp8 MMA over the concatenated K:
        #   [q8(a_hi) | q8(a_lo)] . [q8(w_lo) ; q8(w_hi)]   (lo parts carry their own power-of-two scale)
        return ah @ wh + q_e4m3(ah.to(torch.float32)) @ q_e4m3(wl.to(torch.float32)) \
            + q_e4m3(al.to(torch.float32)) @ q_e4m3(wh.to(torch.float32))
    raise ValueError(mode)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=100)
    ap.add_argument("--N", type=int, default=131072)
    ap.add_argument("--rows", type=int, default=8192)
    args = ap.parse_args()
    import __graft_entry__ as g
    import oracle
    pkg = g.load_package()
    dev = torch.device("cuda", 0)
    w = pkg.synthetic.WORKLOADS["c5"]
    N, K = args.N, 4
    xs, ys = pkg.synthetic.make_data(w, N, K)
    model = pkg.densenet(w.widths, layernorm=True, rng=np.random.default_rng(w.seed + 1))
    out = {"workload": f"c5 network {w.widths}, N={N}, K={K}, minibatch 65536, Adam", "rows": args.rows, "stages": {}}
    data = pkg.SimulationData(xs, ys, featurizer=pkg.FeaturesAll())
    iso = pkg.Iso(data, opt=pkg.AdamRegularized(), model=model, minibatch=65536)
    perms = pkg.synthetic.make_perms(w, N, args.iters)
    sub = np.arange(0, N, N // args.rows)[:args.rows]
    feats = oracle.flatpairdists(np.ascontiguousarray(xs.T[sub]))                          # fp32 features
    done = 0
    for stage, upto in (("random init", 0), (f"after {args.iters} iterations", args.iters)):
        if upto > done:
            pkg.run_(iso, upto - done, perms=perms[done:upto])
            done = upto
        flat = iso.engine.download_params()
        om = oracle.unflatten_params(oracle.Model(list(w.widths), True), flat)
        f64 = torch.from_numpy(feats).to(dev).to(torch.float64)
        mu = f64.mean(1, keepdim=True)
        xc = f64 - mu
        xhat = (xc / torch.sqrt((xc * xc).mean(1, keepdim=True) + 1e-10)).to(torch.float32)
        W = [torch.from_numpy(x).to(dev) for x in om.W]
        b = [torch.from_numpy(x).to(dev).to(torch.float64) for x in om.b]
        gam, bet = torch.from_numpy(om.ln_scale).to(dev), torch.from_numpy(om.ln_bias).to(dev)
        # folded first layer exactly like the library: W1' = diag(gamma) W1, b1' = b1 + W1^T beta
        W1f = (gam[:, None] * W[0]).to(torch.float32)
        b1f = b[0] + W[0].to(torch.float64).T @ bet.to(torch.float64)

        def net(mode):
            z1 = torch.sigmoid(gemm(xhat, W1f, mode) + b1f).to(torch.float32)
            z2 = torch.sigmoid(gemm(z1, W[1], mode) + b[1]).to(torch.float32)
            return (z2.to(torch.float64) @ W[2].to(torch.float64) + b[2]).squeeze(1)

        ref = net("fp32")
        spread = float(ref.max() - ref.min())
        rows = {}
        for mode in ["bf16x1", "bf16x2a", "bf16x2w", "bf16x3", "fp16x1", "fp16x2a", "fp16x2w", "fp16x3",
                     "bf16+fp8corr", "fp16+fp8corr"]:
            e = net(mode) - ref
            rows[mode] = {"max_abs": float(e.abs().max()), "rms": float(torch.sqrt((e * e).mean())),
                          "max_abs_after_shiftscale": float(e.abs().max()) / spread}
        chi_lib = torch.from_numpy(np.ascontiguousarray(pkg.chis(iso)[0, sub])).to(dev).to(torch.float64)
        e = chi_lib - ref
        rows["library (3 x bf16 on tcgen05, fp32 TMEM accumulation)"] = {
            "max_abs": float(e.abs().max()), "rms": float(torch.sqrt((e * e).mean())),
            "max_abs_after_shiftscale": float(e.abs().max()) / spread}
        out["stages"][stage] = {"chi_spread": spread, "loss": float(iso.losses[-1]) if iso.losses else None,
                                "errors_vs_fp64": rows}
    # ---- what would an fp8 correction MMA cost?  sustained cuBLASLt rates under this GPU's power cap
    M = 8192
    a16 = torch.randn(M, M, device=dev, dtype=torch.bfloat16)
    b16 = torch.randn(M, M, device=dev, dtype=torch.bfloat16)
    a8 = a16.to(torch.float8_e4m3fn)
    b8 = b16.to(torch.float8_e4m3fn).t().contiguous().t()          # column-major B as cuBLASLt wants it
    one = torch.ones((), device=dev)

    def sustained(fn, seconds=3.0):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        t0, n = time.perf_counter(), 0
        while time.perf_counter() - t0 < seconds:
            for _ in range(20):
                fn()
            torch.cuda.synchronize()
            n += 20
        return 2.0 * M ** 3 * n / (time.perf_counter() - t0) / 1e12
    rates = {"bf16_tflops": sustained(lambda: a16 @ b16)}
    try:
        rates["fp8_e4m3_tflops"] = sustained(lambda: torch._scaled_mm(a8, b8, scale_a=one, scale_b=one,
                                                                        out_dtype=torch.bfloat16))
        rates["fp8_over_bf16"] = rates["fp8_e4m3_tflops"] / rates["bf16_tflops"]
    except Exception as ex:  # noqa: BLE001
        rates["fp8_error"] = repr(ex)
    out["cublaslt_sustained_8192"] = rates
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
