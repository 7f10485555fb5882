"""Why three bf16 MMAs per product: error of the wide c5 network (F=595 -> 2048 -> 2048 -> 1) on chi when the two
Dense GEMMs run with different operand formats, emulated on the CPU (products exact in fp64, operands rounded as
the tensor core would see them, fp32 accumulation error ignored -- it is common to all variants).

    python profiles/scripts/split_accuracy.py      # prints the table quoted in DESIGN.md section 4
"""
import numpy as np


def rne(x, keep_bits):
    """round fp32 to `keep_bits` explicit mantissa bits (7 = bf16, 10 = tf32/fp16 mantissa), nearest-even"""
    u = np.asarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    drop = 23 - keep_bits
    u = u + ((1 << (drop - 1)) - 1) + ((u >> drop) & 1)
    u = (u >> drop) << drop
    return u.astype(np.uint32).view(np.float32)


def split(x, bits):
    hi = rne(x, bits)
    lo = rne((np.asarray(x, np.float32) - hi).astype(np.float32), bits)
    return hi.astype(np.float64), lo.astype(np.float64)


def gemm(a, w, mode):
    a = a.astype(np.float32); w = w.astype(np.float32)
    if mode == "fp32":
        return a.astype(np.float64) @ w.astype(np.float64)
    bits = {"bf16": 7, "tf32": 10, "fp16": 10}[mode.split(":")[0]]
    terms = int(mode.split(":")[1])
    ah, al = split(a, bits)
    wh, wl = split(w, bits)
    out = ah @ wh
    if terms >= 2:
        out += al @ wh
    if terms >= 3:
        out += ah @ wl
    return out


def main():
    rng = np.random.default_rng(0)
    M, F, H = 4096, 595, 2048
    x = rng.normal(size=(M, F)).astype(np.float32)                      # LayerNorm output
    lim = lambda i, o: np.sqrt(6.0 / (i + o))
    W1 = rng.uniform(-lim(F, H), lim(F, H), size=(F, H)).astype(np.float32)
    W2 = rng.uniform(-lim(H, H), lim(H, H), size=(H, H)).astype(np.float32)
    W3 = rng.uniform(-lim(H, 1), lim(H, 1), size=(H, 1)).astype(np.float32)
    sig = lambda z: 1.0 / (1.0 + np.exp(-z))

    def net(mode):
        a1 = sig(gemm(x, W1, mode)).astype(np.float32)
        a2 = sig(gemm(a1, W2, mode)).astype(np.float32)
        return a2.astype(np.float64) @ W3.astype(np.float64)

    ref = net("fp32")
    print("| operands / MMAs per product | max abs error on chi | rms |")
    print("|---|---:|---:|")
    for mode in ["bf16:1", "bf16:2", "bf16:3", "tf32:1", "fp16:2", "fp16:3"]:
        e = net(mode) - ref
        print(f"| {mode.replace(':', ' x ')} | {np.abs(e).max():.2e} | {np.sqrt((e ** 2).mean()):.2e} |")
    print(f"(chi spread for scale: std {ref.std():.3f})")


if __name__ == "__main__":
    main()
