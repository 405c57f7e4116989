"""GPU micro-benchmark of the GEMM kernels on the production shapes (not a pytest file).
    python tests/gemm_bench.py            # stream-K tcgen05 kernel
Weights rotate over enough distinct buffers that no launch finds its weights in L2."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from infinisst_b200 import tiny_config
from infinisst_b200.engine import Engine

SHAPES = [
    # name, M, N, K, kwargs
    ("dec qkv", 64, 6144, 4096, {}), ("dec o", 64, 4096, 4096, {"resid": True}),
    ("dec gateup", 64, 14336, 4096, {"dual": True}), ("dec down", 64, 4096, 14336, {"resid": True}),
    ("dec lm_head", 64, 128263, 4096, {"out_f32": True}),
    ("b1 qkv", 1, 6144, 4096, {}), ("b1 gateup", 1, 14336, 4096, {"dual": True}), ("b1 down", 1, 4096, 14336, {"resid": True}),
    ("pre qkv", 1408, 6144, 4096, {}), ("pre o", 1408, 4096, 4096, {"resid": True}),
    ("pre gateup", 1408, 14336, 4096, {"dual": True}), ("pre down", 1408, 4096, 14336, {"resid": True}),
    ("enc qkv", 3072, 3072, 1024, {"bias": True}), ("enc out", 3072, 1024, 1024, {"bias": True, "resid": True}),
    ("enc fc1", 3072, 4096, 1024, {"bias": True, "gelu": True}), ("enc fc2", 3072, 1024, 4096, {"bias": True, "resid": True}),
    ("beam qkv", 256, 6144, 4096, {}), ("beam gateup", 256, 14336, 4096, {"dual": True}), ("beam down", 256, 4096, 14336, {"resid": True}),
]


def bench(eng, name, M, N, K, kw, iters=20):
    dev = "cuda:0"
    dual = kw.get("dual", False)
    rows = N * (2 if dual else 1)
    nbuf = max(2, min(16, int(400e6 // (rows * K * 2)) + 1))
    ws = [(torch.randn(rows, K, device=dev) * K ** -0.5).bfloat16() for _ in range(nbuf)]
    a = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
    bias = torch.randn(N, device=dev) if kw.get("bias") else None
    resid = torch.randn(M, N, device=dev).bfloat16() if kw.get("resid") else None
    call = lambda i: eng.op_gemm(a, ws[i % nbuf], bias=bias, gelu=kw.get("gelu", False), resid=resid, dual=dual,
                                 out_f32=kw.get("out_f32", False))
    for i in range(3):
        call(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        call(i)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / iters * 1e3
    flops = 2.0 * M * rows * K
    by = rows * K * 2 + M * K * 2 + M * N * (4 if kw.get("out_f32") else 2) * (2 if resid is not None else 1)
    return us, flops / us / 1e6, by / us / 1e3


def stamps(eng, M, N, K, kw):
    dev = "cuda:0"
    dual = kw.get("dual", False)
    w = (torch.randn(N * (2 if dual else 1), K, device=dev) * K ** -0.5).bfloat16()
    a = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
    resid = torch.randn(M, N, device=dev).bfloat16() if kw.get("resid") else None
    eng.debug(2)
    for _ in range(2):
        eng.op_gemm(a, w, bias=(torch.randn(N, device=dev) if kw.get("bias") else None), gelu=kw.get("gelu", False), resid=resid, dual=dual, out_f32=kw.get("out_f32", False))
    torch.cuda.synchronize()
    t = eng.read_tap("gemm_stamps", torch.int64).view(4096, 8)[:148].double()
    eng.debug(0)
    t = t[t[:, 0] > 0]                  # CTAs that ran (grids smaller than the SM count leave zero rows)
    t0 = t[:, 0].min()
    rel = (t[:, :6] - t0) / 1e3
    return [f"{rel[:, i].mean():.1f}/{rel[:, i].max():.1f}" for i in (0, 1, 2, 3, 5, 4)]


def main():
    out = []
    for mode in (["pair", "single"] if "--ab" in sys.argv else ["sk"]):
        eng = Engine(tiny_config(), device=0, max_streams=2)
        if "--no-pair" in sys.argv or mode == "single":
            eng.option("gemm_pair", 0)              # one CTA per tile above 128 rows (A/B against the CTA-pair kernel)
        for (name, M, N, K, kw) in SHAPES:
            if mode != "sk" and M <= 128:
                continue
            us, tf, gbs = bench(eng, name, M, N, K, kw)
            line = f"[{mode}] {name:12s} M={M:5d} N={N:6d} K={K:5d}  {us:8.1f} us  {tf:7.1f} TFLOP/s  {gbs:7.1f} GB/s"
            if mode == "sk" and "--stamps" in sys.argv:
                line += "  stamps(start, first_tma, first_acc, walk_done, reduce_done, exit) mean/max us: " + " ".join(stamps(eng, M, N, K, kw))
            print(line, flush=True)
        eng.close()


if __name__ == "__main__":
    main()
