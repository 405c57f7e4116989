"""Bring-up / A-B of the CTA-pair GEMM (not a pytest file): correctness of every epilogue against torch fp32, then
timings against the one-CTA-per-tile kernel on the production prefill / encoder shapes."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from infinisst_b200 import tiny_config
from infinisst_b200.engine import Engine
from gemm_bench import SHAPES, bench

CASES = [
    (256, 256, 256, {}), (256, 384, 512, {}), (200, 384, 512, {}), (513, 520, 256, {"out_f32": True}),
    (1408, 6144, 4096, {}), (1408, 4096, 4096, {"resid": True}), (1408, 4096, 14336, {"resid": True}),
    (3072, 4096, 1024, {"bias": True, "gelu": True}), (3072, 1024, 4096, {"bias": True, "resid": True}),
    (300, 768, 512, {"dual": True}), (400, 640, 1024, {"dual": True}), (1408, 14336, 4096, {"dual": True}),
    (129, 128, 256, {}), (4000, 256, 320, {"bias": True}),
]


def rel_l2(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-20)).item()


def main():
    dev = "cuda:0"
    eng = Engine(tiny_config(), device=0, max_streams=2)
    torch.manual_seed(0)
    bad = 0
    for (M, N, K, kw) in CASES:
        a = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
        dual = kw.get("dual", False)
        w = (torch.randn(N * (2 if dual else 1), K, device=dev) * (K ** -0.5)).bfloat16()
        bias = torch.randn(N, device=dev) if kw.get("bias") else None
        resid = torch.randn(M, N, device=dev).bfloat16() if kw.get("resid") else None
        ref = a.float() @ w.float().t()
        if dual:
            ref = torch.nn.functional.silu(ref[:, :N]) * ref[:, N:]
        if bias is not None:
            ref = ref + bias
        if kw.get("gelu"):
            ref = torch.nn.functional.gelu(ref)
        if resid is not None:
            ref = ref + resid.float()
        n0 = eng.path_count("gemm_pair") + eng.path_count("gemm_pair_dual")
        out = eng.op_gemm(a, w, bias=bias, gelu=kw.get("gelu", False), resid=resid, dual=dual, out_f32=kw.get("out_f32", False))
        torch.cuda.synchronize()
        took = eng.path_count("gemm_pair") + eng.path_count("gemm_pair_dual") - n0
        err = rel_l2(out, ref)
        ok = err < (2e-5 if kw.get("out_f32") else 4e-3) and took == 1
        bad += not ok
        print(f"{'ok ' if ok else 'BAD'} M={M} N={N} K={K} {kw} rel_l2={err:.2e} pair_launches={took}", flush=True)
    if bad:
        print("FAILED", bad)
        sys.exit(1)
    for mode in ("pair", "single"):
        eng.option("gemm_pair", 1 if mode == "pair" else 0)
        for (name, M, N, K, kw) in SHAPES:
            if M <= 128:
                continue
            us, tf, gbs = bench(eng, name, M, N, K, kw)
            print(f"[{mode}] {name:12s} M={M:5d} N={N:6d} K={K:5d}  {us:8.1f} us  {tf:7.1f} TFLOP/s", flush=True)
    eng.close()


if __name__ == "__main__":
    main()
