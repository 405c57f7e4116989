"""A/B of the two GEMM kernels and torch.matmul (cuBLAS) with the modes interleaved round by round, so that clock /
power drift hits all of them alike (not a pytest file).  Weights rotate over distinct buffers as in gemm_bench.py."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from infinisst_b200 import tiny_config
from infinisst_b200.engine import Engine

SHAPES = [
    ("square 8192", 8192, 8192, 8192, {}),
    ("pre qkv", 1408, 6144, 4096, {}), ("pre o", 1408, 4096, 4096, {"resid": True}),
    ("pre gateup", 1408, 14336, 4096, {"dual": True}), ("pre down", 1408, 4096, 14336, {"resid": True}),
    ("enc qkv", 3072, 3072, 1024, {"bias": True}), ("enc out", 3072, 1024, 1024, {"bias": True, "resid": True}),
    ("enc fc1", 3072, 4096, 1024, {"bias": True, "gelu": True}), ("enc fc2", 3072, 1024, 4096, {"bias": True, "resid": True}),
]


def main():
    dev = "cuda:0"
    eng = Engine(tiny_config(), device=0, max_streams=2)
    rounds, iters = 6, 10
    for (name, M, N, K, kw) in SHAPES:
        dual = kw.get("dual", False)
        rows = N * (2 if dual else 1)
        nbuf = max(2, min(8, int(400e6 // (rows * K * 2)) + 1))
        ws = [(torch.randn(rows, K, device=dev) * K ** -0.5).bfloat16() for _ in range(nbuf)]
        a = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
        bias = torch.randn(N, device=dev) if kw.get("bias") else None
        resid = torch.randn(M, N, device=dev).bfloat16() if kw.get("resid") else None
        mine = lambda i: eng.op_gemm(a, ws[i % nbuf], bias=bias, gelu=kw.get("gelu", False), resid=resid, dual=dual)
        blas = lambda i: torch.matmul(a, ws[i % nbuf].t())
        tot = {"pair": 0.0, "single": 0.0, "cublas": 0.0}
        for r in range(rounds + 1):
            for mode in ("pair", "single", "cublas"):
                if mode != "cublas":
                    eng.option("gemm_pair", 1 if mode == "pair" else 0)
                fn = blas if mode == "cublas" else mine
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for i in range(iters):
                    fn(i)
                e1.record()
                torch.cuda.synchronize()
                if r > 0:                              # round 0 is warm-up
                    tot[mode] += e0.elapsed_time(e1) / iters * 1e3 / rounds
        fl = 2.0 * M * rows * K
        print(f"{name:12s} M={M:5d} N={N:6d} K={K:5d} " +
              "  ".join(f"{m} {tot[m]:7.1f} us {fl / tot[m] / 1e6:7.1f} TF/s" for m in tot), flush=True)
    eng.close()


if __name__ == "__main__":
    main()
