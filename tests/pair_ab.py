"""A/B of an engine option on the tensor-bound GEMM shapes, modes interleaved round by round (not a pytest file):
    python tests/pair_ab.py gemm_pair_l2_ahead 0 8 16"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from infinisst_b200 import tiny_config
from infinisst_b200.engine import Engine
from gemm_bench import SHAPES


def main():
    key, vals = sys.argv[1], [int(v) for v in sys.argv[2:]]
    dev = "cuda:0"
    eng = Engine(tiny_config(), device=0, max_streams=2)
    rounds, iters = 5, 10
    for (name, M, N, K, kw) in SHAPES:
        if M <= 256:
            continue
        dual = kw.get("dual", False)
        rows = N * (2 if dual else 1)
        nbuf = max(2, min(8, int(400e6 // (rows * K * 2)) + 1))
        ws = [(torch.randn(rows, K, device=dev) * K ** -0.5).bfloat16() for _ in range(nbuf)]
        a = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
        bias = torch.randn(N, device=dev) if kw.get("bias") else None
        resid = torch.randn(M, N, device=dev).bfloat16() if kw.get("resid") else None
        tot = {v: 0.0 for v in vals}
        for r in range(rounds + 1):
            for v in vals:
                eng.option(key, v)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for i in range(iters):
                    eng.op_gemm(a, ws[i % nbuf], bias=bias, gelu=kw.get("gelu", False), resid=resid, dual=dual)
                e1.record()
                torch.cuda.synchronize()
                if r > 0:
                    tot[v] += e0.elapsed_time(e1) / iters * 1e3 / rounds
        print(f"{name:12s} M={M:5d} N={N:6d} K={K:5d} " + "  ".join(f"{key}={v}: {tot[v]:7.1f} us" for v in vals), flush=True)
    eng.close()


if __name__ == "__main__":
    main()
