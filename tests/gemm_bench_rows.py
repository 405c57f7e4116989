"""GPU micro-benchmark: the LLM decode GEMM shapes at 64 / 128 / 256 token rows (greedy 64 streams; beam search
4 x 32 / 4 x 64 rows) in both tile modes (force_swap 0: tokens on the 128-lane operand, 1: weights on it).
    python tests/gemm_bench_rows.py
Not a pytest file.  Weights rotate over distinct buffers so no launch finds its weights in L2."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from infinisst_b200 import tiny_config
from infinisst_b200.engine import Engine

SHAPES = [("qkv", 6144, 4096, {}), ("o", 4096, 4096, {"resid": True}), ("gateup", 14336, 4096, {"dual": True}),
          ("down", 4096, 14336, {"resid": True})]


def bench(eng, M, N, K, kw, force_swap, iters=20):
    dev = "cuda:0"
    dual = kw.get("dual", False)
    rows = N * (2 if dual else 1)
    nbuf = max(2, min(16, int(400e6 // (rows * K * 2)) + 1))
    ws = [(torch.randn(rows, K, device=dev) * K ** -0.5).bfloat16() for _ in range(nbuf)]
    a = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
    resid = torch.randn(M, N, device=dev).bfloat16() if kw.get("resid") else None
    call = lambda i: eng.op_gemm(a, ws[i % nbuf], resid=resid, dual=dual, force_swap=force_swap)
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
    return us, 2.0 * M * rows * K / us / 1e6, rows * K * 2 / us / 1e3


def main():
    eng = Engine(tiny_config(), device=0, max_streams=2)
    for M in (64, 128, 256):
        tot = {0: 0.0, 1: 0.0}
        for (name, N, K, kw) in SHAPES:
            row = f"M={M:4d} {name:7s}"
            for fs in (0, 1):
                try:
                    us, tf, gbs = bench(eng, M, N, K, kw, fs)
                    tot[fs] += us
                    row += f"   swap={fs}: {us:7.1f} us {tf:7.1f} TF/s {gbs:7.1f} GB/s(w)"
                except Exception as ex:          # noqa: BLE001
                    row += f"   swap={fs}: failed ({str(ex)[:60]})"
            print(row, flush=True)
        print(f"M={M:4d} layer sum: swap=0 {tot[0]:.1f} us, swap=1 {tot[1]:.1f} us (HBM ideal 67 us, tensor ideal {2.0 * M * 218.1e6 / 1.4e9:.0f} us)", flush=True)
    eng.close()


if __name__ == "__main__":
    main()
