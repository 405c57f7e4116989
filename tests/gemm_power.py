"""SM clock and board power while one GEMM kernel runs back to back for a few seconds (not a pytest file): tells a
power-capped kernel (clock pulled down, ~1 kW) from one that stalls at full clock."""
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pynvml
import torch

from infinisst_b200 import tiny_config
from infinisst_b200.engine import Engine


def sample(stop, out, h):
    while not stop.is_set():
        out.append((pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetPowerUsage(h) / 1e3))
        time.sleep(0.05)


def main():
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(0)
    dev = "cuda:0"
    eng = Engine(tiny_config(), device=0, max_streams=2)
    secs = float(os.environ.get("SECS", "3"))
    for (name, M, N, K, dual) in [("square 8192", 8192, 8192, 8192, False), ("pre gateup", 1408, 14336, 4096, True),
                                  ("pre down", 1408, 4096, 14336, False)]:
        rows = N * (2 if dual else 1)
        a = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
        w = (torch.randn(rows, K, device=dev) * K ** -0.5).bfloat16()
        fl = 2.0 * M * rows * K
        for mode in ("cublas", "pair", "single"):
            if mode != "cublas":
                eng.option("gemm_pair", 1 if mode == "pair" else 0)
            fn = (lambda: torch.matmul(a, w.t())) if mode == "cublas" else (lambda: eng.op_gemm(a, w, dual=dual))
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            stop, smp = threading.Event(), []
            th = threading.Thread(target=sample, args=(stop, smp, h))
            th.start()
            n, t0 = 0, time.time()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            while time.time() - t0 < secs:
                for _ in range(20):
                    fn()
                n += 20
                torch.cuda.synchronize()
            e1.record()
            torch.cuda.synchronize()
            stop.set()
            th.join()
            us = e0.elapsed_time(e1) / n * 1e3
            smp = smp[len(smp) // 3:]                    # settled part
            clk = sorted(s[0] for s in smp)[len(smp) // 2]
            pw = sorted(s[1] for s in smp)[len(smp) // 2]
            print(f"{name:12s} {mode:7s} {us:8.1f} us {fl / us / 1e6:7.1f} TF/s  sm_clock median {clk} MHz  power median {pw:.0f} W", flush=True)
            time.sleep(1.0)
    eng.close()


if __name__ == "__main__":
    main()
