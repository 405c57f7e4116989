"""One GEMM shape launched a few times (for ncu captures; not a pytest file):  python tests/one_gemm.py M N K [pair=0|1] [bias] [resid] [dual] [gelu]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from infinisst_b200 import tiny_config
from infinisst_b200.engine import Engine

M, N, K = (int(x) for x in sys.argv[1:4])
flags = sys.argv[4:]
eng = Engine(tiny_config(), device=0, max_streams=2)
eng.option("gemm_pair", 0 if "pair=0" in flags else 1)
dev = "cuda:0"
dual = "dual" in flags
a = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
w = (torch.randn(N * (2 if dual else 1), K, device=dev) * K ** -0.5).bfloat16()
bias = torch.randn(N, device=dev) if "bias" in flags else None
resid = torch.randn(M, N, device=dev).bfloat16() if "resid" in flags else None
for _ in range(4):
    eng.op_gemm(a, w, bias=bias, gelu="gelu" in flags, resid=resid, dual=dual)
torch.cuda.synchronize()
eng.close()
