"""GPU micro-benchmark of decode attention over synthetic paged KV (not a pytest file)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from infinisst_b200 import production_config
from infinisst_b200.engine import Engine


def main():
    cfg = production_config()
    cfg.enc.layers, cfg.llm.layers = 1, 8
    n = int(os.environ.get("N_STREAMS", "64"))
    eng = Engine(cfg, device=0, max_streams=2 * n + 2, max_batch=n)
    splits = int(os.environ.get("DEC_SPLITS", "0"))          # 0 = automatic
    eng.option("decode_splits", splits)
    for L in (250, 500, 1000):
        ms = eng.decode_attention_bench(n, L, 40)
        by = n * (L + 1) * cfg.llm.kv_heads * cfg.llm.head_dim * 2 * 2
        print(f"splits={splits or 'auto'!s:>4s} n={n} L={L:5d}: {ms * 1e3:7.1f} us  {by / ms / 1e6:7.1f} GB/s", flush=True)
    eng.close()


if __name__ == "__main__":
    main()
