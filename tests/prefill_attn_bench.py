"""Diagnostic (not a pytest file): prefill attention class time per step (isst_profile) for a few L2 look-ahead depths."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from infinisst_b200 import production_config
from infinisst_b200.engine import Engine
from infinisst_b200.runner import LockstepRunner
from infinisst_b200.synthetic import make_state_dict

SEG = 15360


def main():
    B, layers = 64, 4
    cfg = production_config()
    cfg.enc.layers, cfg.llm.layers = 1, layers
    sd = make_state_dict(cfg, seed=0, device="cuda:0", dtype=torch.bfloat16)
    eng = Engine(cfg, device=0, max_streams=B, max_batch=B)
    eng.load_state_dict(sd)
    r = LockstepRunner(eng, cfg, B)
    g = torch.Generator(device="cuda:0").manual_seed(1)
    for c in range(33):
        r.step_device(0.1 * torch.randn(B, SEG + (399 if c == 0 else 0), device="cuda:0", generator=g))
    torch.cuda.synchronize()
    for ahead in (0, 1, 2, 4, 8, 2, 0):
        eng.option("prefill_l2_ahead", ahead)
        eng.profile(True)
        eng.profile_reset()
        for _ in range(3):
            r.step_device(0.1 * torch.randn(B, SEG, device="cuda:0", generator=g))
        torch.cuda.synchronize()
        p = eng.profile_read()["attn_prefill"]
        eng.profile(False)
        print(f"l2_ahead={ahead}: attn_prefill {p['ms'] / p['launches'] * 1e3:.1f} us per launch, {p['bytes'] / p['ms'] / 1e6:.0f} GB/s", flush=True)
    eng.close()


if __name__ == "__main__":
    main()
