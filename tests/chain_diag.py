"""Diagnostic (not a pytest file): fused decode chain vs operator-per-kernel path, step logits compared per step / stream."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch

from infinisst_b200 import production_config
from infinisst_b200.engine import Engine
from infinisst_b200.runner import LockstepRunner
from infinisst_b200.synthetic import make_audio, make_state_dict

SEG = 15360


def run(cfg, sd, B, use_chain, pdl, forced, n_chunks=1):
    eng = Engine(cfg, device=0, max_streams=B, max_batch=B)
    eng.load_state_dict(sd)
    eng.option("decode_chain", use_chain)
    eng.option("pdl", pdl)
    eng.debug(True)
    r = LockstepRunner(eng, cfg, B)
    audios = [make_audio(n_chunks * SEG / 16000.0, seed=300 + b) for b in range(B)]
    out = []
    for c in range(n_chunks):
        pcm = torch.cat([torch.cat([torch.zeros(1, 399), a[None, :SEG]], 1) if c == 0 else a[None, c * SEG:(c + 1) * SEG] for a in audios], 0)
        r.step_device(pcm, forced=None if forced is None else forced[c])
        out.append((r.last_tokens, eng.read_tap("step_logits", torch.float32).clone()))
    r.close()
    eng.close()
    return out


def main():
    for layers in (1, 2, 3):
        cfg = production_config()
        cfg.enc.layers, cfg.llm.layers = 1, layers
        sd = make_state_dict(cfg, seed=0, device="cuda:0", dtype=torch.bfloat16)
        for B in (64, 20, 1):
            ref = run(cfg, sd, B, 0, 1, None)
            forced = [t for t, _ in ref]
            V = cfg.llm.vocab
            for pdl in (1, 0):
                got = run(cfg, sd, B, 1, pdl, forced)
                la, lb = ref[0][1].view(10, B, V), got[0][1].view(10, B, V)
                d = (la - lb).abs()
                per_step = [float(d[s].max()) for s in range(10)]
                bad_streams = [b for b in range(B) if float(d[:, b].max()) > 0]
                print(f"layers={layers} B={B} pdl={pdl}: per-step max diff {['%.3g' % x for x in per_step]} "
                      f"streams differing {len(bad_streams)}/{B}", flush=True)
                if bad_streams:
                    s = next(i for i in range(10) if per_step[i] > 0)
                    b = bad_streams[0]
                    dv = d[s, b]
                    tiles = (dv.view(-1)[: (V // 128) * 128].view(-1, 128).max(dim=1).values > 0).sum().item()
                    print(f"    first bad step {s}: stream {b}: {int((dv > 0).sum())} of {V} logits differ, {tiles} of {V // 128} "
                          f"vocab tiles touched, rel-L2 {float(dv.norm() / la[s, b].norm()):.3e}", flush=True)
        del sd


if __name__ == "__main__":
    main()
