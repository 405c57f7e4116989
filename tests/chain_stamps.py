"""Diagnostic (not a pytest file): %globaltimer phase stamps of the fused decode chain of layer 1 in the last decode step
(64 streams, production widths, KV ~ n_prime chunks), plus CUDA-event time of whole decode steps."""
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
    B = int(os.environ.get("N_STREAMS", "64"))
    layers = int(os.environ.get("LAYERS", "4"))
    cfg = production_config()
    cfg.enc.layers, cfg.llm.layers = 1, layers
    sd = make_state_dict(cfg, seed=0, device="cuda:0", dtype=torch.bfloat16)
    eng = Engine(cfg, device=0, max_streams=B, max_batch=B)
    eng.load_state_dict(sd)
    r = LockstepRunner(eng, cfg, B)
    g = torch.Generator(device="cuda:0").manual_seed(1)
    for c in range(int(os.environ.get("PRIME", "33"))):
        pcm = 0.1 * torch.randn(B, SEG + (399 if c == 0 else 0), device="cuda:0", generator=g)
        r.step_device(pcm)
    torch.cuda.synchronize()
    for use_chain, pref in ((1, 1), (1, 0), (1, 1), (1, 0)):      # (chain, folded norms)
        eng.option("decode_chain", use_chain)
        eng.option("chain_fold", pref)
        ts = []
        for _ in range(4):
            pcm = 0.1 * torch.randn(B, SEG, device="cuda:0", generator=g)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r.step_device(pcm)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        print(f"chain={use_chain} fold={pref}: step ms {['%.2f' % t for t in ts]} (enc 1 + llm {layers} layers, {B} streams, kv {eng.kv_len(r.sids[0])})", flush=True)
    eng.option("decode_chain", 1)
    eng.option("chain_fold", int(os.environ.get("FOLD", "1")))
    eng.debug(2)
    pcm = 0.1 * torch.randn(B, SEG, device="cuda:0", generator=g)
    r.step_device(pcm)
    torch.cuda.synchronize()
    t = eng.read_tap("gemm_stamps", torch.int64).view(-1, 32)[:148].double()
    t0 = t[:, 0].min()
    rel = (t - t0) / 1e3
    fold = int(os.environ.get("FOLD", "1"))
    names = ["o_proj+reduce", "gate_up", "down+reduce", "qkv"] if fold else ["o_proj", "rows", "gate_up", "down", "rows", "qkv"]
    print("CTA start: mean %.2f max %.2f us" % (rel[:, 0].mean(), rel[:, 0].max()))
    for pi, nm in enumerate(names):
        cols = rel[:, 1 + 3 * pi: 4 + 3 * pi]
        def stat(v):
            v = v[v > 0]
            return "   -   " if v.numel() == 0 else "%6.2f/%6.2f/%6.2f" % (v.min(), v.mean(), v.max())
        print(f"phase {pi} {nm:8s}: dep resolved (min/mean/max us) {stat(cols[:, 0])} | first acc / rows start {stat(cols[:, 1])} | done {stat(cols[:, 2])}", flush=True)
    eng.close()


if __name__ == "__main__":
    main()
