"""BASELINE.json configs[3]: ONE unbounded stream of `--seconds` (default 3600 s = 3750 chunks of 960 ms) of synthetic
16 kHz audio through the reference-facing API (model.generate with the agent's kwargs + sliding-window eviction) at
production size (wav2vec2-large + Llama-3.1-8B, bf16 random init, max_llm_cache_size 1000 + pinned 40-token system
prompt, encoder window 576 frames).  Checks, for every chunk: the eviction (kept index ranges) equals the oracle's
integer model of agents/infinisst.py:337-361 bit-exactly, the KV length stays bounded, the page pool does not
leak; reports wall time, speech-s/s and chunk latency percentiles as one JSON line.

    python tests/hour_stream.py [--seconds 3600] [--beam 1] > profiles/rNN_hour_stream.json
Not a bench value (the oracle integer model is the checker here, tests-style); needs a B200."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

CHUNK = 15360


def pct(v, q):
    s = sorted(v)
    return s[min(len(s) - 1, int(q * len(s)))]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=3600.0)
    ap.add_argument("--beam", type=int, default=1)
    args = ap.parse_args()
    from infinisst_b200 import production_config
    from infinisst_b200.engine import Engine
    from infinisst_b200.runner import LockstepRunner
    from infinisst_b200.synthetic import make_state_dict
    from oracle import infinisst_oracle as O          # integer eviction model: the checker

    cfg = production_config()
    g = cfg.gen
    eng = Engine(cfg, device=0, max_streams=2, max_batch=args.beam, max_prompt=64, max_beams=args.beam)
    eng.load_state_dict(make_state_dict(cfg, seed=0, device="cuda:0", dtype=torch.bfloat16))
    free0 = eng.pages_free()
    run = LockstepRunner(eng, cfg, 1, beam=args.beam)
    n_chunks = int(round(args.seconds / 0.96))
    gen = torch.Generator().manual_seed(998244353)
    buf = torch.empty(1, CHUNK, dtype=torch.float32).pin_memory()
    first = torch.cat([torch.zeros(1, 399), 0.1 * torch.randn(1, CHUNK, generator=gen)], 1).pin_memory()
    ost = O.EvictionState()
    sys_n = len(cfg.tpl.system_ids)
    lat, kv_max, n_evict, n_tokens, mismatches = [], 0, 0, 0, 0
    t_all = time.perf_counter()
    for c in range(n_chunks):
        if c:
            buf.copy_(0.1 * torch.randn(1, CHUNK, generator=gen))
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = run.step_api(first if c == 0 else buf)
        torch.cuda.synchronize()
        lat.append(1e3 * (time.perf_counter() - t0))
        n_tokens += len(out[0])
        plan = run.evict_log[-1][0]                     # (keep_prefix, drop_upto, cur) or None
        cur = plan[2] if plan is not None else eng.kv_len(run.sids[0])
        kept = O.evict(ost, cur, g.max_llm_cache_size, g.always_cache_system_prompt, sys_n)
        want = None if (kept is None or kept[1] == cur) else (kept[0], cur - kept[1])
        got = None if plan is None else (plan[0], plan[1])
        mismatches += want != got
        n_evict += plan is not None
        kv_max = max(kv_max, cur)
        run.evict_log.clear()
    wall = time.perf_counter() - t_all
    kv_end = eng.kv_len(run.sids[0])
    enc_steps = eng.enc_steps(run.sids[0])
    run.close()
    leaked = free0 - eng.pages_free()
    line = {"workload": f"BASELINE.json configs[3]: one {args.seconds:.0f} s stream, {n_chunks} chunks of 960 ms, production size, "
                        + ("greedy" if args.beam == 1 else f"beam {args.beam}"),
            "chunks": n_chunks, "wall_s": wall, "speech_s_per_s": n_chunks * 0.96 / wall,
            "latency_ms": {"p50": pct(lat, 0.5), "p99": pct(lat, 0.99), "max": max(lat)},
            "evictions": n_evict, "eviction_mismatches_vs_integer_oracle": mismatches, "kv_len_max": kv_max,
            "kv_len_end": kv_end, "kv_bound": g.max_llm_cache_size + sys_n + 64 + g.max_new_tokens,
            "encoder_frames": enc_steps, "tokens_emitted": n_tokens, "kv_pages_leaked": leaked}
    print(json.dumps(line))
    ok = mismatches == 0 and leaked == 0 and kv_max <= line["kv_bound"] and enc_steps == 48 * n_chunks
    eng.close()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
