"""Diagnostic (not a pytest file): per-step logits of a tiny stream with ISST_DEC_FUSE=1 vs 0 (two processes; the
second run is teacher-forced with the tokens of the first).  Usage: python tests/dec_fuse_diff.py [multiplier]"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
M = int(os.environ.get("DIFF_M", "4"))
if len(sys.argv) > 2:
    import torch
    from infinisst_b200 import tiny_config
    from infinisst_b200.engine import Engine
    from infinisst_b200.synthetic import make_audio, make_state_dict
    from oracle import infinisst_oracle as O
    from parity_utils import bf16_weights, slot_map
    from infinisst_b200.agent import S2TAgentStates, evict_plan
    out_path, forced_path = sys.argv[1], sys.argv[2]
    forced_all = torch.load(forced_path) if os.path.exists(forced_path) else None
    cfg = tiny_config(max_cache_size=192, max_llm_cache_size=300)
    cfg.gen.latency_multiplier, cfg.gen.max_new_tokens = M, 10 * M
    sd = bf16_weights(make_state_dict(cfg, seed=0))
    eng = Engine(cfg, device=0, max_streams=2, max_multiplier=M, max_prompt=64 + 12 * M); eng.load_state_dict(sd); eng.debug(True)
    SEG = 15360 * M
    n = 8
    audio = make_audio(n * SEG / 16000.0)
    sid = eng.open_stream()
    st = S2TAgentStates(); st.system_prompt_size = len(cfg.tpl.system_ids)
    outs, toks_all, target, kvs = [], [], [], []
    for c in range(n):
        pcm = audio[c * SEG:(c + 1) * SEG][None].clone()
        if c == 0: pcm = torch.cat([torch.zeros(1, 399), pcm], 1)
        eng.encode_chunk([sid], pcm, M)
        ids = O.build_prompt(cfg.tpl, c == 0, M)
        kv0 = eng.kv_len(sid)
        toks = eng.generate([sid], [ids], [slot_map(cfg, ids)], [target[-100:]], cfg.gen, pin_prefix=len(cfg.tpl.system_ids),
                            forced=None if forced_all is None else [forced_all[c]])[0]
        toks_all.append(toks)
        kvs.append(kv0 + len(ids))
        outs.append(eng.read_tap("step_logits", torch.float32).view(cfg.gen.max_new_tokens, cfg.llm.vocab).clone())
        target.extend(toks[:-1])
        plan = evict_plan(st, eng.kv_len(sid), cfg.gen.max_llm_cache_size, True)
        if plan is not None: eng.kv_evict(sid, plan[0], plan[1])
    torch.save({"logits": outs, "kv": kvs}, out_path)
    if forced_all is None: torch.save(toks_all, forced_path)
    eng.close()
else:
    import torch
    for f in ("/tmp/dfd_tok.pt",):
        if os.path.exists(f): os.unlink(f)
    for v in ("1", "0"):
        subprocess.run([sys.executable, __file__, f"/tmp/dfd_{v}.pt", "/tmp/dfd_tok.pt"], env=dict(os.environ, ISST_DEC_FUSE=v), check=True)
    a, b = torch.load("/tmp/dfd_1.pt"), torch.load("/tmp/dfd_0.pt")
    toks = torch.load("/tmp/dfd_tok.pt")
    worst = []
    for c in range(len(a["logits"])):
        for s in range(len(toks[c])):
            la, lb = a["logits"][c][s], b["logits"][c][s]
            d = float((la - lb).norm() / lb.norm())
            worst.append((d, c, s, a["kv"][c] + s))
    worst.sort(reverse=True)
    ds = sorted(w[0] for w in worst)
    print("steps", len(ds), "median", ds[len(ds) // 2], "p90", ds[int(0.9 * len(ds))], "max", ds[-1])
    for w in worst[:12]:
        print("  rel diff %.5f  chunk %d step %d  L_old(kv before the step) %d" % w)
