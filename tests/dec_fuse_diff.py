"""Diagnostic (not a pytest file): step logits of a tiny stream with ISST_DEC_FUSE=1 vs 0, in two processes."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
if len(sys.argv) > 1:
    import torch
    from infinisst_b200 import tiny_config
    from infinisst_b200.engine import Engine
    from infinisst_b200.synthetic import make_audio, make_state_dict
    from oracle import infinisst_oracle as O
    from parity_utils import bf16_weights, slot_map
    import numpy as np
    gold = np.load(os.path.join(ROOT, "tests/golden/tiny_stream.npz"))
    cfg = tiny_config(max_cache_size=96, max_llm_cache_size=150)
    sd = bf16_weights(make_state_dict(cfg, seed=0))
    eng = Engine(cfg, device=0, max_streams=2); eng.load_state_dict(sd); eng.debug(True)
    audio = make_audio(8 * 15360 / 16000.0)
    sid = eng.open_stream()
    from infinisst_b200.agent import S2TAgentStates, evict_plan
    st = S2TAgentStates(); st.system_prompt_size = len(cfg.tpl.system_ids)
    outs, target = [], []
    for c in range(8):
        pcm = audio[c * 15360:(c + 1) * 15360][None].clone()
        if c == 0: pcm = torch.cat([torch.zeros(1, 399), pcm], 1)
        eng.encode_chunk([sid], pcm, 1)
        ids = O.build_prompt(cfg.tpl, c == 0)
        forced = gold[f"c{c}_sequence"].tolist()[len(ids):]
        eng.generate([sid], [ids], [slot_map(cfg, ids)], [target[-100:]], cfg.gen, pin_prefix=len(cfg.tpl.system_ids), forced=[forced])
        outs.append(eng.read_tap("step_logits", torch.float32).view(cfg.gen.max_new_tokens, cfg.llm.vocab).clone())
        target.extend(gold[f"c{c}_output_ids"].tolist())
        plan = evict_plan(st, eng.kv_len(sid), cfg.gen.max_llm_cache_size, True)
        if plan is not None: eng.kv_evict(sid, plan[0], plan[1])
    torch.save(torch.stack(outs), sys.argv[1])
    eng.close()
else:
    import torch, numpy as np
    for v in ("1", "0"):
        subprocess.run([sys.executable, __file__, f"/tmp/dec_fuse_{v}.pt"], env=dict(os.environ, ISST_DEC_FUSE=v), check=True)
    a, b = torch.load("/tmp/dec_fuse_1.pt"), torch.load("/tmp/dec_fuse_0.pt")
    gold = np.load(os.path.join(ROOT, "tests/golden/tiny_stream.npz"))
    for c in range(8):
        g = torch.from_numpy(gold[f"c{c}_step_logits"])
        for s in range(10):
            da = float((a[c, s] - g[s]).norm() / g[s].norm()); db = float((b[c, s] - g[s]).norm() / g[s].norm())
            dab = float((a[c, s] - b[c, s]).norm() / b[c, s].norm())
            print(f"chunk {c} step {s}: fused-vs-oracle {da:.4f}  unfused-vs-oracle {db:.4f}  fused-vs-unfused {dab:.5f}")
