"""GPU parity tests: the CUDA path (through the C-ABI) against the oracle on the same seeded
inputs, against the committed golden vectors, and - at production widths - against the oracle
executed in fp32 on the same device.

Tolerances (north_star: "encoder outputs and logits within a stated bf16 tolerance"):
  * encoder taps / speech features: rel-L2 <= 3e-2 against the fp32 oracle;
  * last-position logits:           rel-L2 <= 5e-2 against the fp32 oracle, and never worse than
    2x the error of the oracle itself run in bf16 eager (the reference's own numerics) + 1e-2;
  * greedy tokens: identical under teacher forcing except at near-ties (oracle margin < TIE_EPS);
  * KV lengths and eviction ranges: bit-exact.
"""
import os

import numpy as np
import pytest
import torch

from infinisst_b200 import production_config, tiny_config
from infinisst_b200.synthetic import make_audio, make_state_dict
from oracle import infinisst_oracle as O
from parity_utils import OracleStream, bf16_weights, max_abs, rel_l2, slot_map

pytestmark = pytest.mark.gpu

# the oracle also runs on the GPU as an fp32 checker at production widths: keep it true fp32
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False

ENC_TOL, LOGIT_TOL, TIE_EPS = 3e-2, 5e-2, 0.35
LOGIT_TOL_FULL = 1e-1      # 24 + 32 layers deep: the bf16-eager oracle itself is at 5.5e-2; the yardstick test is the binding one
SEG = 15360
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tiny_stream.npz")


def _engine(cfg, sd, **kw):
    from infinisst_b200.engine import Engine
    eng = Engine(cfg, device=0, **kw)
    eng.load_state_dict(sd)
    return eng


def _chunk_pcm(audio, c):
    pcm = audio[c * SEG:(c + 1) * SEG][None].clone()
    if c == 0:
        pcm = torch.cat([torch.zeros(1, 399), pcm], 1)
    return pcm


# ----------------------------------------------------------------------------------------------
# operators
# ----------------------------------------------------------------------------------------------
GEMM_CASES = [
    (128, 128, 64, {}), (128, 128, 256, {}), (256, 384, 512, {}), (100, 200, 192, {}),
    (3072, 1024, 1024, {}), (1408, 4096, 4096, {}),
    (1, 128, 64, {}), (1, 4096, 4096, {}), (16, 256, 512, {}), (22, 6144, 4096, {}), (48, 3072, 1024, {}),
    (64, 1024, 4096, {}), (12, 519, 512, {"out_f32": True}), (1, 128263, 4096, {"out_f32": True}),
    (64, 128263, 4096, {"out_f32": True}),
    (22, 4096, 4096, {"force_splits": 4}), (1, 4096, 14336, {"force_splits": 8}), (300, 512, 1024, {"force_splits": 3}),
    (48, 1024, 1024, {"bias": True, "resid": True}), (48, 4096, 1024, {"bias": True, "gelu": True}),
    (3072, 4096, 1024, {"bias": True, "gelu": True}), (3072, 1024, 4096, {"bias": True, "resid": True}),
    (22, 768, 512, {"dual": True}), (1, 14336, 4096, {"dual": True}), (1408, 14336, 4096, {"dual": True}),
    (64, 14336, 4096, {"dual": True, "force_splits": 2}), (100, 256, 512, {"force_swap": 1}),
    (30, 256, 512, {"force_swap": 0}),
    # CTA-pair kernel (more than 128 rows): ragged token / feature tiles, stream-K remainder, gate/up with ragged tiles
    (129, 128, 256, {}), (200, 384, 512, {}), (513, 520, 256, {"out_f32": True}), (1408, 6144, 4096, {}),
    (1408, 4096, 14336, {"resid": True}), (4000, 256, 320, {"bias": True}), (300, 768, 512, {"dual": True}),
    (400, 640, 1024, {"dual": True}),
]


@pytest.fixture(scope="module")
def op_engine():
    from infinisst_b200.engine import Engine
    eng = Engine(tiny_config(), device=0, max_streams=2)
    yield eng
    eng.close()


def test_gemm_vs_torch_fp32(op_engine):
    """out = act . W^T with every fused epilogue, against a plain PyTorch fp32 reference of the op.
    Inputs are bf16, accumulation fp32: the only error is the bf16 rounding of the output."""
    torch.manual_seed(0)
    dev = "cuda:0"
    for (M, N, K, kw) in GEMM_CASES:
        a = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
        dual = kw.get("dual", False)
        w = (torch.randn(N * (2 if dual else 1), K, device=dev) * (K ** -0.5)).bfloat16()
        bias = torch.randn(N, device=dev) if kw.get("bias") else None
        resid = torch.randn(M, N, device=dev).bfloat16() if kw.get("resid") else None
        ref = a.float() @ w.float().t()
        if dual:
            ref = torch.nn.functional.silu(ref[:, :N]) * ref[:, N:]
        if bias is not None:
            ref = ref + bias
        if kw.get("gelu"):
            ref = torch.nn.functional.gelu(ref)
        if resid is not None:
            ref = ref + resid.float()
        pair0 = op_engine.path_count("gemm_pair") + op_engine.path_count("gemm_pair_dual")
        out = op_engine.op_gemm(a, w, bias=bias, gelu=kw.get("gelu", False), resid=resid, dual=dual,
                                out_f32=kw.get("out_f32", False), force_swap=kw.get("force_swap", -1),
                                force_splits=kw.get("force_splits", 0))
        torch.cuda.synchronize()
        # above 128 rows a plain GEMM runs on CTA pairs (cta_group::2) unless the test forces the one-CTA kernel
        took_pair = op_engine.path_count("gemm_pair") + op_engine.path_count("gemm_pair_dual") - pair0
        assert took_pair == int(M > 128 and K >= 256 and not kw.get("force_splits") and kw.get("force_swap", -1) != 1), (M, N, K, kw)
        tol = 2e-5 if kw.get("out_f32") else 4e-3            # fp32 out: accumulation order only; bf16 out: 2^-8 rounding
        assert rel_l2(out, ref) < tol, (M, N, K, kw, rel_l2(out, ref))


# ----------------------------------------------------------------------------------------------
# the per-chunk step, tiny config (production head sizes -> production kernels)
# ----------------------------------------------------------------------------------------------
def _run_stream(cfg, n_chunks, check_taps=True, yardstick=False, m=1):
    SEG = 15360 * m                                   # samples per policy call at latency multiplier m
    cfg.gen.latency_multiplier, cfg.gen.max_new_tokens = m, 10 * m
    sd = bf16_weights(make_state_dict(cfg, seed=0))
    eng = _engine(cfg, sd, max_streams=2, max_multiplier=m, max_prompt=64 + 12 * m)
    eng.debug(True)
    audio = make_audio(n_chunks * SEG / 16000.0)
    orc = OracleStream(cfg, sd)
    orc16 = OracleStream(cfg, sd, torch.bfloat16) if yardstick else None
    sid = eng.open_stream()
    target, ck = [], O.EvictionState()
    stats = {"near_ties": 0, "steps": 0, "evictions": 0, "worst_logit": 0.0}
    for c in range(n_chunks):
        out_o, rec, taps = orc.chunk(audio[: (c + 1) * SEG].tolist())
        ids = O.build_prompt(cfg.tpl, c == 0, m)
        forced = rec.sequences[0][len(ids):]
        if yardstick:
            _, rec16, taps16 = orc16.chunk(audio[: (c + 1) * SEG].tolist(), forced=forced)
        pcm = audio[c * SEG:(c + 1) * SEG][None].clone()
        if c == 0:
            pcm = torch.cat([torch.zeros(1, 399), pcm], 1)
        feats = eng.encode_chunk([sid], pcm, m, return_feats=True)
        torch.cuda.synchronize()
        if check_taps:
            for name, okey in [("enc_conv", "conv"), ("enc_post_proj", "post_proj"), ("enc_layer_0", "enc_layer_0"),
                               ("enc_layer_1", "enc_layer_1"), ("enc_out", "enc_out"), ("speech_feats", "speech_feats")]:
                got = eng.read_tap(name).float()
                ref = taps[okey].flatten()
                e = rel_l2(got[: ref.numel()], ref)
                assert e < ENC_TOL, (c, name, e)
                if yardstick:
                    e16 = rel_l2(taps16[okey].flatten(), ref)
                    assert e <= 2 * e16 + 1e-2, (c, name, e, e16)
        assert rel_l2(feats.cpu(), taps["speech_feats"]) < ENC_TOL
        toks = eng.generate([sid], [ids], [slot_map(cfg, ids)], [target[-100:]], cfg.gen,
                            pin_prefix=len(cfg.tpl.system_ids), forced=[forced])[0]
        logits = eng.read_tap("step_logits", torch.float32).view(cfg.gen.max_new_tokens, 1, cfg.llm.vocab)
        assert toks == forced, (c, toks, forced)
        for s in range(len(rec.step_logits)):
            e = rel_l2(logits[s, 0], rec.step_logits[s][0])
            stats["worst_logit"] = max(stats["worst_logit"], e)
            assert e < LOGIT_TOL, (c, s, e)
            if yardstick:
                e16 = rel_l2(rec16.step_logits[s][0], rec.step_logits[s][0])
                assert e <= 2 * e16 + 1e-2, (c, s, e, e16)
            # the token the CUDA path would pick on its own (same processors, applied by the oracle code to
            # the CUDA logits) must be the oracle's token, or the oracle itself must be at a near-tie
            call_ids = ids + forced[:s]
            sc_gpu = O.process_logits(logits[s, 0], call_ids, target[-100:], cfg.gen)
            pick = int(sc_gpu.argmax())
            stats["steps"] += 1
            if pick != forced[s]:
                sc_o = rec.step_scores[s][0]
                assert sc_o[pick] >= sc_o.max() - TIE_EPS, (c, s, pick, forced[s], float(sc_o.max() - sc_o[pick]))
                stats["near_ties"] += 1
        assert eng.kv_len(sid) == orc.st.kv_log[-1]["cur"]
        target.extend(out_o)
        cur = eng.kv_len(sid)
        kept = O.evict(ck, cur, cfg.gen.max_llm_cache_size, True, len(cfg.tpl.system_ids))
        if kept is not None:
            eng.kv_evict(sid, kept[0], cur - kept[1])
            stats["evictions"] += 1
        assert eng.kv_len(sid) == orc.st.llm_cache.length()          # eviction bit-exact
        assert eng.enc_steps(sid) == orc.st.enc_cache.n_steps
    eng.close()
    return stats


def test_stream_vs_oracle_tiny():
    st = _run_stream(tiny_config(), 4, yardstick=True)
    assert st["near_ties"] <= 0.01 * st["steps"] + 1


def test_stream_with_both_windows_sliding():
    """Encoder window 96 frames and LLM window 150 tokens: ring wrap-around, page recycling and
    position shifts after eviction are all exercised within 12 chunks."""
    st = _run_stream(tiny_config(max_cache_size=96, max_llm_cache_size=150), 12, check_taps=True)
    assert st["evictions"] >= 6
    # near-ties (oracle margin < TIE_EPS) flip with the summation order of the kernels: a count, not a correctness bound
    assert st["near_ties"] <= 0.02 * st["steps"] + 1


@pytest.mark.parametrize("m", [2, 4])
def test_latency_multipliers(m):
    """48*m-frame chunks (block size scales with m, speech_encoder.py:143-145), 12*m speech
    tokens per turn and max_new_tokens = 10*m (agents/infinisst.py:125-128,245)."""
    st = _run_stream(tiny_config(max_cache_size=192, max_llm_cache_size=300), 6, check_taps=True, m=m)
    assert st["evictions"] >= 1
    # every flip was already checked to be a near-tie of the oracle itself (margin < TIE_EPS); how many of them flip
    # depends on the kernels' summation order, so the count only guards against a systematic bias
    assert st["near_ties"] <= 0.05 * st["steps"] + 1


def test_long_stream_no_drift():
    """BASELINE.json configs[3] scaled down: 60 chunks, ~55 evictions; the logits error must not grow."""
    st = _run_stream(tiny_config(max_cache_size=96, max_llm_cache_size=150), 60, check_taps=False)
    assert st["evictions"] >= 50 and st["worst_logit"] < LOGIT_TOL


def test_absolute_position_invariance():
    """Keys are rotated once at their absolute index (fp64 angles): a stream whose absolute positions start
    2 000 000 tokens later (hours of speech) must produce the same logits and tokens as a fresh one,
    through prefill, decode and sliding-window eviction with a pinned system prompt."""
    cfg = tiny_config(max_cache_size=96, max_llm_cache_size=150)
    sd = bf16_weights(make_state_dict(cfg, seed=0))
    eng = _engine(cfg, sd, max_streams=2)
    eng.debug(True)
    n_chunks = 10
    audio = make_audio(n_chunks * SEG / 16000.0)
    from infinisst_b200.agent import S2TAgentStates, evict_plan
    runs = []
    for shift in (0, 2_000_000):
        sid = eng.open_stream()
        st = S2TAgentStates()
        st.system_prompt_size = len(cfg.tpl.system_ids)
        target, rec = [], []
        for c in range(n_chunks):
            eng.encode_chunk([sid], _chunk_pcm(audio, c), 1)
            ids = O.build_prompt(cfg.tpl, c == 0)
            if c == 0 and shift:
                eng.debug_shift_positions(sid, shift)        # pinned prefix keys stay at 0..39, everything else moves
            forced = [runs[0][c][0]] if runs else None       # the shifted run is teacher-forced with the fresh run's tokens
            toks = eng.generate([sid], [ids], [slot_map(cfg, ids)], [target[-100:]], cfg.gen,
                                pin_prefix=len(cfg.tpl.system_ids), forced=forced)[0]
            rec.append((toks, eng.read_tap("step_logits", torch.float32).clone()))
            target.extend(toks[:-1])
            plan = evict_plan(st, eng.kv_len(sid), cfg.gen.max_llm_cache_size, True)
            if plan is not None:
                eng.kv_evict(sid, plan[0], plan[1])
        runs.append(rec)
        eng.close_stream(sid)
    V = cfg.llm.vocab
    for c, ((ta, la), (tb, lb)) in enumerate(zip(*runs)):
        assert ta == tb
        la, lb = la.view(-1, V)[: len(ta)], lb.view(-1, V)[: len(ta)]
        # same arithmetic at other absolute angles: only the bf16 rounding of the rotated q / k differs
        assert rel_l2(lb, la) < 1e-2, (c, rel_l2(lb, la))
    eng.close()


def test_golden_stream():
    """Committed golden vectors (tests/golden/make_golden.py): features, logits, tokens, KV log."""
    gold = np.load(GOLD)
    n = int(gold["n_chunks"])
    cfg = tiny_config(max_cache_size=int(gold["max_cache"]), max_llm_cache_size=int(gold["max_llm"]))
    sd = bf16_weights(make_state_dict(cfg, seed=0))
    eng = _engine(cfg, sd, max_streams=2)
    eng.debug(True)
    audio = make_audio(n * SEG / 16000.0)
    sid = eng.open_stream()
    from infinisst_b200.agent import S2TAgentStates, evict_plan
    st = S2TAgentStates()
    st.system_prompt_size = len(cfg.tpl.system_ids)
    target = []
    for c in range(n):
        feats = eng.encode_chunk([sid], _chunk_pcm(audio, c), 1, return_feats=True)
        assert rel_l2(feats[0].cpu(), torch.from_numpy(gold[f"c{c}_speech_feats"])) < ENC_TOL
        ids = O.build_prompt(cfg.tpl, c == 0)
        seq = gold[f"c{c}_sequence"].tolist()
        assert seq[:len(ids)] == ids
        forced = seq[len(ids):]
        toks = eng.generate([sid], [ids], [slot_map(cfg, ids)], [target[-100:]], cfg.gen,
                            pin_prefix=len(cfg.tpl.system_ids), forced=[forced])[0]
        assert toks == forced
        logits = eng.read_tap("step_logits", torch.float32).view(cfg.gen.max_new_tokens, cfg.llm.vocab)
        g = torch.from_numpy(gold[f"c{c}_step_logits"])
        for s in range(g.shape[0]):
            assert rel_l2(logits[s], g[s]) < LOGIT_TOL
        target.extend(gold[f"c{c}_output_ids"].tolist())
        cur, kp, kt, after = gold[f"c{c}_kv"].tolist()
        assert eng.kv_len(sid) == cur
        plan = evict_plan(st, cur, cfg.gen.max_llm_cache_size, True)
        if kp < 0:
            assert plan is None
        else:
            assert plan == (kp, cur - kt)                    # kept = [0, kp) U [cur - kt, cur): bit-exact
            eng.kv_evict(sid, plan[0], plan[1])
        assert eng.kv_len(sid) == after
    eng.close()


def test_batched_streams_equal_independent_streams():
    """BASELINE.json configs[2] scaled down: 5 distinct streams in lock-step through one batched call
    each chunk == the same streams run one by one (oracle = loop of B=1 oracles)."""
    from infinisst_b200.engine import Engine
    from infinisst_b200.runner import LockstepRunner
    cfg = tiny_config(max_cache_size=96, max_llm_cache_size=150)
    sd = bf16_weights(make_state_dict(cfg, seed=0))
    B, n_chunks = 5, 7
    audios = [make_audio(n_chunks * SEG / 16000.0, seed=100 + b) for b in range(B)]
    eng = _engine(cfg, sd, max_streams=B)
    eng.debug(True)
    run = LockstepRunner(eng, cfg, B)
    orcs = [OracleStream(cfg, sd) for _ in range(B)]
    for c in range(n_chunks):
        recs = [o.chunk(a[: (c + 1) * SEG].tolist()) for o, a in zip(orcs, audios)]
        ids = O.build_prompt(cfg.tpl, c == 0)
        forced = [r[1].sequences[0][len(ids):] for r in recs]
        pcm = torch.cat([_chunk_pcm(a, c) for a in audios], 0)
        outs = run.step_device(pcm, forced=forced)
        logits = eng.read_tap("step_logits", torch.float32).view(cfg.gen.max_new_tokens, B, cfg.llm.vocab)
        for b in range(B):
            assert run.last_tokens[b] == forced[b]
            assert outs[b] == recs[b][0]
            for s in range(len(recs[b][1].step_logits)):
                assert rel_l2(logits[s, b], recs[b][1].step_logits[s][0]) < LOGIT_TOL, (c, b, s)
            assert eng.kv_len(run.sids[b]) == orcs[b].st.llm_cache.length()
            log = orcs[b].st.kv_log[-1]
            mine = run.evict_log[-1][b]
            if log["kept"] is None:
                assert mine is None
            else:
                assert mine == (log["kept"][0], log["cur"] - log["kept"][1], log["cur"])
    assert run.evictions >= B
    run.close()
    eng.close()


def test_agent_drop_in_api():
    """The SimulEval-facing agent (same class/method names as agents/infinisst.py): audio arrives as
    growing Python lists on `states.source`; tokens equal the oracle's free-running greedy tokens up
    to the first near-tie; KV lengths follow the integer eviction model."""
    import argparse
    from infinisst_b200.agent import InfiniSST
    cfg = tiny_config(max_cache_size=96, max_llm_cache_size=150)
    sd = bf16_weights(make_state_dict(cfg, seed=0))
    p = argparse.ArgumentParser()
    InfiniSST.add_args(p)
    args = p.parse_args(["--w2v2-type", "w2v2", "--block-size", "48", "--max-cache-size", "96", "--xpos", "0",
                         "--latency-multiplier", "1", "--max-latency-multiplier", "1", "--max-new-tokens", "10",
                         "--no-repeat-ngram-size", "5", "--max-llm-cache-size", "150", "--always-cache-system-prompt",
                         "--beam", "1"])
    args.model_config, args.state_dict = cfg, sd
    agent = InfiniSST(args)
    states = agent.build_states()
    states.source_sample_rate = 16000
    n_chunks = 8
    audio = make_audio(n_chunks * SEG / 16000.0)
    orc = OracleStream(cfg, sd)
    agree = True
    for c in range(n_chunks):
        states.source = audio[: (c + 1) * SEG].tolist()
        states.source_finished = c == n_chunks - 1
        n_before = len(states.target_ids)
        act = agent.policy(states)
        out_o, rec, _ = orc.chunk(audio[: (c + 1) * SEG].tolist())
        assert not act.is_read()
        if agree and states.target_ids[n_before:] != out_o:
            agree = False                                  # a near-tie flipped: later chunks legitimately differ
            margins = [float(s[0].max() - s[0].topk(2).values[1]) for s in rec.step_scores]
            assert min(margins) < TIE_EPS, (c, margins)
        if agree:
            assert states.past_key_values[0][0].size(2) == orc.st.llm_cache.length()
    assert states.past_key_values[0][0].size(2) <= 150 + 40
    assert len(agent.chunk_latencies) == n_chunks
    assert act.finished
    states.reset()
    agent.model.engine.close()


@pytest.mark.parametrize("flags", [["--latency-multiplier", "2", "--max-new-tokens", "20"],
                                   ["--latency-multiplier", "4", "--max-new-tokens", "40"],
                                   ["--latency-multiplier", "3"],
                                   ["--latency-multiplier", "2", "--max-new-tokens", "20", "--pseudo-batch-size", "3"]],
                         ids=["m2", "m4", "m3_default_max_new_and_cache", "m2_pseudo_batch3"])
def test_agent_sizes_its_engine_from_the_flags(flags):
    """The agent's own flags size the engine (first-chunk prompt = 40 + 9 + 12 m tokens: 73 / 85 / 97 at m = 2 / 3 / 4;
    `--max-new-tokens` defaults to 1000 and `--max-llm-cache-size` to 10000, agents/infinisst.py:185-198 +
    agents/options.py).  m = 2 is held against the oracle (tokens up to the first near-tie, KV lengths); the default
    flags must simply run: 1000 new tokens per call, a 10000-token window that never evicts here."""
    import argparse
    from infinisst_b200.agent import InfiniSST
    cfg = tiny_config(max_cache_size=192, max_llm_cache_size=300)
    sd = bf16_weights(make_state_dict(cfg, seed=0))
    p = argparse.ArgumentParser()
    InfiniSST.add_args(p)
    base = ["--w2v2-type", "w2v2", "--block-size", "48", "--max-cache-size", "192", "--xpos", "0", "--no-repeat-ngram-size", "5",
            "--always-cache-system-prompt", "--beam", "1"]
    defaults = "--max-new-tokens" not in flags
    if not defaults:
        base += ["--max-llm-cache-size", "300"]
    args = p.parse_args(base + flags)
    args.model_config, args.state_dict, args.log_chunks = cfg, sd, False
    agent = InfiniSST(args)
    m = args.latency_multiplier
    states = agent.build_states()
    states.source_sample_rate = 16000
    n_calls = 2 if defaults else 5
    audio = make_audio(n_calls * m * SEG / 16000.0)
    cfg_o = tiny_config(max_cache_size=192, max_llm_cache_size=300)
    cfg_o.gen.latency_multiplier, cfg_o.gen.max_new_tokens = m, args.max_new_tokens
    orc = OracleStream(cfg_o, sd)
    agree = not defaults
    for c in range(n_calls):
        states.source = audio[: (c + 1) * m * SEG].tolist()
        states.source_finished = c == n_calls - 1
        n_before = len(states.target_ids)
        act = agent.policy(states)
        assert not act.is_read()
        kv = states.past_key_values[0][0].size(2)
        if defaults:
            assert len(states.target_ids) - n_before <= 999 and kv <= 10000 + 40
            continue
        out_o, rec, _ = orc.chunk(audio[: (c + 1) * m * SEG].tolist())
        if agree and states.target_ids[n_before:] != out_o:
            agree = False
            margins = [float(s[0].max() - s[0].topk(2).values[1]) for s in rec.step_scores]
            assert min(margins) < TIE_EPS, (c, margins)
        if agree:
            assert kv == orc.st.llm_cache.length()
        assert kv <= 300 + 40 + 22 + 12 * m + 10 * m
    assert act.finished and states.segment_idx == 0          # -1 on the last segment (agents/infinisst.py:303-304), then += 1
    states.reset()
    agent.model.engine.close()


# ----------------------------------------------------------------------------------------------
# production widths
# ----------------------------------------------------------------------------------------------
def _production(enc_layers, llm_layers):
    cfg = production_config()
    cfg.enc.layers, cfg.llm.layers = enc_layers, llm_layers
    return cfg


def _oracle_on_gpu_stream(cfg, sd_dev):
    """The oracle is plain PyTorch: at production widths it runs in fp32 on the GPU as the checker."""
    class _S:
        pass
    s = _S()
    s.cfg, s.sd, s.st = cfg, sd_dev, O.StreamState()
    return s


@pytest.mark.parametrize("layers", [(2, 2), (24, 32)], ids=["slice_2+2_layers", "full_24+32_layers"])
def test_production_widths_vs_oracle(layers):
    """wav2vec2-large widths (512-ch extractor, d=1024, ffn 4096, 16x64 heads) and Llama-3.1-8B widths
    (4096, 32/8x128, ffn 14336, vocab 128263); `full` is the whole BASELINE.json configs[1] model."""
    cfg = _production(*layers)
    dev = "cuda:0"
    sd16 = make_state_dict(cfg, seed=0, device=dev, dtype=torch.bfloat16)
    eng = _engine(cfg, sd16, max_streams=2)
    eng.debug(True)
    sd32 = {k: v.float() for k, v in sd16.items()}
    n_chunks = 3
    audio = make_audio(n_chunks * SEG / 16000.0)
    st = O.StreamState()
    st16 = O.StreamState()      # the oracle in bf16 eager on the GPU = the reference's own numerics (yardstick)
    sid = eng.open_stream()
    target = []
    worst = {"feat": 0.0, "logit": 0.0}
    flips = flips16 = steps = 0
    with torch.inference_mode():
        for c in range(n_chunks):
            taps = {}
            out_o, rec = O.policy_chunk(sd32, cfg, st, audio[: (c + 1) * SEG].tolist(), torch.float32, taps)
            ids = O.build_prompt(cfg.tpl, c == 0)
            forced = rec.sequences[0][len(ids):]
            taps16 = {}
            _, rec16 = O.policy_chunk(sd16, cfg, st16, audio[: (c + 1) * SEG].tolist(), torch.bfloat16, taps16, forced)
            feats = eng.encode_chunk([sid], _chunk_pcm(audio, c), 1, return_feats=True)
            e = rel_l2(feats, taps["speech_feats"])
            e16 = rel_l2(taps16["speech_feats"], taps["speech_feats"])
            worst["feat"] = max(worst["feat"], e)
            worst["feat16"] = max(worst.get("feat16", 0.0), e16)
            assert e < ENC_TOL and e <= 2 * e16 + 1e-2, (c, e, e16)
            toks = eng.generate([sid], [ids], [slot_map(cfg, ids)], [target[-100:]], cfg.gen,
                                pin_prefix=len(cfg.tpl.system_ids), forced=[forced])[0]
            assert toks == forced
            logits = eng.read_tap("step_logits", torch.float32).view(cfg.gen.max_new_tokens, cfg.llm.vocab)
            for s in range(len(rec.step_logits)):
                ref = rec.step_logits[s][0].cpu()
                e = rel_l2(logits[s], ref)
                e16 = rel_l2(rec16.step_logits[s][0].cpu(), ref)
                worst["logit"] = max(worst["logit"], e)
                worst["logit16"] = max(worst.get("logit16", 0.0), e16)
                assert e < (LOGIT_TOL_FULL if layers[1] > 8 else LOGIT_TOL) and e <= 2 * e16 + 1e-2, (c, s, e, e16)
                sc = O.process_logits(logits[s], ids + forced[:s], target[-100:], cfg.gen)
                steps += 1
                sc16 = O.process_logits(rec16.step_logits[s][0].cpu(), ids + forced[:s], target[-100:], cfg.gen)
                flips16 += int(sc16.argmax()) != forced[s]
                if int(sc.argmax()) != forced[s]:
                    so = rec.step_scores[s][0].cpu()
                    # a flip must be explained by the measured numerical noise: oracle margin below 4x the rms logit error
                    err_rms = float((logits[s] - ref).pow(2).mean().sqrt())
                    assert so[int(sc.argmax())] >= so.max() - max(TIE_EPS, 4.0 * err_rms), (c, s, err_rms)
                    flips += 1
            assert eng.kv_len(sid) == st.llm_cache.length()
            target.extend(out_o)
    print(f"production {layers}: worst feat rel_l2 {worst['feat']:.3e} (bf16-eager oracle {worst['feat16']:.3e}), "
          f"worst logit rel_l2 {worst['logit']:.3e} (bf16-eager oracle {worst['logit16']:.3e}), near-tie flips {flips}/{steps} "
          f"(bf16-eager oracle vs fp32 oracle: {flips16}/{steps})")
    # 128 263 near-iid random logits: the top-2 gap is below the bf16 error for a few percent of the steps
    # (SURVEY §7 hard part 2); every flip was checked above to be such a near-tie, and the CUDA path must not
    # flip more often than the reference's own bf16-eager numerics do against the fp32 oracle (measured: 1 / 30 and
    # 2 / 30 against 1 / 30 and 6 / 30; token-level agreement over long streams: tests/test_gpu_headline.py)
    assert flips <= max(flips16 + 1, 3)
    eng.close()


def test_decode_attention_bench_runs(op_engine):
    ms = op_engine.decode_attention_bench(2, 500, 4)
    assert ms > 0


def test_profile_counters(op_engine):
    """The roofline leg of bench.py: per-class launches and algorithmic work are recorded."""
    cfg = tiny_config()
    sd = bf16_weights(make_state_dict(cfg, seed=0))
    eng = _engine(cfg, sd, max_streams=2)
    eng.profile(True)
    eng.profile_reset()
    sid = eng.open_stream()
    audio = make_audio(SEG / 16000.0)
    eng.encode_chunk([sid], _chunk_pcm(audio, 0), 1)
    ids = O.build_prompt(cfg.tpl, True)
    eng.generate([sid], [ids], [slot_map(cfg, ids)], [[]], cfg.gen, pin_prefix=40)
    prof = eng.profile_read()
    assert prof["gemm_stream"]["launches"] > 0 and prof["gemm_stream"]["bytes"] > 0
    assert prof["attn_decode"]["launches"] == cfg.llm.layers * (cfg.gen.max_new_tokens - 1)
    assert prof["attn_encoder"]["launches"] == cfg.enc.layers and prof["attn_prefill"]["launches"] == cfg.llm.layers
    assert all(v["ms"] >= 0 for v in prof.values())
    eng.close()


def test_streams_join_a_running_batch():
    """Continuous serving: streams arrive at different times and share the batched calls from then on - a joining
    stream's first chunk (61-token system + turn prompt, 79+320 zero offset = the library's zero carried tail) runs
    in the same batch as 22-token later turns of the others (right-padded ids + attention_mask through
    `model.generate`).  Every stream is checked against ITS OWN oracle: teacher-forced tokens, step logits, KV
    lengths, evictions - exactly as if it ran alone."""
    from infinisst_b200.agent import S2TAgentStates, evict_plan
    from infinisst_b200.model import SpeechLlamaForCausalLM
    cfg = tiny_config(max_cache_size=96, max_llm_cache_size=150)
    sd = bf16_weights(make_state_dict(cfg, seed=0))
    eng = _engine(cfg, sd, max_streams=4)
    eng.debug(True)
    model = SpeechLlamaForCausalLM(cfg, engine=eng)
    g = cfg.gen
    n_calls, joins = 8, [0, 1, 3]
    audios = [make_audio(n_calls * SEG / 16000.0, seed=100 + j) for j in range(len(joins))]
    orcs = [OracleStream(cfg, sd) for _ in joins]
    states = [S2TAgentStates() for _ in joins]
    free0 = eng.pages_free()
    evictions = 0
    for call in range(n_calls):
        live = [j for j, start in enumerate(joins) if call >= start]
        recs, rows, forced = {}, [], []
        for j in live:
            c = call - joins[j]
            recs[j] = orcs[j].chunk(audios[j][: (c + 1) * SEG].tolist())
            ids = O.build_prompt(cfg.tpl, c == 0)
            rows.append(ids)
            forced.append(recs[j][1].sequences[0][len(ids):])
            if c == 0:
                states[j].system_prompt_size = len(cfg.tpl.system_ids)
        width = max(len(r) for r in rows)
        ids_t = torch.full((len(live), width), g.pad_token_id, dtype=torch.long)
        mask = torch.zeros(len(live), width, dtype=torch.long)
        for k, r in enumerate(rows):
            ids_t[k, :len(r)] = torch.tensor(r)
            mask[k, :len(r)] = 1
        # every row brings exactly one chunk of samples: the joining stream's zero offset is its zero carried tail
        pcm = torch.cat([audios[j][(call - joins[j]) * SEG:(call - joins[j] + 1) * SEG][None] for j in live], 0)
        out = model.generate(attention_mask=mask, input_ids=ids_t, speech_batch=pcm, num_beams=1,
                             max_new_tokens=g.max_new_tokens, encoder_input_ids=[states[j].target_ids[-100:] for j in live],
                             encoder_no_repeat_ngram_size=g.no_repeat_ngram_size, no_repeat_ngram_size=g.no_repeat_ngram_size,
                             repetition_penalty=g.repetition_penalty, pad_token_id=g.pad_token_id,
                             states=[states[j] for j in live], multiplier=1, forced_tokens=forced,
                             pin_prefix=len(cfg.tpl.system_ids))
        logits = eng.read_tap("step_logits", torch.float32).view(g.max_new_tokens, len(live), cfg.llm.vocab)
        for k, j in enumerate(live):
            rec = recs[j][1]
            assert out.generated[k] == forced[k], (call, j)
            assert out.sequences[k, :len(rows[k]) + len(forced[k])].tolist() == rows[k] + forced[k]
            for s_ in range(len(rec.step_logits)):
                assert rel_l2(logits[s_, k], rec.step_logits[s_][0]) < LOGIT_TOL, (call, j, s_)
            st = states[j]
            st.target_ids.extend(recs[j][0])
            st.past_key_values = st.speech_cache
            cur = st.speech_cache.kv_len
            plan = evict_plan(st, cur, g.max_llm_cache_size, g.always_cache_system_prompt)
            if plan is not None:
                eng.kv_evict(st.speech_cache.sid, plan[0], plan[1])
                evictions += 1
            log = orcs[j].st.kv_log[-1]
            assert cur == log["cur"] and st.speech_cache.kv_len == log["after"], (call, j)
            assert st.speech_cache.n_steps == orcs[j].st.enc_cache.n_steps
    assert evictions >= 4
    for st in states:
        st.speech_cache.close()
    assert eng.pages_free() == free0
    eng.close()


@pytest.mark.parametrize("beam", [1, 4], ids=["greedy", "beam4"])
def test_agent_policy_batch_accepts_joining_streams(beam):
    """The agent-level form of the above: `policy_batch` with streams that start at different calls returns one
    action per stream, keeps every KV window bounded and gives every page back (greedy and the shipped `--beam 4`)."""
    import argparse
    from infinisst_b200.agent import InfiniSST
    cfg = tiny_config(max_cache_size=96, max_llm_cache_size=150)
    sd = bf16_weights(make_state_dict(cfg, seed=0))
    p = argparse.ArgumentParser()
    InfiniSST.add_args(p)
    args = p.parse_args(["--w2v2-type", "w2v2", "--block-size", "48", "--max-cache-size", "96", "--xpos", "0",
                         "--latency-multiplier", "1", "--max-latency-multiplier", "1", "--max-new-tokens", "10",
                         "--no-repeat-ngram-size", "5", "--max-llm-cache-size", "150", "--always-cache-system-prompt",
                         "--beam", str(beam)])
    args.model_config, args.state_dict, args.max_streams = cfg, sd, 4
    agent = InfiniSST(args)
    free0 = agent.model.engine.pages_free()
    n_calls, joins = 8, [0, 1, 3]
    audios = [make_audio(n_calls * SEG / 16000.0, seed=100 + j) for j in range(len(joins))]
    states = [agent.build_states() for _ in joins]
    solo_first = []
    for j in range(len(joins)):          # each stream's first chunk alone (the explicit-offset form of the reference)
        st = agent.build_states()
        st.source_sample_rate = 16000
        st.source = audios[j][:SEG].tolist()
        agent.policy(st)
        solo_first.append(list(st.target_ids))
        st.reset()
    for st in states:
        st.source_sample_rate = 16000
    for call in range(n_calls):
        live = [j for j, start in enumerate(joins) if call >= start]
        for j in live:
            states[j].source = audios[j][: (call - joins[j] + 1) * SEG].tolist()
            states[j].source_finished = call == n_calls - 1
        acts = agent.policy_batch([states[j] for j in live])
        assert len(acts) == len(live) and all(a is not None for a in acts)
        for j in live:
            if call == joins[j] and j > 0 and beam == 1:
                # joined a running batch: same emitted ids as alone up to the first near-tie (random weights; the fp32
                # sums of the attention key splits depend on the batch).  Not asserted for beam search, where one flip
                # among near-equal hypotheses replaces the whole sequence (tests/test_gpu_beam.py checks those by score)
                a, b = states[j].target_ids, solo_first[j]
                same = next((i for i, (x, y) in enumerate(zip(a, b)) if x != y), min(len(a), len(b)))
                assert same >= min(len(a), len(b)) - 3, (j, a, b)
            assert states[j].past_key_values[0][0].size(2) <= 150 + 40
    assert all(len(st.target_ids) > 20 for st in states)
    for st in states:
        st.reset()
    assert agent.model.engine.pages_free() == free0
    agent.model.engine.close()


def test_model_forward_returns_every_position_like_the_reference():
    """`model.forward` (model/llm.py:192-270) applies lm_head to all T positions (:236-237): logits are
    [B, T, vocab].  Two calls on one stream (first chunk with the spliced speech tokens, then a follow-up turn over
    the cached KV), every position held against the oracle; `last_only=True` equals the last row."""
    from infinisst_b200.model import SpeechLlamaForCausalLM
    cfg = tiny_config()
    sd = bf16_weights(make_state_dict(cfg, seed=0))
    eng = _engine(cfg, sd, max_streams=3)
    model = SpeechLlamaForCausalLM(cfg, engine=eng)
    osd = O.cast_state_dict(sd, torch.float32)
    audio = make_audio(2 * SEG / 16000.0)

    class _St:
        speech_cache = None
    st_a, st_b = _St(), _St()
    enc_cache, llm_cache = None, O.LlmCache.empty(cfg.llm.layers)
    for c in range(2):
        ids = O.build_prompt(cfg.tpl, c == 0)
        pcm = _chunk_pcm(audio, c)
        feats, enc_cache = O.encode_speech(osd, cfg.enc, pcm, enc_cache)
        emb = O.splice_embeddings(osd, cfg.llm, torch.tensor([ids]), feats)
        ref = O.llama_forward(osd, cfg.llm, emb, llm_cache)[0]                       # [T, V]
        ids_t = torch.tensor([ids], dtype=torch.long)
        out = model.forward(input_ids=ids_t, speech_batch=pcm, states=st_a, multiplier=1,
                            pin_prefix=len(cfg.tpl.system_ids))
        assert tuple(out.logits.shape) == (1, len(ids), cfg.llm.vocab)
        got = out.logits[0].float().cpu()
        worst = max(rel_l2(got[t], ref[t]) for t in range(len(ids)))
        assert worst < LOGIT_TOL, (c, worst)
        last = model.forward(input_ids=ids_t, speech_batch=pcm, states=st_b, multiplier=1,
                             pin_prefix=len(cfg.tpl.system_ids), last_only=True)
        assert tuple(last.logits.shape) == (1, 1, cfg.llm.vocab)
        assert rel_l2(last.logits[0, 0].float().cpu(), got[-1]) < 1e-3
        assert out.past_key_values[0][0].size(2) == llm_cache.length()
    eng.close()


def test_error_convention_and_recovery():
    """SURVEY §8b error convention: every misuse comes back as a non-zero status + isst_last_error() (raised as
    IsstError by the shim), nothing is silently truncated, and the context keeps working afterwards."""
    from infinisst_b200._lib import IsstError
    cfg = tiny_config(max_cache_size=96, max_llm_cache_size=150)
    sd = bf16_weights(make_state_dict(cfg, seed=0))
    eng = _engine(cfg, sd, max_streams=2)
    a, b = eng.open_stream(), eng.open_stream()
    with pytest.raises(IsstError, match="no free stream slot"):
        eng.open_stream()
    audio = make_audio(2 * SEG / 16000.0)
    with pytest.raises(IsstError, match="n_samples"):
        eng.encode_chunk([a], audio[None, :1000].clone(), 1)                  # not a whole chunk
    with pytest.raises(IsstError, match="multiplier"):
        eng.encode_chunk([a], audio[None, :2 * SEG].clone(), 2)              # engine built for multiplier 1
    with pytest.raises(IsstError):
        eng.encode_chunk([a, a], torch.cat([_chunk_pcm(audio, 0)] * 2, 0), 1)  # the same stream twice in a batch
    with pytest.raises(IsstError):
        eng.kv_len(7)                                                          # not an open stream
    ids = O.build_prompt(cfg.tpl, True)
    with pytest.raises(IsstError, match="prompt length"):
        eng.generate([a], [ids * 4], [[-1] * (4 * len(ids))], [[]], cfg.gen, pin_prefix=len(cfg.tpl.system_ids))
    # the context is intact: a normal chunk on both streams still matches the oracle
    orc = OracleStream(cfg, sd)
    out_o, rec, taps = orc.chunk(audio[:SEG].tolist())
    forced = rec.sequences[0][len(ids):]
    pcm = torch.cat([_chunk_pcm(audio, 0)] * 2, 0)
    eng.encode_chunk([a, b], pcm, 1)
    toks = eng.generate([a, b], [ids, ids], [slot_map(cfg, ids)] * 2, [[], []], cfg.gen,
                        pin_prefix=len(cfg.tpl.system_ids), forced=[forced, forced])
    assert toks[0] == forced and toks[1] == forced
    assert eng.kv_len(a) == eng.kv_len(b) == orc.st.llm_cache.length()
    with pytest.raises(IsstError, match="keep_prefix"):
        eng.kv_evict(a, 3, 50)                                                 # not the stream's pinned prefix
    cur = eng.kv_len(a)
    with pytest.raises(IsstError):
        eng.kv_evict(a, len(cfg.tpl.system_ids), cur + 5)                      # beyond the cache
    assert eng.kv_len(a) == cur
    eng.kv_evict(a, len(cfg.tpl.system_ids), len(cfg.tpl.system_ids) + 10)
    assert eng.kv_len(a) == cur - 10
    eng.close_stream(a)
    c = eng.open_stream()                                                      # the slot is reusable
    assert eng.kv_len(c) == 0 and eng.enc_steps(c) == 0
    eng.close()


def test_update_multiplier_mid_stream():
    """`InfiniSST.update_multiplier` (agents/infinisst.py:125-128) may change the latency multiplier between policy
    calls of one stream: chunk length, block size of the encoder mask (set_blocksize, speech_encoder.py:143-145),
    speech tokens per turn and max_new_tokens all follow.  One stream through m = 1, 2, 2, 1, 4, 1, 2 against the
    oracle (teacher-forced tokens, logits, features, KV lengths and evictions)."""
    cfg = tiny_config(max_cache_size=192, max_llm_cache_size=300)
    sd = bf16_weights(make_state_dict(cfg, seed=0))
    pattern = [1, 2, 2, 1, 4, 1, 2]
    eng = _engine(cfg, sd, max_streams=2, max_multiplier=4, max_prompt=64 + 12 * 4)
    eng.debug(True)
    audio = make_audio(sum(pattern) * SEG / 16000.0)
    orc = OracleStream(cfg, sd)
    sid = eng.open_stream()
    target, ck, pos, evictions = [], O.EvictionState(), 0, 0
    for c, m in enumerate(pattern):
        cfg.gen.latency_multiplier, cfg.gen.max_new_tokens = m, 10 * m
        n_new = SEG * m
        out_o, rec, taps = orc.chunk(audio[: pos + n_new].tolist())
        ids = O.build_prompt(cfg.tpl, c == 0, m)
        forced = rec.sequences[0][len(ids):]
        pcm = audio[pos:pos + n_new][None].clone()
        if c == 0:
            pcm = torch.cat([torch.zeros(1, 399), pcm], 1)
        pos += n_new
        feats = eng.encode_chunk([sid], pcm, m, return_feats=True)
        assert feats.shape[1] == 12 * m
        assert rel_l2(feats.cpu(), taps["speech_feats"]) < ENC_TOL, (c, m)
        toks = eng.generate([sid], [ids], [slot_map(cfg, ids)], [target[-100:]], cfg.gen,
                            pin_prefix=len(cfg.tpl.system_ids), forced=[forced])[0]
        assert toks == forced, (c, m)
        logits = eng.read_tap("step_logits", torch.float32).view(-1, cfg.llm.vocab)
        for s in range(len(rec.step_logits)):
            assert rel_l2(logits[s], rec.step_logits[s][0]) < LOGIT_TOL, (c, m, s)
        target.extend(out_o)
        cur = eng.kv_len(sid)
        assert cur == orc.st.kv_log[-1]["cur"]
        kept = O.evict(ck, cur, cfg.gen.max_llm_cache_size, True, len(cfg.tpl.system_ids))
        if kept is not None:
            eng.kv_evict(sid, kept[0], cur - kept[1])
            evictions += 1
        assert eng.kv_len(sid) == orc.st.llm_cache.length()
        assert eng.enc_steps(sid) == orc.st.enc_cache.n_steps
    assert evictions >= 1
    eng.close()


@pytest.mark.parametrize("shape", ["tiny_1", "tiny_5", "tiny_100", "tiny_200", "prod_slice_64", "prod_slice_20", "prod_slice_130"])
def test_decode_chain_bit_identical_to_operator_path(shape):
    """The fused decode-layer chain (decode_chain.cuh: o_proj -> RMSNorm -> gate/up -> down -> RMSNorm -> next QKV in
    one persistent kernel with grid barriers) keeps the k-split ranges, the accumulation order and the partial-sum
    order of the operator-per-kernel path: with the RMSNorms as row phases (`chain_fold` = 0) the raw step logits of both
    paths must be IDENTICAL bit for bit, for 1, 5, 20 and 64 rows (token tiles of 16, 32 and 64 columns), tiny and
    production widths; the product default (`chain_fold` = 1: norms folded into the decode GEMMs) must stay within bf16
    noise of them."""
    from infinisst_b200.runner import LockstepRunner
    if shape.startswith("tiny"):
        cfg = tiny_config(max_cache_size=96, max_llm_cache_size=150)
        sd = bf16_weights(make_state_dict(cfg, seed=0))
        B = int(shape.split("_")[1])
        n_chunks = 7 if B <= 16 else 3
    else:
        cfg = _production(1, 3)
        sd = make_state_dict(cfg, seed=0, device="cuda:0", dtype=torch.bfloat16)
        B, n_chunks = int(shape.split("_")[2]), 3
    audios = [make_audio(n_chunks * SEG / 16000.0, seed=300 + b) for b in range(B)]
    res = []
    # (chain, folded norms): the product default, the chain with RMSNorm row phases, the operator-per-kernel path
    for use_chain, fold in ((1, 1), (1, 0), (0, 0)):
        eng = _engine(cfg, sd, max_streams=B, max_batch=B)
        eng.option("decode_chain", use_chain)
        eng.option("chain_fold", fold)
        eng.option("defer_splits_as_chain", 1)                   # operator path: the chain's k-ranges (its units are tile pairs)
        eng.debug(True)
        run = LockstepRunner(eng, cfg, B)
        rec = []
        for c in range(n_chunks):
            pcm = torch.cat([_chunk_pcm(a, c) for a in audios], 0)
            forced = res[0][c][0] if res else None               # later runs are teacher-forced with the first run's tokens
            run.step_device(pcm, forced=forced)
            rec.append((run.last_tokens, eng.read_tap("step_logits", torch.float32).clone()))
        bn = next(x for x in (16, 32, 64, 128, 256) if B <= x)
        n_chain = eng.path_count(f"decode_chain{bn}")
        assert (n_chain > 0) == bool(use_chain), (use_chain, n_chain)
        if use_chain:
            assert n_chain == n_chunks * (cfg.gen.max_new_tokens - 1) * (cfg.llm.layers + 1)
        res.append(rec)
        run.close()
        eng.close()
    for c, ((tf, lf), (ta, la), (tb, lb)) in enumerate(zip(*res)):
        assert tf == ta == tb
        if B <= 64:
            assert torch.equal(la, lb), (c, float((la - lb).abs().max()))
        else:
            # beyond 64 rows the operator path puts tokens on the 128-lane operand (stream-K, other k-ranges): the fp32
            # sums differ in the last bits, a bf16 rounding flips here and there and propagates through the layers - equal
            # to well below the bf16 tolerance of the parity tests (5e-2 against the fp32 oracle; measured here: 7e-3)
            assert rel_l2(la, lb) < 1.5e-2, (c, rel_l2(la, lb))
        # folded norms: the normalised activations are rounded to bf16 once (x * w) instead of twice (x / rms, then * w), so
        # the logits move by bf16 noise (as much as between any two bf16 implementations)
        assert rel_l2(lf, lb) < 1.5e-2, (c, rel_l2(lf, lb))
        print(f"chunk {c}: folded vs operator path rel_l2 {rel_l2(lf, lb):.2e}")
