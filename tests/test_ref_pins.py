"""Pins produced by EXECUTING THE REFERENCE'S OWN CODE (tests/golden/make_ref_pins.py runs
/root/reference's agent, model, encoder patch and LLM attention patch on the tiny configuration with
stand-ins only for the absent third-party packages; tests/golden/ref_standins.py lists which is which).

CPU (`-m "not gpu"`): the oracle must reproduce the reference's numbers - features, per-step logits, token
sequences, KV lengths, kept KV index sets, eviction timelines at production parameters, attention masks.
GPU (`-m gpu`): the CUDA path (through the C-ABI) must reproduce them within the stated bf16 tolerance,
integers bit-exact.
"""
import os

import numpy as np
import pytest
import torch

from infinisst_b200 import tiny_config
from infinisst_b200.synthetic import make_audio, make_state_dict
from oracle import infinisst_oracle as O
from parity_utils import OracleStream, bf16_weights, rel_l2, slot_map

PINS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_tiny_stream.npz")
SEG = 15360
SCENARIOS = ["prod", "prod_ragged", "nosys", "tight", "overflow"]


@pytest.fixture(scope="module")
def pins():
    return np.load(PINS)


def _runs_of(pins, name, n_chunks):
    runs = [[] for _ in range(n_chunks)]
    for ci, a, b in pins[f"evict_{name}_runs"].tolist():
        runs[ci].append((a, b))
    return runs


def test_oracle_reproduces_reference_stream(pins):
    """Encoder features, logits, greedy tokens, KV lengths and kept index sets of the reference agent."""
    n = int(pins["n_chunks"])
    cfg = tiny_config(max_cache_size=int(pins["max_cache"]), max_llm_cache_size=int(pins["max_llm"]))
    sd = bf16_weights(make_state_dict(cfg, seed=0))
    audio = make_audio(n * SEG / 16000.0)
    orc = OracleStream(cfg, sd)
    evictions = 0
    for c in range(n):
        out_ids, rec, taps = orc.chunk(audio[: (c + 1) * SEG].tolist())
        # same fp32 operator sequence: features agree to rounding, logits to accumulation order
        np.testing.assert_allclose(taps["speech_feats"][0].numpy(), pins[f"c{c}_speech_feats"], atol=1e-5, rtol=1e-5)
        got = torch.stack([l[0] for l in rec.step_logits]).numpy()
        np.testing.assert_allclose(got, pins[f"c{c}_step_logits"], atol=2e-4, rtol=1e-4)
        assert rec.sequences[0] == pins[f"c{c}_sequence"].tolist()
        assert out_ids == pins[f"c{c}_output_ids"].tolist()
        log = orc.st.kv_log[-1]
        cur, after = pins[f"c{c}_kv"].tolist()
        assert (log["cur"], log["after"]) == (cur, after)
        kept = list(range(cur)) if log["kept"] is None else \
            list(range(log["kept"][0])) + list(range(cur - log["kept"][1], cur))
        assert kept == pins[f"c{c}_kept_idx"].tolist()                    # eviction indices bit-exact
        assert orc.st.enc_cache.n_steps == int(pins[f"c{c}_enc_steps"])
        evictions += log["kept"] is not None
    assert evictions >= 3


def test_oracle_reproduces_reference_stream_multiplier_2(pins):
    """Latency multiplier 2: 1920 ms calls, block size 96, 24 speech tokens per turn, 20 new tokens."""
    n, m = int(pins["m2_n_chunks"]), 2
    cfg = tiny_config(max_cache_size=int(pins["max_cache"]), max_llm_cache_size=int(pins["m2_max_llm"]))
    cfg.gen.latency_multiplier, cfg.gen.max_new_tokens = m, 10 * m
    sd = bf16_weights(make_state_dict(cfg, seed=0))
    audio = make_audio(int(pins["n_chunks"]) * SEG / 16000.0)
    orc = OracleStream(cfg, sd)
    for c in range(n):
        # the last call is a SHORT FINAL chunk: half a segment of new audio, padded by the agent to ONE segment
        # (agents/infinisst.py:211-213) -> 12 features for a prompt with 24 <sp_patch> slots; the reference's splice
        # (model/llm.py:101-110) then feeds a shorter sequence to the LLM
        src = audio[: (c + 1) * SEG * m] if c < n - 1 else audio[: c * SEG * m + SEG // 2]
        out_ids, rec, taps = orc.chunk(src.tolist())
        assert taps["speech_feats"].shape[1] == int(pins[f"m2_c{c}_n_feats"]) == (24 if c < n - 1 else 12)
        np.testing.assert_allclose(taps["speech_feats"][0].numpy(), pins[f"m2_c{c}_speech_feats"], atol=1e-5, rtol=1e-5)
        got = torch.stack([l[0] for l in rec.step_logits]).numpy()
        np.testing.assert_allclose(got, pins[f"m2_c{c}_step_logits"], atol=2e-4, rtol=1e-4)
        assert rec.sequences[0] == pins[f"m2_c{c}_sequence"].tolist()
        log = orc.st.kv_log[-1]
        assert [log["cur"], log["after"]] == pins[f"m2_c{c}_kv"].tolist()


@pytest.mark.parametrize("name", SCENARIOS)
def test_eviction_timelines_match_reference(pins, name):
    """agents/infinisst.py:334-361 executed by the reference on a serial-numbered cache vs the oracle's integer
    model and the product's `evict_plan`, including production parameters (window 1000, pinned 40-token system
    prompt), ragged turn lengths and the two quirks (window < turn: `k[:, :, -0:]` keeps everything and
    re-concatenates the system prompt)."""
    from infinisst_b200.agent import S2TAgentStates, evict_plan
    max_llm, keep_sys, n_chunks, sys_n = pins[f"evict_{name}_cfg"].tolist()
    gen = pins[f"evict_{name}_gen"].tolist()
    ref_runs = _runs_of(pins, name, n_chunks)
    ref_cur_after = pins[f"evict_{name}_cur_after"].tolist()
    st_o = O.EvictionState()
    st_p = S2TAgentStates()
    st_p.system_prompt_size = sys_n
    kv_p = 0
    quirk = name in ("tight", "overflow")
    alive, serial = [], 0              # the oracle's cache as token serial numbers (the quirk duplicates entries)
    for c in range(n_chunks):
        prompt = (sys_n + 21) if c == 0 else 22
        grow = prompt + gen[c] - 1
        # oracle (literal restatement, quirks included), applied to the serial-numbered cache like
        # O.apply_eviction applies it to tensors
        pre = alive + list(range(serial, serial + grow))
        serial += grow
        cur = len(pre)
        kept = O.evict(st_o, cur, max_llm, bool(keep_sys), sys_n)
        alive = pre if kept is None else pre[:kept[0]] + pre[cur - kept[1]:]
        idx = [pre.index(t) for t in alive]                                # same read-out as make_ref_pins.py
        runs, a = [], 0
        while a < len(idx):
            b = a
            while b + 1 < len(idx) and idx[b + 1] == idx[b] + 1:
                b += 1
            runs.append((idx[a], idx[b] + 1))
            a = b + 1
        assert [cur, len(alive)] == ref_cur_after[c], (c, cur, len(alive), ref_cur_after[c])
        assert runs == ref_runs[c], (c, runs, ref_runs[c])
        if quirk:
            continue          # the product deliberately does not duplicate the system prompt
        # product host logic: (keep_prefix, drop_upto) -> evicted = [keep_prefix, drop_upto)
        cur_p = kv_p + grow
        plan = evict_plan(st_p, cur_p, max_llm, bool(keep_sys))
        runs_p = [(0, cur_p)] if plan is None else ([(0, plan[0])] if plan[0] else []) + [(plan[1], cur_p)]
        kv_p = sum(b - a for a, b in runs_p)
        assert runs_p == ref_runs[c], (c, runs_p, ref_runs[c])


def test_masks_match_reference(pins):
    """get_attn_mask_training / get_attn_mask_inference (patch_speech_encoder.py:30-77) vs the oracle's
    restatement and the closed form the CUDA kernel evaluates."""
    grid = pins["mask_grid"].tolist()
    for i, (seq, prefix, cache, bs) in enumerate(grid):
        shape = tuple(pins[f"mask_{i}_shape"].tolist())
        ref = np.unpackbits(pins[f"mask_{i}"])[: shape[0] * shape[1]].reshape(shape).astype(bool)
        m = O.mask_streaming(seq, prefix, cache, bs) if prefix > 0 else O.mask_offline(seq, cache, bs)
        assert tuple(m.shape) == shape
        assert np.array_equal((m == 0).numpy(), ref), (seq, prefix, cache, bs)
        cf = O.mask_closed_form(seq, prefix, cache, bs)
        assert np.array_equal((cf == 0).numpy(), ref), (seq, prefix, cache, bs)


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="the reference tree only exists in the build container")
def test_pins_regenerate_from_reference(pins, tmp_path):
    """In the build container: re-run the reference and require the committed fixture to be what it produces."""
    import subprocess
    import sys
    script = os.path.join(os.path.dirname(PINS), "make_ref_pins.py")
    env = dict(os.environ, REF_PINS_OUT=str(tmp_path / "pins.npz"), REF_PINS_CHUNKS="3")
    r = subprocess.run([sys.executable, script], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    fresh = np.load(tmp_path / "pins.npz")
    for c in range(3):
        np.testing.assert_allclose(fresh[f"c{c}_speech_feats"], pins[f"c{c}_speech_feats"], atol=1e-6)
        np.testing.assert_allclose(fresh[f"c{c}_step_logits"], pins[f"c{c}_step_logits"], atol=1e-4)
        assert fresh[f"c{c}_sequence"].tolist() == pins[f"c{c}_sequence"].tolist()


@pytest.mark.gpu
def test_cuda_path_reproduces_reference_stream(pins):
    """The CUDA path against what the reference itself computed (tolerances of tests/test_gpu_parity.py)."""
    from infinisst_b200.agent import S2TAgentStates, evict_plan
    from infinisst_b200.engine import Engine
    ENC_TOL, LOGIT_TOL = 3e-2, 5e-2
    n = int(pins["n_chunks"])
    cfg = tiny_config(max_cache_size=int(pins["max_cache"]), max_llm_cache_size=int(pins["max_llm"]))
    sd = bf16_weights(make_state_dict(cfg, seed=0))
    eng = Engine(cfg, device=0, max_streams=2)
    eng.load_state_dict(sd)
    eng.debug(True)
    audio = make_audio(n * SEG / 16000.0)
    sid = eng.open_stream()
    st = S2TAgentStates()
    st.system_prompt_size = len(cfg.tpl.system_ids)
    target = []
    for c in range(n):
        pcm = audio[c * SEG:(c + 1) * SEG][None].clone()
        if c == 0:
            pcm = torch.cat([torch.zeros(1, 399), pcm], 1)
        feats = eng.encode_chunk([sid], pcm, 1, return_feats=True)
        assert rel_l2(feats[0].cpu(), torch.from_numpy(pins[f"c{c}_speech_feats"])) < ENC_TOL
        seq = pins[f"c{c}_sequence"].tolist()
        ids = O.build_prompt(cfg.tpl, c == 0)
        assert seq[:len(ids)] == ids                                       # the reference agent built the same prompt
        forced = seq[len(ids):]
        toks = eng.generate([sid], [ids], [slot_map(cfg, ids)], [target[-100:]], cfg.gen,
                            pin_prefix=len(cfg.tpl.system_ids), forced=[forced])[0]
        assert toks == forced
        logits = eng.read_tap("step_logits", torch.float32).view(cfg.gen.max_new_tokens, cfg.llm.vocab)
        g = torch.from_numpy(pins[f"c{c}_step_logits"])
        for s in range(g.shape[0]):
            assert rel_l2(logits[s], g[s]) < LOGIT_TOL, (c, s)
            # un-forced choice of the CUDA path: same processors on the CUDA logits must pick the reference's token
            sc = O.process_logits(logits[s], ids + forced[:s], target[-100:], cfg.gen)
            assert int(sc.argmax()) == forced[s] or float(sc.max() - sc[forced[s]]) < 0.35, (c, s)
        target.extend(pins[f"c{c}_output_ids"].tolist())
        cur, after = pins[f"c{c}_kv"].tolist()
        assert eng.kv_len(sid) == cur
        plan = evict_plan(st, cur, cfg.gen.max_llm_cache_size, True)
        kept = list(range(cur)) if plan is None else list(range(plan[0])) + list(range(plan[1], cur))
        assert kept == pins[f"c{c}_kept_idx"].tolist()                    # eviction indices bit-exact vs the reference
        if plan is not None:
            eng.kv_evict(sid, plan[0], plan[1])
        assert eng.kv_len(sid) == after
        assert eng.enc_steps(sid) == int(pins[f"c{c}_enc_steps"])
    eng.close()


@pytest.mark.gpu
def test_cuda_path_reproduces_reference_stream_multiplier_2(pins):
    """Latency multiplier 2 through `model.generate` (the reference-facing call), ending with the SHORT FINAL chunk
    of the pins: 12 speech features for 24 <sp_patch> slots - the surplus slots are not fed (model/llm.py:101-110),
    `sequences` still carries the whole prompt, the KV grows by the shorter sequence."""
    from infinisst_b200.agent import S2TAgentStates, evict_plan
    from infinisst_b200.engine import Engine
    from infinisst_b200.model import SpeechLlamaForCausalLM
    n, m = int(pins["m2_n_chunks"]), 2
    cfg = tiny_config(max_cache_size=int(pins["max_cache"]), max_llm_cache_size=int(pins["m2_max_llm"]))
    cfg.gen.latency_multiplier, cfg.gen.max_new_tokens = m, 10 * m
    g = cfg.gen
    sd = bf16_weights(make_state_dict(cfg, seed=0))
    eng = Engine(cfg, device=0, max_streams=2, max_multiplier=m, max_prompt=64 + 12 * m)
    eng.load_state_dict(sd)
    eng.debug(True)
    model = SpeechLlamaForCausalLM(cfg, engine=eng)
    audio = make_audio(int(pins["n_chunks"]) * SEG / 16000.0)
    st = S2TAgentStates()
    st.system_prompt_size = len(cfg.tpl.system_ids)
    for c in range(n):
        last = c == n - 1
        pcm = audio[c * SEG * m:(c + 1) * SEG * m] if not last else \
            torch.cat([audio[c * SEG * m: c * SEG * m + SEG // 2], torch.zeros(SEG - SEG // 2)])
        pcm = pcm[None].clone()
        if c == 0:
            pcm = torch.cat([torch.zeros(1, 399), pcm], 1)
        seq = pins[f"m2_c{c}_sequence"].tolist()
        ids = O.build_prompt(cfg.tpl, c == 0, m)
        assert seq[:len(ids)] == ids
        forced = seq[len(ids):]
        out = model.generate(input_ids=torch.tensor([ids]), speech_batch=pcm, num_beams=1, max_new_tokens=g.max_new_tokens,
                             encoder_input_ids=[st.target_ids[-100:]], encoder_no_repeat_ngram_size=g.no_repeat_ngram_size,
                             no_repeat_ngram_size=g.no_repeat_ngram_size, repetition_penalty=g.repetition_penalty,
                             pad_token_id=g.pad_token_id, states=st, multiplier=m, forced_tokens=[forced],
                             pin_prefix=len(cfg.tpl.system_ids))
        assert out.sequences[0].tolist() == seq                              # whole prompt + chosen tokens
        assert eng.speech_tokens == int(pins[f"m2_c{c}_n_feats"])
        feats = eng.read_tap("speech_feats").float().view(-1, cfg.llm.hidden)[: eng.speech_tokens]
        assert rel_l2(feats, torch.from_numpy(pins[f"m2_c{c}_speech_feats"])) < 3e-2, c
        logits = eng.read_tap("step_logits", torch.float32).view(g.max_new_tokens, cfg.llm.vocab)
        ref = torch.from_numpy(pins[f"m2_c{c}_step_logits"])
        for s_ in range(ref.shape[0]):
            assert rel_l2(logits[s_], ref[s_]) < 5e-2, (c, s_)
        st.target_ids.extend(forced[:-1])
        cur, after = pins[f"m2_c{c}_kv"].tolist()
        assert st.speech_cache.kv_len == cur, (c, st.speech_cache.kv_len, cur)
        plan = evict_plan(st, cur, g.max_llm_cache_size, True)
        if plan is not None:
            eng.kv_evict(st.speech_cache.sid, plan[0], plan[1])
        assert st.speech_cache.kv_len == after
    eng.close()
