"""SURVEY §8f item 3 against something other than the product itself: a REAL tokenizers-backed HF tokenizer with a
Llama-3.1-style Jinja chat template (tests/golden/llama3_style_tokenizer; 25-token default system header) was handed to
the REFERENCE agent (agents/infinisst.py run from /root/reference under the stand-ins, tests/golden/make_ref_tokenizer_pins.py):
its own `preprocess` extended the tokenizer, its own `_prepare_inputs` built every prompt.  Here the product must build
the same ids from the same tokenizer files (CPU) and reproduce the reference's stream through the CUDA path (GPU)."""
import argparse
import json
import os
import shutil

import numpy as np
import pytest
import torch

from infinisst_b200 import checkpoint as ck
from infinisst_b200 import tiny_config
from infinisst_b200.synthetic import make_audio, make_state_dict
from parity_utils import bf16_weights, rel_l2

HERE = os.path.dirname(os.path.abspath(__file__))
TOK_DIR = os.path.join(HERE, "golden", "llama3_style_tokenizer")
PINS = os.path.join(HERE, "golden", "ref_tokenizer_pins.npz")
SEG = 15360


@pytest.fixture(scope="module")
def pins():
    return np.load(PINS)


def _host_agent(tok, cfg, llama31=True, m=1):
    """The agent's host logic only (no engine, no GPU): what `_prepare_inputs` needs."""
    from infinisst_b200.agent import InfiniSST
    a = InfiniSST.__new__(InfiniSST)
    a.tokenizer, a.cfg, a.llama31 = tok, cfg, llama31
    a.args = argparse.Namespace(block_size=48)
    a.latency_multiplier, a.source_lang, a.target_lang = m, "English", "German"
    return a


def test_product_builds_the_reference_prompts_from_the_same_tokenizer(pins):
    from infinisst_b200.agent import InfiniSST, S2TAgentStates
    tok = InfiniSST._hf_tokenizer(TOK_DIR)                       # AutoTokenizer.from_pretrained(dir, padding_side="right", use_fast=False)
    assert tok is not None and len(tok) == 512 and tok.pad_token == "<|finetune_right_pad_id|>"
    cfg = tiny_config()
    ck.preprocess_tokenizer(tok, cfg, 4)                          # model/llm.py:149-190
    sp, user, assist, start, n_tok, pad = pins["l31_config_ids"].tolist()
    assert (cfg.llm.sp_patch_token_id, cfg.llm.user_token_id, cfg.llm.assist_token_id, cfg.llm.start_header_id) == (sp, user, assist, start)
    assert len(tok) == n_tok == cfg.llm.vocab and tok.pad_token_id == pad == cfg.gen.pad_token_id
    # Llama-3.1 branch: first chunk = system + user + assistant minus the last token; later chunks strip the 25-token
    # default system header and keep its <|eot_id|> (agents/infinisst.py:254-264)
    a = _host_agent(tok, cfg)
    st = S2TAgentStates()
    first = a._prepare_inputs(st)[0].tolist()
    assert first == pins["c0_prompt"].tolist()
    assert st.system_prompt_size == int(pins["system_prompt_size"])
    st.speech_cache = object()
    later = a._prepare_inputs(st)[0].tolist()
    assert later == pins["c1_prompt"].tolist() == pins["c3_prompt"].tolist()
    assert later[0] == tok.convert_tokens_to_ids("<|eot_id|>") and len(later) == 22
    # the splice finds its slots in both (model/llm.py:86-113)
    from infinisst_b200.model import SpeechLlamaForCausalLM
    sm = SpeechLlamaForCausalLM._slot_map(argparse.Namespace(cfg=cfg), first)
    assert sorted(s for s in sm if s >= 0) == list(range(12)) and all(first[i] == sp for i, s in enumerate(sm) if s >= 0)
    # latency multiplier 2: 24 slots, <latency_2> in the system turn
    a2 = _host_agent(tok, cfg, m=2)
    assert a2._prepare_inputs(S2TAgentStates())[0].tolist() == pins["m2_first"].tolist()
    # a model name without "3.1": nothing is stripped, position 0 becomes eos (agents/infinisst.py:265-266)
    a3 = _host_agent(tok, cfg, llama31=False)
    st3 = S2TAgentStates()
    assert a3._prepare_inputs(st3)[0].tolist() == pins["l3_first"].tolist()
    st3.speech_cache = object()
    l3 = a3._prepare_inputs(st3)[0].tolist()
    assert l3 == pins["l3_later"].tolist() and l3[0] == tok.eos_token_id and len(l3) == 22 + 25
    # template_from_tokenizer = the system turn the first chunk carries
    ck.template_from_tokenizer(tok, cfg, "English", "German", 1)
    assert cfg.tpl.system_ids == first[: st.system_prompt_size]


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="the reference tree only exists in the build container")
def test_tokenizer_pins_regenerate_from_reference(pins, tmp_path):
    import subprocess
    import sys
    script = os.path.join(HERE, "golden", "make_ref_tokenizer_pins.py")
    r = subprocess.run([sys.executable, script], env=dict(os.environ, REF_PINS_OUT=str(tmp_path / "p.npz")),
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    fresh = np.load(tmp_path / "p.npz")
    for k in ("c0_prompt", "c1_prompt", "l3_first", "l3_later", "m2_first", "l31_config_ids", "c2_sequence"):
        assert fresh[k].tolist() == pins[k].tolist(), k
    np.testing.assert_allclose(fresh["c3_step_logits"], pins["c3_step_logits"], atol=1e-4)


@pytest.mark.gpu
def test_cuda_agent_from_tokenizer_files_reproduces_reference_stream(pins, tmp_path):
    """`load_model` from files only (agents/infinisst.py:130-183): --model-name is a directory with the tokenizer files,
    config.json and generation_config.json; --state-dict-path an un-pruned checkpoint.  Prompts come from the agent's
    own `_prepare_inputs`; the CUDA path, teacher-forced with the reference's tokens, must reproduce the reference's
    features, step logits and KV lengths."""
    from infinisst_b200.agent import InfiniSST
    cfg = tiny_config(max_cache_size=96, max_llm_cache_size=150)
    sd = bf16_weights(make_state_dict(cfg, seed=0))
    torch.save({"model." + k: v.bfloat16() for k, v in sd.items()}, tmp_path / "pytorch_model.bin")
    d = tmp_path / "Llama-3.1-tiny"
    shutil.copytree(TOK_DIR, d)
    import transformers
    tok0 = transformers.AutoTokenizer.from_pretrained(TOK_DIR)
    eos = [int(tok0.convert_tokens_to_ids(t)) for t in ("<|end_of_text|>", "<|eom_id|>", "<|eot_id|>")]
    (d / "config.json").write_text(json.dumps({            # the fields of a Llama config.json (AutoTokenizer reads it too)
        "model_type": "llama", "architectures": ["LlamaForCausalLM"], "hidden_size": cfg.llm.hidden,
        "intermediate_size": cfg.llm.ffn, "num_hidden_layers": cfg.llm.layers, "vocab_size": 512,
        "max_position_embeddings": 131072, "num_attention_heads": cfg.llm.heads, "num_key_value_heads": cfg.llm.kv_heads,
        "rms_norm_eps": cfg.llm.rms_eps, "rope_theta": cfg.llm.rope_theta,
        "rope_scaling": dict(cfg.llm.rope_scaling, rope_type="llama3")}))
    (d / "generation_config.json").write_text(json.dumps({"eos_token_id": eos}))
    p = argparse.ArgumentParser()
    InfiniSST.add_args(p)
    args = p.parse_args(["--w2v2-type", "w2v2", "--block-size", "48", "--max-cache-size", "96", "--xpos", "0",
                         "--latency-multiplier", "1", "--max-latency-multiplier", "4", "--max-new-tokens", "10",
                         "--no-repeat-ngram-size", "5", "--max-llm-cache-size", "150", "--always-cache-system-prompt",
                         "--beam", "1", "--length-shrink-cfg", "[(128,2,2)] * 2", "--state-dict-path", str(tmp_path / "pytorch_model.bin"),
                         "--model-name", str(d)])
    args.log_chunks = False
    agent = InfiniSST(args)
    assert agent.llama31 and agent.cfg.gen.eos_token_ids == eos and type(agent.tokenizer).__name__ != "TemplateTokenizer"
    eng = agent.model.engine
    eng.debug(True)
    n = int(pins["n_chunks"])
    audio = make_audio(n * SEG / 16000.0)
    st = agent.build_states()
    st.source_sample_rate = 16000
    g = agent.cfg.gen
    for c in range(n):
        st.source = audio[: (c + 1) * SEG].tolist()
        speech = agent._prepare_speech(st)
        ids = agent._prepare_inputs(st)
        assert ids[0].tolist() == pins[f"c{c}_prompt"].tolist()
        seq = pins[f"c{c}_sequence"].tolist()
        forced = seq[ids.shape[1]:]
        out = agent.model.generate(input_ids=ids, speech_batch=speech, num_beams=1, max_new_tokens=10,
                                   encoder_input_ids=[st.target_ids[-100:]], encoder_no_repeat_ngram_size=5,
                                   no_repeat_ngram_size=5, repetition_penalty=1.2, pad_token_id=agent.tokenizer.pad_token_id,
                                   states=st, multiplier=1, forced_tokens=[forced], pin_prefix=st.system_prompt_size)
        assert out.sequences[0].tolist() == seq
        feats = eng.read_tap("speech_feats").float().view(-1, agent.cfg.llm.hidden)[:12]
        assert rel_l2(feats, torch.from_numpy(pins[f"c{c}_speech_feats"])) < 3e-2
        logits = eng.read_tap("step_logits", torch.float32).view(10, agent.cfg.llm.vocab)
        ref = torch.from_numpy(pins[f"c{c}_step_logits"])
        for s in range(ref.shape[0]):
            assert rel_l2(logits[s], ref[s]) < 5e-2, (c, s)
        st.target_ids.extend(forced[:-1])
        st.past_key_values = st.speech_cache
        cur, after = pins[f"c{c}_kv"].tolist()
        assert st.past_key_values[0][0].size(2) == cur
        agent._evict(st)
        assert st.past_key_values[0][0].size(2) == after
    st.reset()
    eng.close()
