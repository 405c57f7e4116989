"""Generates tests/golden/ref_tiny_stream.npz by EXECUTING THE REFERENCE'S OWN CODE from /root/reference.

The reference's third-party dependencies (fairseq, rotary_embedding_torch, simuleval, lightning, transformers
4.47 internals) are absent from this image; tests/golden/ref_standins.py supplies minimal stand-ins for them
(its header lists exactly which lines are the reference's and which are stand-ins).  With those in place this
script drives the reference agent itself:

    InfiniSST.policy (agents/infinisst.py:270-394)
      -> _prepare_speech / _prepare_inputs            (:200-268)
      -> model.generate  [HF greedy `_sample` stand-in; beam search needs transformers 4.47]
           -> SpeechLlamaForCausalLM.forward / SpeechLlamaModel.forward   (model/llm.py:51-126,192-270)
                -> SpeechEncoderW2V2RoPE.encode_speech                      (model/speech_encoder.py:219-236)
                     -> uni_w2v2_forward / encoder / uni_mha_forward         (model/patches/patch_speech_encoder.py)
                -> transformers LlamaModel with llama_sdpa_attention_new_forward (model/patches/patch_llm.py:231-336)
      -> KV eviction + drop-last output slicing        (:334-363)

on the tiny configuration (BASELINE.json configs[0]), the same synthetic weights (reference state-dict key
layout, loaded with the agent's own `load_state_dict`) and the same synthetic audio as tests/golden/make_golden.py,
and records what the reference computed: speech features, per-step last-position logits, sequences, emitted ids,
KV length before / after eviction.  tests/test_ref_pins.py then requires the oracle (CPU) and the CUDA path (GPU)
to reproduce them.  It also records the reference's attention masks for a grid of shapes.

The fixture cannot be regenerated on the GPU box (/root/reference does not exist there): it is committed.
Run:  python tests/golden/make_ref_pins.py
"""
import os
import sys
import tempfile
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, HERE)

import ref_standins as RS                                             # noqa: E402
from infinisst_b200 import tiny_config                                 # noqa: E402
from infinisst_b200.synthetic import make_audio, make_state_dict      # noqa: E402
from parity_utils import bf16_weights                                  # noqa: E402
from oracle import infinisst_oracle as O                               # noqa: E402  (prompt ids for the fake tokenizer only)

N_CHUNKS = int(os.environ.get("REF_PINS_CHUNKS", "8"))
MAX_CACHE, MAX_LLM = 96, 150
SEG_MS = 960


class FakeTokenizer:
    """Stands in for the Llama-3.1 tokenizer (not in this image): `apply_chat_template` returns the synthetic
    template of infinisst_b200.config.TemplateConfig laid out the way the real template is - the agent's own
    slicing (`[:, :-1]`, `[:, 25:]`, system_prompt_size) then runs on it unchanged."""

    def __init__(self, cfg):
        self.cfg = cfg
        self.pad_token = None
        self.pad_token_id = cfg.gen.pad_token_id
        self.eos_token_id = cfg.tpl.eot_id

    def _turn(self, role_id, body):
        t = self.cfg.tpl
        return [t.start_header_id, role_id, t.end_header_id, t.nl_id] + body + [t.eot_id]

    def apply_chat_template(self, conversations, return_tensors="pt", **kw):
        t = self.cfg.tpl
        (messages,) = conversations
        ids = []
        if messages and messages[0]["role"] == "system":
            ids += list(t.system_ids)
            messages = messages[1:]
        else:
            # Llama-3.1's template emits a default system turn when there is no system message: 25 tokens (BOS,
            # header, "Cutting Knowledge Date ... Today Date ...") and then <|eot_id|>.  The agent strips exactly
            # the 25 and keeps the EOT ("to remove system prompt and preserve last EOT", agents/infinisst.py:262-264)
            ids += [t.system_ids[0]] + [1] * 24 + [t.eot_id]
        for m in messages:
            if m["role"] == "user":
                n = m["content"].count("<sp_patch>")
                ids += self._turn(t.user_token_id, [t.sp_patch_id] * n)
            else:
                ids += self._turn(t.assist_token_id, [])
        return torch.tensor([ids], dtype=torch.long)

    def decode(self, ids, skip_special_tokens=True):
        return " ".join(str(i) for i in ids)


def build_reference_agent(cfg, sd, xpos=0, rope=1, tokenizer=None, model_name="synthetic-llama-3.1"):
    """`tokenizer`: a real HF tokenizer to hand to the agent (then the reference's OWN `preprocess`, model/llm.py:149-190,
    runs on it); default: the FakeTokenizer over the synthetic template with a stand-in `preprocess`."""
    RS.install()
    import transformers
    from transformers import LlamaConfig
    import model.llm as ref_llm
    import model.speech_encoder as ref_se
    import model.patches.patch_llm as ref_pl
    import agents.infinisst as ref_agent
    import transformers.models.llama.modeling_llama as ml

    e, l, g = cfg.enc, cfg.llm, cfg.gen

    hf_cfg = ref_llm.SpeechLlamaConfig(
        vocab_size=l.vocab, hidden_size=l.hidden, intermediate_size=l.ffn, num_hidden_layers=l.layers,
        num_attention_heads=l.heads, num_key_value_heads=l.kv_heads, head_dim=l.head_dim, rms_norm_eps=l.rms_eps,
        rope_theta=l.rope_theta, max_position_embeddings=131072, attention_bias=False, mlp_bias=False,
        tie_word_embeddings=False, attn_implementation="eager",
        rope_scaling=dict(rope_type="llama3", rope_theta=l.rope_theta, **l.rope_scaling))

    def from_pretrained(cls, name, torch_dtype=None, device_map=None, **kw):
        torch.manual_seed(0)
        m = cls(hf_cfg)
        for layer in m.model.layers:
            object.__setattr__(layer.self_attn, "rotary_emb", m.model.rotary_emb)
        return m.float()
    ref_llm.SpeechLlamaForCausalLM.from_pretrained = classmethod(from_pretrained)

    def preprocess(self, tokenizer, max_multiplier=4, resize=True):
        # llm.py:149-190 adds 7 tokens to a real tokenizer and records these ids on the config; the synthetic
        # vocabulary already contains them
        self.config.sp_patch_token_id = l.sp_patch_token_id
        self.config.user_token_id = l.user_token_id
        self.config.assist_token_id = l.assist_token_id
        self.config.start_header_id = l.start_header_id
    if tokenizer is None:
        ref_llm.SpeechLlamaForCausalLM.preprocess = preprocess

    def _load_w2v2(self, path, finetuned):
        # fairseq checkpoint loading is third-party; the container classes come from the stand-in, their
        # forward methods from the reference's patch_w2v2 (already applied by load_model)
        from fairseq.models.wav2vec import Wav2Vec2Model
        return Wav2Vec2Model(e.conv_layers, e.embed_dim, e.ffn_dim, e.heads, e.layers), e.embed_dim, e.layers
    ref_se.SpeechEncoderW2V2RoPE._load_w2v2 = _load_w2v2

    transformers.AutoTokenizer.from_pretrained = staticmethod(lambda *a, **k: tokenizer if tokenizer is not None else FakeTokenizer(cfg))

    with tempfile.NamedTemporaryFile(suffix=".bin", delete=False) as f:
        torch.save({k: v.clone() for k, v in sd.items()}, f.name)
        sd_path = f.name

    adapter = "[" + ", ".join(str(tuple(x)) for x in e.adapter_layers) + "]"
    args = types.SimpleNamespace(
        min_start_sec=0, latency_multiplier=1, source_segment_size=SEG_MS, max_latency_multiplier=4,
        source_lang="English", target_lang="German", beam=2, no_repeat_ngram_lookback=g.no_repeat_ngram_lookback,
        no_repeat_ngram_size=g.no_repeat_ngram_size, repetition_penalty=g.repetition_penalty,
        suppress_non_language=False, max_len_a=1, max_len_b=256, max_new_tokens=g.max_new_tokens, do_sample=False,
        top_p=1.0, top_k=0, epsilon_cutoff=0.0, temperature=1.0, pseudo_batch_size=1,
        max_llm_cache_size=g.max_llm_cache_size, always_cache_system_prompt=g.always_cache_system_prompt,
        dpo_sampling=False, model_name=model_name, w2v2_path="synthetic", ctc_finetuned=True,
        w2v2_type="w2v2", length_shrink_cfg=adapter, block_size=e.block_size, max_cache_size=e.max_cache_size,
        xpos=xpos, rope=rope, state_dict_path=sd_path)
    # torch.cuda device placement: the agent moves tensors to model.device, which is the CPU here
    agent = ref_agent.InfiniSST(args)          # runs the reference's __init__ and load_model
    os.unlink(sd_path)
    # transformers 5.5 has one LlamaAttention class: route it to the reference's SDPA patch - what patch_llm()
    # (called by load_model above) installs on 4.47's LlamaSdpaAttention, the class inference instantiates since
    # the agent passes no attn_implementation (agents/infinisst.py:150-154).
    def attn_forward(self, hidden_states, position_embeddings=None, attention_mask=None, past_key_values=None,
                     cache_position=None, **kw):
        out, _, _ = ref_pl.llama_sdpa_attention_new_forward(
            self, hidden_states=hidden_states, attention_mask=attention_mask, past_key_value=past_key_values,
            use_cache=True, cache_position=cache_position, position_embeddings=position_embeddings)
        return out, None
    ml.LlamaAttention.forward = attn_forward
    agent.beam = 1                               # greedy: the `assert beam > 1` guards the shipped beam path

    taps = {}
    se = agent.model.model.speech_encoder
    orig_encode = se.encode_speech

    def encode_tap(*a, **k):
        feat, cache = orig_encode(*a, **k)
        taps["speech_feats"] = feat.detach().clone()
        return feat, cache
    se.encode_speech = encode_tap

    def generate(**kw):
        out = RS.greedy_generate(agent.model, eos_token_ids=g.eos_token_ids, **kw)
        out.past_key_values_pre = [out.past_key_values.key_cache[-1].clone()]   # last layer: rows depend on the context, so they are unique
        taps["gen"] = out
        return out
    agent.model.generate = generate
    return agent, taps


def main():
    torch.set_num_threads(4)
    cfg = tiny_config(max_cache_size=MAX_CACHE, max_llm_cache_size=MAX_LLM)
    sd = bf16_weights(make_state_dict(cfg, seed=0))
    agent, taps = build_reference_agent(cfg, sd)
    seg = 15360
    audio = make_audio(N_CHUNKS * seg / 16000.0)
    states = agent.build_states()
    states.reset()                                   # SimulEval resets the states before every instance
    states.source_sample_rate = 16000
    out = {"n_chunks": np.int32(N_CHUNKS), "max_cache": np.int32(MAX_CACHE), "max_llm": np.int32(MAX_LLM)}
    for c in range(N_CHUNKS):
        states.source = audio[: (c + 1) * seg].tolist()
        kv_before = 0 if states.past_key_values is None else states.past_key_values[0][0].size(2)
        agent.policy(states)
        k_pre = taps["gen"].past_key_values_pre[0][0, 0]       # last-layer K right after generate (un-rotated rows)
        k_post = states.past_key_values[len(states.past_key_values) - 1][0][0, 0]
        kept_idx = [int((k_pre == row).all(dim=1).nonzero()[0]) for row in k_post]
        out[f"c{c}_kept_idx"] = np.array(kept_idx, dtype=np.int32)
        gen = taps["gen"]
        out[f"c{c}_speech_feats"] = taps["speech_feats"][0].numpy().astype(np.float32)
        out[f"c{c}_step_logits"] = torch.stack([x[0] for x in gen.step_logits]).numpy().astype(np.float32)
        out[f"c{c}_sequence"] = gen.sequences[0].numpy().astype(np.int32)
        n_prompt = gen.sequences.size(1) - len(gen.step_logits)
        out[f"c{c}_output_ids"] = gen.sequences[0, n_prompt:-1].numpy().astype(np.int32)
        cur = kv_before + gen.sequences.size(1) - 1                     # prompt + generated[:-1]
        after = states.past_key_values[0][0].size(2)
        out[f"c{c}_kv"] = np.array([cur, after], dtype=np.int32)
        out[f"c{c}_enc_steps"] = np.int32(states.speech_cache.n_steps)
        out[f"c{c}_target_len"] = np.int32(len(states.target_ids))
        print(f"chunk {c}: kv {kv_before} -> {cur} -> {after}, emitted {out[f'c{c}_output_ids'].tolist()}")
    # ---- latency multiplier 2: 1920 ms policy calls, 96-frame blocks, 24 speech tokens per
    #      turn, max_new_tokens 20 (agents/infinisst.py:125-128,245; speech_encoder.py:143-145) ----
    if N_CHUNKS >= 8:
        m = 2
        agent.update_multiplier(m)
        agent.max_llm_cache_size, agent.cache_checkpoints = 200, []
        st_m = agent.build_states()
        st_m.reset()
        st_m.source_sample_rate = 16000
        n_m = 4
        out["m2_n_chunks"], out["m2_max_llm"] = np.int32(n_m), np.int32(200)
        for c in range(n_m):
            st_m.source = audio[: (c + 1) * seg * m].tolist()
            if c == n_m - 1:
                # short FINAL chunk at m = 2: the agent pads to ONE segment only
                # (agents/infinisst.py:211-213): half a segment of new audio becomes 48 frames = 12 features for a
                # prompt that still has 24 <sp_patch> slots; the splice (model/llm.py:101-110) then yields a SHORTER
                # sequence, and states.source_finished sets segment_idx = -1 (:303-304)
                st_m.source = audio[: c * seg * m + seg // 2].tolist()
                st_m.source_finished = True
            kv_before = 0 if st_m.past_key_values is None else st_m.past_key_values[0][0].size(2)
            agent.policy(st_m)
            gen = taps["gen"]
            out[f"m2_c{c}_speech_feats"] = taps["speech_feats"][0].numpy().astype(np.float32)
            out[f"m2_c{c}_step_logits"] = torch.stack([x[0] for x in gen.step_logits]).numpy().astype(np.float32)
            out[f"m2_c{c}_sequence"] = gen.sequences[0].numpy().astype(np.int32)
            # KV length right after generate = last-layer K rows of the returned cache (a short final chunk feeds
            # fewer positions than `sequences` has prompt ids)
            out[f"m2_c{c}_kv"] = np.array([gen.past_key_values_pre[0].size(2), st_m.past_key_values[0][0].size(2)], dtype=np.int32)
            out[f"m2_c{c}_n_feats"] = np.int32(taps["speech_feats"].shape[1])
            print(f"m=2 chunk {c}: kv {kv_before} -> {out[f'm2_c{c}_kv'].tolist()}")
        agent.update_multiplier(1)
    # ---- eviction timelines: the agent's own eviction code (agents/infinisst.py:334-361) on a cache whose
    #      entries are token serial numbers, so the kept index set can be read off directly.  `generate` is a
    #      stub that only grows the cache by prompt + n_generated - 1 entries (drop-last rule). ----
    scenarios = [  # (name, max_llm_cache_size, always_cache_system_prompt, n_chunks, gen-length pattern)
        ("prod", 1000, True, 80, [10]),                                  # no EOS, 9 forwarded tokens
        ("prod_ragged", 1000, True, 120, [10, 3, 7, 1, 10, 5, 2, 9]),     # early EOS: variable turn lengths
        ("nosys", 300, False, 40, [10, 4, 8]),                           # --always-cache-system-prompt off
        ("tight", 64, True, 12, [10]),                                   # window barely larger than a turn
        ("overflow", 20, True, 6, [10]),                                 # a single turn exceeds the window (quirk: -0 slice)
    ]
    for name, max_llm, keep_sys, n_chunks, pattern in scenarios:
        agent.max_llm_cache_size, agent.always_cache_system_prompt = max_llm, keep_sys
        agent.cache_checkpoints = []
        st2 = agent.build_states()
        st2.reset()
        st2.source_sample_rate = 16000
        serial = [0]
        step = [0]

        def fake_generate(input_ids=None, past_key_values=None, max_new_tokens=10, **kw):
            n_gen = min(pattern[step[0] % len(pattern)], max_new_tokens)
            step[0] += 1
            n_add = input_ids.size(1) + n_gen - 1
            tags = torch.arange(serial[0], serial[0] + n_add, dtype=torch.float32).view(1, 1, n_add, 1)
            serial[0] += n_add
            cache = past_key_values if past_key_values is not None else RS.DynamicCache447()
            cache.update(tags, tags.clone(), 0)
            st2.speech_cache = object()                                   # "not the first chunk any more"
            seq = torch.cat([input_ids, torch.full((1, n_gen), 7, dtype=torch.long)], dim=1)
            return types.SimpleNamespace(sequences=seq, past_key_values=cache)
        agent.model.generate = fake_generate
        rows = []
        alive = []                                                        # serial numbers currently in the cache
        for c in range(n_chunks):
            st2.source = [0.0] * ((c + 1) * seg)
            serial_before = serial[0]
            agent.policy(st2)
            kept = st2.past_key_values[0][0][0, 0, :, 0].to(torch.int64).tolist()
            pre = alive + list(range(serial_before, serial[0]))          # cache contents right after generate
            idx = [pre.index(t) for t in kept]                            # logical indices kept (first match for duplicates)
            rows.append((len(pre), len(kept), idx))
            alive = kept
        out[f"evict_{name}_cfg"] = np.array([max_llm, int(keep_sys), n_chunks, agent.system_prompt_size], dtype=np.int32)
        out[f"evict_{name}_gen"] = np.array([pattern[i % len(pattern)] for i in range(n_chunks)], dtype=np.int32)
        out[f"evict_{name}_cur_after"] = np.array([(r[0], r[1]) for r in rows], dtype=np.int32)
        # kept logical indices as (start, stop) runs
        runs = []
        for ci, (_, _, idx) in enumerate(rows):
            a = 0
            while a < len(idx):
                b = a
                while b + 1 < len(idx) and idx[b + 1] == idx[b] + 1:
                    b += 1
                runs.append((ci, idx[a], idx[b] + 1))
                a = b + 1
        out[f"evict_{name}_runs"] = np.array(runs, dtype=np.int32)
        n_ev = sum(1 for r in rows if r[1] != r[0])
        print(f"eviction scenario {name}: {n_chunks} chunks, {n_ev} evictions, final kv {rows[-1][1]}")
    # the reference's masks (patch_speech_encoder.py:30-77) on a grid of shapes
    import model.patches.patch_speech_encoder as ref_pse
    grid = []
    for bs in (4, 48):
        for cache in (8, 96, 576):
            for seq in (bs, 2 * bs):
                grid.append((seq, 0, cache, bs))
                for prefix in (bs, 3 * bs, cache, cache + bs, 2 * cache + 3 * bs):
                    grid.append((seq, prefix, cache, bs))
    out["mask_grid"] = np.array(grid, dtype=np.int32)
    for i, (seq, prefix, cache, bs) in enumerate(grid):
        m = (ref_pse.get_attn_mask_inference(seq, prefix, cache, bs, "cpu") if prefix > 0
             else ref_pse.get_attn_mask_training(seq, cache, bs, "cpu"))
        out[f"mask_{i}"] = np.packbits((m == 0).numpy())
        out[f"mask_{i}_shape"] = np.array(m.shape, dtype=np.int32)
    path = os.environ.get("REF_PINS_OUT") or os.path.join(HERE, "ref_tiny_stream.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
