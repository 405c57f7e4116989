"""Checkpoint / tokenizer ingestion (agents/infinisst.py:130-183, model/llm.py:149-190,
speech_encoder.py:147-172, train/prune_bin.py).  Host logic only: runs without a GPU."""
import argparse
import json
import os

import pytest
import torch

from infinisst_b200 import checkpoint as ck
from infinisst_b200.config import production_config, tiny_config
from infinisst_b200.synthetic import make_state_dict, weight_specs


def _tiny_sd():
    cfg = tiny_config()
    return cfg, make_state_dict(cfg, seed=0)


def test_infer_config_roundtrip_tiny(tmp_path):
    cfg, sd = _tiny_sd()
    p = tmp_path / "pytorch_model.bin"
    torch.save(sd, p)
    got, sd2 = ck.load_checkpoint(str(p), block_size=48, max_cache_size=576)
    assert got.enc.conv_layers == [tuple(x) for x in cfg.enc.conv_layers]
    for f in ("embed_dim", "ffn_dim", "heads", "layers", "llm_dim", "block_size", "max_cache_size"):
        assert getattr(got.enc, f) == getattr(cfg.enc, f), f
    assert got.enc.adapter_layers == [tuple(x) for x in cfg.enc.adapter_layers]
    for f in ("hidden", "layers", "heads", "kv_heads", "head_dim", "ffn", "vocab", "rms_eps", "rope_theta"):
        assert getattr(got.llm, f) == getattr(cfg.llm, f), f
    assert got.llm.rope_scaling == cfg.llm.rope_scaling
    assert set(sd2) == set(sd)


def test_infer_config_production_shapes_without_allocating():
    """The production architecture is recovered from shapes alone (meta tensors: no 16 GB allocation)."""
    cfg = production_config()
    sd = {k: torch.empty(shape, device="meta") for k, shape, _, _ in weight_specs(cfg)}
    got = ck.infer_config(sd, block_size=48, max_cache_size=576, length_shrink_cfg="[(1024,2,2)] * 2")
    assert (got.enc.layers, got.enc.embed_dim, got.enc.ffn_dim, got.enc.heads) == (24, 1024, 4096, 16)
    assert got.enc.conv_layers == [(512, 10, 5)] + [(512, 3, 2)] * 4 + [(512, 2, 2)] * 2
    assert got.enc.adapter_layers == [(1024, 2, 2)] * 2
    l = got.llm
    assert (l.layers, l.hidden, l.ffn, l.heads, l.kv_heads, l.head_dim, l.vocab) == (32, 4096, 14336, 32, 8, 128, 128263)


def test_lightning_prefix_is_pruned(tmp_path):
    """train/prune_bin.py:7-9: an un-pruned Lightning dump has `model.` in front of every key."""
    cfg, sd = _tiny_sd()
    p = tmp_path / "unpruned.bin"
    torch.save({"model." + k: v for k, v in sd.items()}, p)
    assert set(ck.load_reference_state_dict(str(p))) == set(sd)
    p2 = tmp_path / "lightning.ckpt"
    torch.save({"state_dict": {"model." + k: v for k, v in sd.items()}, "epoch": 3}, p2)
    assert set(ck.load_reference_state_dict(str(p2))) == set(sd)


def test_strict_load_errors_use_torch_wording():
    cfg, sd = _tiny_sd()
    bad = dict(sd)
    del bad["model.norm.weight"]
    bad["model.layers.0.extra.weight"] = torch.zeros(1)
    bad["lm_head.weight"] = torch.zeros(3, 3)
    with pytest.raises(RuntimeError) as e:
        ck.check_state_dict(bad, cfg)
    msg = str(e.value)
    assert 'Missing key(s) in state_dict: "model.norm.weight"' in msg
    assert 'Unexpected key(s) in state_dict: "model.layers.0.extra.weight"' in msg
    assert "size mismatch for lm_head.weight" in msg
    # tensors the reference module tree holds but the step never reads are tolerated
    ok = dict(sd)
    ok[ck.ENC + "mask_emb"] = torch.zeros(cfg.enc.embed_dim)
    ok[ck.ENC + "encoder.pos_conv.0.bias"] = torch.zeros(cfg.enc.embed_dim)
    ck.check_state_dict(ok, cfg)


def test_groupnorm_extractor_is_rejected_like_the_reference():
    cfg, sd = _tiny_sd()
    gn = {k: v for k, v in sd.items() if ".2.1." not in k or "length_shrink" in k}
    gn[ck.ENC + "feature_extractor.conv_layers.0.2.weight"] = torch.ones(cfg.enc.conv_layers[0][0])
    with pytest.raises(NotImplementedError, match="layer_norm_first"):
        ck.infer_config(gn)


@pytest.mark.parametrize("text,want", [
    ("[(1024,2,2)] * 2", [(1024, 2, 2)] * 2),
    ("[(512,10,5)] + [(512,3,2)] * 4 + [(512,2,2)] * 2", [(512, 10, 5)] + [(512, 3, 2)] * 4 + [(512, 2, 2)] * 2),
    ("[(64, 2, 2)]", [(64, 2, 2)]),
])
def test_parse_conv_cfg(text, want):
    assert ck.parse_conv_cfg(text) == want


def test_parse_conv_cfg_refuses_code():
    with pytest.raises(ValueError):
        ck.parse_conv_cfg("__import__('os').system('true')")


def test_hf_config_and_generation_config(tmp_path):
    cfg, sd = _tiny_sd()
    d = tmp_path / "Llama-tiny"
    d.mkdir()
    (d / "config.json").write_text(json.dumps({
        "num_attention_heads": cfg.llm.heads, "num_key_value_heads": cfg.llm.kv_heads, "rms_norm_eps": 1e-6,
        "rope_theta": 10000.0, "rope_scaling": None}))
    (d / "generation_config.json").write_text(json.dumps({"eos_token_id": [5, 6]}))
    hf, gen = ck.read_hf_config(str(d))
    got = ck.infer_config(sd, hf_config=hf, generation_config=gen)
    assert got.llm.head_dim == cfg.llm.head_dim and got.llm.rms_eps == 1e-6 and got.llm.rope_theta == 10000.0
    assert got.llm.rope_scaling is None and got.gen.eos_token_ids == [5, 6]
    assert ck.read_hf_config("meta-llama/Llama-3.1-8B-Instruct") == (None, None)     # a hub id: defaults apply


def test_w2v2_checkpoint_args_and_key_mapping(tmp_path):
    cfg, sd = _tiny_sd()
    conv = "[(64,10,5)] + [(64,3,2)] * 4 + [(64,2,2)] * 2"
    ns = argparse.Namespace(conv_feature_layers=conv, encoder_layers=2, encoder_embed_dim=128,
                            encoder_ffn_embed_dim=256, encoder_attention_heads=2, extractor_mode="layer_norm",
                            layer_norm_first=True, conv_bias=True)
    model = {k[len(ck.ENC):]: v for k, v in sd.items() if k.startswith(ck.ENC)}
    p = tmp_path / "w2v2.pt"
    torch.save({"args": ns, "model": model}, p)
    wa = ck.read_w2v2_args(str(p))
    assert wa["conv_feature_layers"] == conv and wa["encoder_layers"] == 2
    got = ck.infer_config(sd, w2v2_args=wa)
    assert got.enc.conv_layers == ck.parse_conv_cfg(conv)
    mapped = ck.w2v2_state_to_reference(model)
    assert all(torch.equal(mapped[k], sd[k]) for k in mapped) and len(mapped) == len(model)
    # CTC fine-tuned layout (speech_encoder.py:157-170)
    ctc = {"w2v_encoder.w2v_model." + k: v for k, v in model.items()}
    ctc["w2v_encoder.proj.weight"] = torch.zeros(4, 4)
    mapped = ck.w2v2_state_to_reference(ctc, ctc_finetuned=True)
    assert set(mapped) == {ck.ENC + k for k in model}
    # cfg-tree form + a post-LN model is refused (patch_speech_encoder.py:571)
    p2 = tmp_path / "w2v2_base.pt"
    torch.save({"cfg": {"model": {"w2v_args": {"model": {"layer_norm_first": False, "extractor_mode": "default"}}}},
                "model": {}}, p2)
    with pytest.raises(NotImplementedError):
        ck.read_w2v2_args(str(p2))


class _ToyTokenizer:
    """Just enough of a HF tokenizer for `preprocess` and the agent's prompt building."""

    def __init__(self, base):
        self.vocab = {t: i for i, t in enumerate(base)}
        self.pad_token, self.pad_token_id = "<|finetune_right_pad_id|>", base.index("<|finetune_right_pad_id|>")

    def add_tokens(self, toks, special_tokens=False):
        n = 0
        for t in toks:
            if t not in self.vocab:
                self.vocab[t] = len(self.vocab)
                n += 1
        return n

    def __len__(self):
        return len(self.vocab)

    def convert_tokens_to_ids(self, t):
        if isinstance(t, (list, tuple)):
            return [self.vocab.get(x) for x in t]
        return self.vocab.get(t)

    def apply_chat_template(self, batch, **kw):
        out = []
        for m in batch[0]:
            out += [self.vocab["<|start_header_id|>"], self.vocab[m["role"]], self.vocab["<|end_header_id|>"]]
            out += [self.vocab["w"]] * len(m["content"].split())
            out += [self.vocab["<|eot_id|>"]]
        return torch.tensor([out])


def test_preprocess_tokenizer_assigns_reference_ids():
    cfg, _ = _tiny_sd()
    base = [f"t{i}" for i in range(cfg.llm.vocab - 7 - 8)] + ["user", "assistant", "system", "w", "<|start_header_id|>",
                                                            "<|end_header_id|>", "<|eot_id|>", "<|finetune_right_pad_id|>"]
    tok = _ToyTokenizer(base)
    V = len(base)
    ck.preprocess_tokenizer(tok, cfg, max_multiplier=4)
    assert len(tok) == cfg.llm.vocab
    # model/llm.py:150-159: <sp_patch>, <sp_start>, <sp_end>, <latency_1..4> take the next 7 ids in this order
    assert tok.convert_tokens_to_ids(["<sp_patch>", "<sp_start>", "<sp_end>", "<latency_1>", "<latency_4>"]) == \
        [V, V + 1, V + 2, V + 3, V + 6]
    assert cfg.llm.sp_patch_token_id == cfg.tpl.sp_patch_id == V
    assert cfg.llm.user_token_id == base.index("user") and cfg.llm.assist_token_id == base.index("assistant")
    assert cfg.llm.start_header_id == base.index("<|start_header_id|>")
    assert cfg.gen.pad_token_id == base.index("<|finetune_right_pad_id|>")
    ck.template_from_tokenizer(tok, cfg, "English", "German", 1)
    assert cfg.tpl.system_ids[:3] == [cfg.llm.start_header_id, base.index("system"), cfg.tpl.end_header_id]
    assert cfg.tpl.system_ids[-1] == cfg.tpl.eot_id
    # a tokenizer / checkpoint size mismatch is an error, not a silent resize (resize=False, agents/infinisst.py:177)
    cfg2, _ = _tiny_sd()
    with pytest.raises(ValueError, match="embedding table"):
        ck.preprocess_tokenizer(_ToyTokenizer(base), cfg2, max_multiplier=2)


@pytest.mark.gpu
def test_agent_loads_from_checkpoint_files_like_the_reference(tmp_path):
    """agents/infinisst.py:130-183 end to end: `--state-dict-path` (an un-pruned Lightning dump), `--model-name`
    (a directory with config.json + generation_config.json), a HF-style tokenizer that goes through `preprocess`
    and `apply_chat_template` (+ the agent's `[:, :-1]` / `[:, 25:]` slicing).  No architecture is handed in: it is
    read off the files.  The stream must equal the one of an agent built from the in-memory config."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    from make_ref_pins import FakeTokenizer
    from infinisst_b200.agent import InfiniSST
    from infinisst_b200.synthetic import make_audio
    from parity_utils import bf16_weights

    cfg = tiny_config(max_cache_size=96, max_llm_cache_size=150)
    sd = bf16_weights(make_state_dict(cfg, seed=0))
    torch.save({"model." + k: v.bfloat16() for k, v in sd.items()}, tmp_path / "pytorch_model.bin")
    d = tmp_path / "Llama-3.1-tiny"
    d.mkdir()
    (d / "config.json").write_text(json.dumps({
        "num_attention_heads": cfg.llm.heads, "num_key_value_heads": cfg.llm.kv_heads, "rms_norm_eps": cfg.llm.rms_eps,
        "rope_theta": cfg.llm.rope_theta, "rope_scaling": dict(cfg.llm.rope_scaling, rope_type="llama3")}))
    (d / "generation_config.json").write_text(json.dumps({"eos_token_id": cfg.gen.eos_token_ids}))

    class Tok(FakeTokenizer):
        names = {"<sp_patch>": cfg.tpl.sp_patch_id, "<sp_start>": cfg.tpl.sp_patch_id + 1, "<sp_end>": cfg.tpl.sp_patch_id + 2,
                 "user": cfg.tpl.user_token_id, "assistant": cfg.tpl.assist_token_id,
                 "<|start_header_id|>": cfg.tpl.start_header_id, "<|end_header_id|>": cfg.tpl.end_header_id,
                 "<|eot_id|>": cfg.tpl.eot_id}

        def add_tokens(self, toks, special_tokens=False):
            return 0                                    # the synthetic vocabulary already holds the 7 speech tokens

        def __len__(self):
            return cfg.llm.vocab

        def convert_tokens_to_ids(self, t):
            return [self.names.get(x) for x in t] if isinstance(t, (list, tuple)) else self.names.get(t)

    def make_args(**kw):
        p = argparse.ArgumentParser()
        InfiniSST.add_args(p)
        a = p.parse_args(["--w2v2-type", "w2v2", "--block-size", "48", "--max-cache-size", "96", "--xpos", "0",
                          "--latency-multiplier", "1", "--max-latency-multiplier", "4", "--max-new-tokens", "10",
                          "--no-repeat-ngram-size", "5", "--max-llm-cache-size", "150", "--always-cache-system-prompt",
                          "--beam", "1", "--length-shrink-cfg", "[(128,2,2)] * 2"])
        for k, v in kw.items():
            setattr(a, k, v)
        return a

    ref_agent = InfiniSST(make_args(model_config=cfg, state_dict=sd))
    agent = InfiniSST(make_args(state_dict_path=str(tmp_path / "pytorch_model.bin"), model_name=str(d), tokenizer=Tok(cfg)))
    got = agent.cfg
    assert (got.llm.layers, got.llm.heads, got.llm.kv_heads, got.llm.head_dim, got.llm.vocab) == \
        (cfg.llm.layers, cfg.llm.heads, cfg.llm.kv_heads, cfg.llm.head_dim, cfg.llm.vocab)
    assert got.enc.conv_layers == [tuple(x) for x in cfg.enc.conv_layers] and got.gen.eos_token_ids == cfg.gen.eos_token_ids
    assert (got.llm.user_token_id, got.llm.assist_token_id, got.llm.sp_patch_token_id) == \
        (cfg.llm.user_token_id, cfg.llm.assist_token_id, cfg.llm.sp_patch_token_id)
    seg, n = 15360, 8
    audio = make_audio(n * seg / 16000.0)
    sa, sb = ref_agent.build_states(), agent.build_states()
    sa.source_sample_rate = sb.source_sample_rate = 16000
    for c in range(n):
        for st in (sa, sb):
            st.source = audio[: (c + 1) * seg].tolist()
            st.source_finished = c == n - 1
        ra, rb = ref_agent.policy(sa), agent.policy(sb)
        assert sa.target_ids == sb.target_ids, f"chunk {c}"
        assert sa.past_key_values[0][0].size(2) == sb.past_key_values[0][0].size(2)
        assert type(ra) is type(rb)
    assert sb.system_prompt_size == len(cfg.tpl.system_ids) and len(sb.target_ids) > 20
    assert sb.past_key_values[0][0].size(2) <= 150 + len(cfg.tpl.system_ids)      # the window slid


@pytest.mark.gpu
def test_engine_ignores_rotary_buffers_some_checkpoints_carry():
    """A `--xpos 1` checkpoint saved with a rotary_embedding_torch version whose `scale` buffer is persistent, or an
    older HF checkpoint with per-layer `rotary_emb.inv_freq`, passes the strict key check AND loads: those buffers are
    recomputed by the library."""
    from infinisst_b200.engine import Engine
    from parity_utils import bf16_weights
    cfg, sd = _tiny_sd()
    sd = bf16_weights(sd)
    extra = dict(sd)
    extra[ck.ENC + "encoder.layers.0.self_attn.rotary_emb.scale"] = torch.ones(cfg.enc.head_dim // 2)
    extra[ck.ENC + "encoder.layers.1.self_attn.rotary_emb.dummy"] = torch.zeros(1)
    extra["model.layers.0.self_attn.rotary_emb.inv_freq"] = torch.ones(cfg.llm.head_dim // 2)
    ck.check_state_dict(extra, cfg)
    eng = Engine(cfg, device=0, max_streams=1)
    eng.load_state_dict(extra)
    sid = eng.open_stream()
    assert eng.kv_len(sid) == 0
    eng.close()
