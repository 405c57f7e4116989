"""Checkpoint and tokenizer ingestion for the agent's `load_model` (agents/infinisst.py:130-183).

The reference builds its model from three artefacts:
  --model-name       a HF Llama directory (config.json, generation_config.json, tokenizer files)  :135-154
  --w2v2-path        a fairseq wav2vec 2.0 checkpoint, read for its architecture arguments       speech_encoder.py:147-172
  --state-dict-path  one `pytorch_model.bin` with every trained tensor, `load_state_dict`'ed      :179-180
This module turns the same artefacts into an `InfiniSSTConfig` plus a reference-layout state dict
for `Engine.load_state_dict` (which repacks into the kernel layouts).  Nothing here computes on the
hot path; it is load-time host code, Python because the reference's is.

  load_reference_state_dict   torch.load(weights_only=True) of the .bin; accepts the un-pruned Lightning
                              layout that train/prune_bin.py:1-11 strips (`model.` in front of every key)
  infer_config                every dimension that a tensor shape determines is read off the state dict;
                              what shapes cannot tell (strides, head size, RoPE scaling, EOS ids) comes
                              from the HF config / w2v2 arguments / flags, with the production defaults
  check_state_dict            strict key + shape check with torch's `load_state_dict` error wording
  read_hf_config              config.json + generation_config.json of --model-name (local directory)
  read_w2v2_args              architecture arguments of a fairseq checkpoint (`args` Namespace or `cfg` tree)
  w2v2_state_to_reference     fairseq `state["model"]` -> `model.speech_encoder.speech_encoder.*` keys
  preprocess_tokenizer        SpeechLlamaForCausalLM.preprocess (model/llm.py:149-190): adds the 3 + m_max
                              special tokens and records the ids the splice needs
  template_from_tokenizer     TemplateConfig for a real tokenizer (system prompt of agents/infinisst.py:229-241)
"""
from __future__ import annotations

import argparse
import ast
import json
import os
import re
from typing import Dict, Iterable, List, Mapping, Optional, Sequence, Tuple

import torch

from .config import (EncoderConfig, GenConfig, InfiniSSTConfig, LLAMA3_ROPE, LlmConfig, TemplateConfig, W2V2_CONV)

ENC = "model.speech_encoder.speech_encoder."
ADAPTER = "model.speech_encoder.length_shrink."
PROJ = "model.speech_encoder.proj."

DEFAULT_SPEECH_PATCH_TOKEN = "<sp_patch>"      # train/dataset.py:52-56
DEFAULT_SPEECH_START_TOKEN = "<sp_start>"
DEFAULT_SPEECH_END_TOKEN = "<sp_end>"
DEFAULT_LATENCY_TOKEN = "<latency_{}>"


# ------------------------------------------------------------------------------------------------ state dicts
def load_reference_state_dict(path: str) -> Dict[str, torch.Tensor]:
    """`torch.load(args.state_dict_path, map_location='cpu', weights_only=True)` (agents/infinisst.py:179).
    A Lightning checkpoint ({"state_dict": ...}) or an un-pruned .bin whose every key carries the
    LightningModule's `model.` attribute prefix (what train/prune_bin.py removes) is normalised."""
    sd = torch.load(path, map_location="cpu", weights_only=True)
    if isinstance(sd, Mapping) and "state_dict" in sd and isinstance(sd["state_dict"], Mapping):
        sd = sd["state_dict"]
    if not isinstance(sd, Mapping) or not sd:
        raise ValueError(f"{path}: not a state dict")
    keys = list(sd.keys())
    if all(k.startswith("model.") for k in keys) and "model.lm_head.weight" in sd:
        sd = {k[6:]: v for k, v in sd.items()}                      # train/prune_bin.py:7-9
    return dict(sd)


def _count(sd: Mapping[str, torch.Tensor], pattern: str) -> int:
    rx = re.compile(pattern)
    idx = {int(m.group(1)) for k in sd for m in [rx.match(k)] if m}
    if idx and idx != set(range(len(idx))):
        raise ValueError(f"non-contiguous layer indices for /{pattern}/: {sorted(idx)}")
    return len(idx)


def parse_conv_cfg(text) -> List[Tuple[int, int, int]]:
    """The reference `eval`s --length-shrink-cfg / conv_feature_layers (speech_encoder.py:128,
    fairseq wav2vec2 `conv_feature_layers`), e.g. "[(1024,2,2)] * 2".  Same grammar, no `eval`:
    int literals, tuples, lists, `*` and `+`."""
    if not isinstance(text, str):
        return [tuple(int(x) for x in t) for t in text]

    def ev(node):
        if isinstance(node, ast.Expression):
            return ev(node.body)
        if isinstance(node, ast.Constant) and isinstance(node.value, int):
            return node.value
        if isinstance(node, ast.Tuple):
            return tuple(ev(e) for e in node.elts)
        if isinstance(node, ast.List):
            return [ev(e) for e in node.elts]
        if isinstance(node, ast.BinOp) and isinstance(node.op, (ast.Mult, ast.Add)):
            a, b = ev(node.left), ev(node.right)
            return a * b if isinstance(node.op, ast.Mult) else a + b
        raise ValueError(f"unsupported expression in conv layer spec: {ast.dump(node)}")

    layers = ev(ast.parse(text.strip(), mode="eval"))
    out = []
    for t in layers:
        if not (isinstance(t, tuple) and len(t) == 3):
            raise ValueError(f"conv layer spec entries must be (dim, kernel, stride): {t!r}")
        out.append((int(t[0]), int(t[1]), int(t[2])))
    return out


def infer_config(sd: Mapping[str, torch.Tensor], *, block_size: int = 48, max_cache_size: int = 576,
                 length_shrink_cfg=None, xpos: bool = False, rope: bool = True, hf_config: Optional[dict] = None,
                 w2v2_args: Optional[dict] = None, generation_config: Optional[dict] = None,
                 name: str = "checkpoint") -> InfiniSSTConfig:
    """Architecture of a reference checkpoint.  Shapes decide: conv channels / kernels, encoder width,
    depth, FFN, head count (from `rotary_emb.freqs`), adapter, LLM width / depth / FFN / vocabulary and the
    q : kv ratio.  Flags / configs decide: conv strides, LLM head size, RMS eps, RoPE base + scaling, EOS."""
    hf = dict(hf_config or {})
    wa = dict(w2v2_args or {})

    def shape(key):
        if key not in sd:
            raise KeyError(f'Missing key(s) in state_dict: "{key}"')
        return tuple(sd[key].shape)

    # ---- speech encoder (fairseq wav2vec 2.0, layer_norm extractor mode)
    n_conv = _count(sd, re.escape(ENC) + r"feature_extractor\.conv_layers\.(\d+)\.0\.weight")
    if n_conv == 0:
        raise KeyError(f'Missing key(s) in state_dict: "{ENC}feature_extractor.conv_layers.0.0.weight"')
    if f"{ENC}feature_extractor.conv_layers.0.2.1.weight" not in sd:
        raise NotImplementedError(
            "extractor_mode='default' (GroupNorm after conv 0, wav2vec2-base) is not supported: those models are "
            "layer_norm_first=False, which the reference itself rejects (patch_speech_encoder.py:571)")
    kern = [shape(f"{ENC}feature_extractor.conv_layers.{j}.0.weight") for j in range(n_conv)]
    if "conv_feature_layers" in wa:
        conv = parse_conv_cfg(wa["conv_feature_layers"])
    elif [k[2] for k in kern] == [k for _, k, _ in W2V2_CONV]:
        conv = [(kern[j][0], W2V2_CONV[j][1], W2V2_CONV[j][2]) for j in range(n_conv)]
    else:
        raise ValueError("conv strides cannot be read off a state dict: pass the w2v2 checkpoint's arguments "
                         "(conv_feature_layers)")
    if len(conv) != n_conv or any((c, k) != (kern[j][0], kern[j][2]) for j, (c, k, _) in enumerate(conv)):
        raise ValueError(f"conv_feature_layers {conv} does not match the checkpoint's conv kernels {kern}")
    embed_dim, conv_dim = shape(ENC + "post_extract_proj.weight")
    if conv_dim != conv[-1][0]:
        raise ValueError("post_extract_proj input width != last conv width")
    n_enc = _count(sd, re.escape(ENC) + r"encoder\.layers\.(\d+)\.fc1\.weight")
    ffn_dim = shape(ENC + "encoder.layers.0.fc1.weight")[0]
    fkey = ENC + "encoder.layers.0.self_attn.rotary_emb.freqs"
    if fkey in sd:
        head_dim = 2 * sd[fkey].numel()                                   # RotaryEmbedding(embed_dim // heads)
    else:
        head_dim = embed_dim // int(wa.get("encoder_attention_heads", 16))
    if embed_dim % head_dim:
        raise ValueError("encoder width is not a multiple of the rotary head size")
    n_ad = _count(sd, re.escape(ADAPTER) + r"conv_layers\.(\d+)\.0\.weight")
    ad_kern = [shape(f"{ADAPTER}conv_layers.{j}.0.weight") for j in range(n_ad)]
    if length_shrink_cfg is not None:
        adapter = parse_conv_cfg(length_shrink_cfg)
    else:
        adapter = [(k[0], k[2], k[2]) for k in ad_kern]                   # production: kernel == stride == 2
    if len(adapter) != n_ad or any((c, k) != (ad_kern[j][0], ad_kern[j][2]) for j, (c, k, _) in enumerate(adapter)):
        raise ValueError(f"--length-shrink-cfg {adapter} does not match the checkpoint's adapter convs {ad_kern}")
    llm_dim = shape(PROJ + "weight")[0]
    enc = EncoderConfig(conv_layers=conv, embed_dim=embed_dim, ffn_dim=ffn_dim, heads=embed_dim // head_dim,
                        layers=n_enc, block_size=block_size, max_cache_size=max_cache_size, adapter_layers=adapter,
                        llm_dim=llm_dim, rope=bool(rope), xpos=bool(xpos))

    # ---- Llama
    vocab, hidden = shape("model.embed_tokens.weight")
    if hidden != llm_dim:
        raise ValueError("speech projection width != LLM width")
    n_llm = _count(sd, r"model\.layers\.(\d+)\.mlp\.gate_proj\.weight")
    ffn = shape("model.layers.0.mlp.gate_proj.weight")[0]
    q_out = shape("model.layers.0.self_attn.q_proj.weight")[0]
    kv_out = shape("model.layers.0.self_attn.k_proj.weight")[0]
    if "head_dim" in hf and hf["head_dim"]:
        hd = int(hf["head_dim"])
    elif "num_attention_heads" in hf:
        hd = q_out // int(hf["num_attention_heads"])
    else:
        hd = 128
    if q_out % hd or kv_out % hd:
        raise ValueError(f"q/k projection widths {q_out}/{kv_out} are not multiples of head_dim {hd}")
    scaling = hf.get("rope_scaling", dict(LLAMA3_ROPE) if not hf else None)
    if scaling is not None:
        kind = scaling.get("rope_type", scaling.get("type", "llama3"))
        if kind != "llama3":
            raise NotImplementedError(f"rope_scaling type {kind!r} (only llama3 and none are built)")
        scaling = {k: float(scaling[k]) for k in LLAMA3_ROPE}
    llm = LlmConfig(hidden=hidden, layers=n_llm, heads=q_out // hd, kv_heads=kv_out // hd, head_dim=hd, ffn=ffn,
                    vocab=vocab, rms_eps=float(hf.get("rms_norm_eps", 1e-5)),
                    rope_theta=float(hf.get("rope_theta", 500000.0)), rope_scaling=scaling)
    gen = GenConfig()
    g = dict(generation_config or {})
    if "eos_token_id" in g:
        e = g["eos_token_id"]
        gen.eos_token_ids = [int(e)] if isinstance(e, int) else [int(x) for x in e]
    cfg = InfiniSSTConfig(enc=enc, llm=llm, tpl=TemplateConfig(speech_tokens_per_chunk=block_size // 4), gen=gen,
                          name=name)
    check_state_dict(sd, cfg)
    return cfg


def expected_shapes(cfg: InfiniSSTConfig) -> Dict[str, Tuple[int, ...]]:
    """Key -> shape of the reference module tree for this architecture (SURVEY §8b 'Weights')."""
    e, l = cfg.enc, cfg.llm
    out: Dict[str, Tuple[int, ...]] = {}
    cin = 1
    for j, (c, k, _) in enumerate(e.conv_layers):
        p = f"{ENC}feature_extractor.conv_layers.{j}."
        out[p + "0.weight"], out[p + "0.bias"] = (c, cin, k), (c,)
        out[p + "2.1.weight"], out[p + "2.1.bias"] = (c,), (c,)
        cin = c
    out[ENC + "layer_norm.weight"], out[ENC + "layer_norm.bias"] = (cin,), (cin,)
    out[ENC + "post_extract_proj.weight"], out[ENC + "post_extract_proj.bias"] = (e.embed_dim, cin), (e.embed_dim,)
    D, F = e.embed_dim, e.ffn_dim
    for i in range(e.layers):
        p = f"{ENC}encoder.layers.{i}."
        for n in ("q_proj", "k_proj", "v_proj", "out_proj"):
            out[p + f"self_attn.{n}.weight"], out[p + f"self_attn.{n}.bias"] = (D, D), (D,)
        out[p + "self_attn.rotary_emb.freqs"] = (e.head_dim // 2,)
        for n in ("self_attn_layer_norm", "final_layer_norm"):
            out[p + n + ".weight"], out[p + n + ".bias"] = (D,), (D,)
        out[p + "fc1.weight"], out[p + "fc1.bias"] = (F, D), (F,)
        out[p + "fc2.weight"], out[p + "fc2.bias"] = (D, F), (D,)
    out[ENC + "encoder.layer_norm.weight"], out[ENC + "encoder.layer_norm.bias"] = (D,), (D,)
    cin = D
    for j, (c, k, _) in enumerate(e.adapter_layers):
        p = f"{ADAPTER}conv_layers.{j}."
        out[p + "0.weight"] = (c, cin, k)
        out[p + "2.1.weight"], out[p + "2.1.bias"] = (c,), (c,)
        cin = c
    out[PROJ + "weight"], out[PROJ + "bias"] = (e.llm_dim, cin), (e.llm_dim,)
    out["model.embed_tokens.weight"] = (l.vocab, l.hidden)
    out["lm_head.weight"] = (l.vocab, l.hidden)
    out["model.norm.weight"] = (l.hidden,)
    for i in range(l.layers):
        p = f"model.layers.{i}."
        out[p + "self_attn.q_proj.weight"] = (l.heads * l.head_dim, l.hidden)
        out[p + "self_attn.k_proj.weight"] = (l.kv_heads * l.head_dim, l.hidden)
        out[p + "self_attn.v_proj.weight"] = (l.kv_heads * l.head_dim, l.hidden)
        out[p + "self_attn.o_proj.weight"] = (l.hidden, l.heads * l.head_dim)
        out[p + "mlp.gate_proj.weight"], out[p + "mlp.up_proj.weight"] = (l.ffn, l.hidden), (l.ffn, l.hidden)
        out[p + "mlp.down_proj.weight"] = (l.hidden, l.ffn)
        out[p + "input_layernorm.weight"] = (l.hidden,)
        out[p + "post_attention_layernorm.weight"] = (l.hidden,)
    return out


# tensors the reference's module tree holds but the per-chunk step never reads
_UNUSED = re.compile(r"(" + re.escape(ENC) + r"(mask_emb|encoder\.pos_conv\..*|quantizer\..*|project_q\..*|final_proj\..*|"
                     r"encoder\.layers\.\d+\.self_attn\.rotary_emb\.(scale|dummy|cached_.*))"
                     r"|model\.rotary_emb\.inv_freq|model\.layers\.\d+\.self_attn\.rotary_emb\.inv_freq)$")


def check_state_dict(sd: Mapping[str, torch.Tensor], cfg: InfiniSSTConfig) -> None:
    """Strict `load_state_dict` (agents/infinisst.py:180): every expected tensor present with the expected
    shape, nothing unknown.  Raises RuntimeError with torch's wording."""
    exp = expected_shapes(cfg)
    errs = []
    missing = [k for k in exp if k not in sd]
    unexpected = [k for k in sd if k not in exp and not _UNUSED.match(k)]
    if missing:
        errs.append("Missing key(s) in state_dict: " + ", ".join(f'"{k}"' for k in missing[:12]) +
                    (f" ... ({len(missing)} in total)" if len(missing) > 12 else ""))
    if unexpected:
        errs.append("Unexpected key(s) in state_dict: " + ", ".join(f'"{k}"' for k in unexpected[:12]) +
                    (f" ... ({len(unexpected)} in total)" if len(unexpected) > 12 else ""))
    for k, s in exp.items():
        if k in sd and tuple(sd[k].shape) != s:
            errs.append(f"size mismatch for {k}: copying a param with shape {tuple(sd[k].shape)} from checkpoint, "
                        f"the shape in current model is {s}.")
    if errs:
        raise RuntimeError("Error(s) in loading state_dict for SpeechLlamaForCausalLM:\n\t" + "\n\t".join(errs))


# ------------------------------------------------------------------------------------------------ side files
def read_hf_config(model_name: Optional[str]) -> Tuple[Optional[dict], Optional[dict]]:
    """config.json / generation_config.json of a local --model-name directory (no hub access here).
    Returns (None, None) when `model_name` is not a directory: the Llama-3.1-8B defaults then apply."""
    if not model_name or not os.path.isdir(model_name):
        return None, None
    out = []
    for fn in ("config.json", "generation_config.json"):
        p = os.path.join(model_name, fn)
        if os.path.isfile(p):
            with open(p) as f:
                out.append(json.load(f))
        else:
            out.append(None)
    return out[0], out[1]


_W2V2_FIELDS = ("conv_feature_layers", "encoder_layers", "encoder_embed_dim", "encoder_ffn_embed_dim",
                "encoder_attention_heads", "extractor_mode", "layer_norm_first", "conv_bias")


def read_w2v2_args(path: str) -> dict:
    """Architecture arguments of a fairseq wav2vec 2.0 checkpoint (speech_encoder.py:147-172: `state["args"]`
    for SSL models, `state["cfg"]["model"]["w2v_args"]["model"]` for CTC fine-tuned ones).  Only plain
    containers and argparse.Namespace are unpickled (weights_only=True)."""
    with torch.serialization.safe_globals([argparse.Namespace]):
        state = torch.load(path, map_location="cpu", weights_only=True)
    node = state.get("args")
    if node is None:
        node = state["cfg"]["model"]
        if "w2v_args" in node and node["w2v_args"] is not None:
            w = node["w2v_args"]
            node = w["model"] if isinstance(w, Mapping) else getattr(w, "model", w)
    get = (lambda k: node.get(k)) if isinstance(node, Mapping) else (lambda k: getattr(node, k, None))
    out = {k: get(k) for k in _W2V2_FIELDS if get(k) is not None}
    if out.get("extractor_mode", "layer_norm") != "layer_norm" or out.get("layer_norm_first") is False:
        raise NotImplementedError("only layer_norm-mode, layer_norm_first wav2vec 2.0 models run through the "
                                  "reference's patched encoder (patch_speech_encoder.py:571)")
    return out


def w2v2_state_to_reference(model_state: Mapping[str, torch.Tensor], ctc_finetuned: bool = False) -> Dict[str, torch.Tensor]:
    """fairseq `state["model"]` -> reference keys (speech_encoder.py:150-170: the SSL model is loaded as is,
    the CTC one drops `w2v_encoder.` / `w2v_model.` and its `proj`)."""
    out = {}
    for k, v in model_state.items():
        if ctc_finetuned:
            k = k.replace("w2v_encoder.", "")
            if k.startswith("proj"):
                continue
            k = k.replace("w2v_model.", "", 1)
        out[ENC + k] = v
    return out


# ------------------------------------------------------------------------------------------------ tokenizer
def preprocess_tokenizer(tokenizer, cfg: InfiniSSTConfig, max_multiplier: int = 4) -> InfiniSSTConfig:
    """SpeechLlamaForCausalLM.preprocess(tokenizer, max_multiplier, resize=False) (model/llm.py:149-190;
    agents/infinisst.py:177): add `<sp_patch>`, `<sp_start>`, `<sp_end>`, `<latency_1..m>` as special tokens
    and store the ids the splice scans for on the config.  The embedding tables come resized from the
    checkpoint, so their row count must equal len(tokenizer)."""
    new = [DEFAULT_SPEECH_PATCH_TOKEN, DEFAULT_SPEECH_START_TOKEN, DEFAULT_SPEECH_END_TOKEN] + \
          [DEFAULT_LATENCY_TOKEN.format(i) for i in range(1, max_multiplier + 1)]
    tokenizer.add_tokens(new, special_tokens=True)
    if getattr(tokenizer, "pad_token_id", None) is None and getattr(tokenizer, "pad_token", None):
        tokenizer.add_tokens([tokenizer.pad_token], special_tokens=True)
    if len(tokenizer) != cfg.llm.vocab:
        raise ValueError(f"tokenizer has {len(tokenizer)} tokens after adding the speech tokens, the checkpoint's "
                         f"embedding table has {cfg.llm.vocab} rows (max_latency_multiplier mismatch?)")
    ids = tokenizer.convert_tokens_to_ids
    l, t = cfg.llm, cfg.tpl
    l.sp_patch_token_id = t.sp_patch_id = int(ids(DEFAULT_SPEECH_PATCH_TOKEN))
    l.user_token_id = t.user_token_id = int(ids("user"))
    l.assist_token_id = t.assist_token_id = int(ids("assistant"))
    l.start_header_id = t.start_header_id = int(ids("<|start_header_id|>"))
    for attr, tok in (("end_header_id", "<|end_header_id|>"), ("eot_id", "<|eot_id|>")):
        v = ids(tok)
        if v is not None:
            setattr(t, attr, int(v))
    if getattr(tokenizer, "pad_token_id", None) is not None:
        cfg.gen.pad_token_id = int(tokenizer.pad_token_id)
    return cfg


def template_from_tokenizer(tokenizer, cfg: InfiniSSTConfig, source_lang: str, target_lang: str,
                            latency_multiplier: int) -> InfiniSSTConfig:
    """System-turn ids exactly as the agent's first chunk forms them (agents/infinisst.py:229-241), stored on
    the template so host-side tools (bench, runner) can build prompts without the tokenizer."""
    latency_token = DEFAULT_LATENCY_TOKEN.format(latency_multiplier)
    messages = [{"role": "system", "content": f"Translate the following speech from {source_lang} to "
                                              f"{target_lang} with latency {latency_token}."}]
    ids = tokenizer.apply_chat_template([messages], return_tensors="pt", padding=True, truncation=False,
                                        add_special_tokens=False)
    ids = ids["input_ids"] if isinstance(ids, Mapping) else ids
    cfg.tpl.system_ids = [int(x) for x in ids[0].tolist()]
    cfg.tpl.speech_tokens_per_chunk = cfg.enc.block_size // 4
    return cfg


def load_checkpoint(state_dict_path: str, *, model_name: Optional[str] = None, w2v2_path: Optional[str] = None,
                    block_size: int = 48, max_cache_size: int = 576, length_shrink_cfg=None, xpos: bool = False,
                    rope: bool = True) -> Tuple[InfiniSSTConfig, Dict[str, torch.Tensor]]:
    """The three artefacts of `load_model` -> (config, reference-layout state dict)."""
    sd = load_reference_state_dict(state_dict_path)
    hf, gen = read_hf_config(model_name)
    wa = read_w2v2_args(w2v2_path) if w2v2_path and os.path.isfile(w2v2_path) else None
    cfg = infer_config(sd, block_size=block_size, max_cache_size=max_cache_size, length_shrink_cfg=length_shrink_cfg,
                       xpos=xpos, rope=rope, hf_config=hf, w2v2_args=wa, generation_config=gen,
                       name=os.path.basename(os.path.normpath(model_name)) if model_name else "checkpoint")
    return cfg, sd
