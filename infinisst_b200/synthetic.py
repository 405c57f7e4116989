"""Synthetic weights and audio (there is no network for checkpoints or datasets).

Weights are produced in the reference's own state-dict key layout (the single
`pytorch_model.bin` the agent loads, agents/infinisst.py:179-180; key nesting from
model/llm.py:133-135 and model/speech_encoder.py:111-121), so the loader that
ingests them is the loader a real checkpoint would go through.
"""
from __future__ import annotations

import math
from typing import Dict, Iterator, Tuple

import torch

from .config import InfiniSSTConfig

ENC = "model.speech_encoder.speech_encoder."
SPE = "model.speech_encoder."


def weight_specs(cfg: InfiniSSTConfig) -> Iterator[Tuple[str, Tuple[int, ...], str, float]]:
    """Yield (key, shape, kind, std) for every tensor on the hot path.
    kind: 'normal' | 'ones' | 'zeros' | 'rope_freqs'."""
    e, l = cfg.enc, cfg.llm
    c_in = 1
    for j, (c, k, _s) in enumerate(e.conv_layers):
        p = f"{ENC}feature_extractor.conv_layers.{j}."
        yield p + "0.weight", (c, c_in, k), "normal", math.sqrt(2.0 / (c_in * k))
        yield p + "0.bias", (c,), "normal", 0.02
        yield p + "2.1.weight", (c,), "ones", 0.1
        yield p + "2.1.bias", (c,), "normal", 0.02
        c_in = c
    yield ENC + "layer_norm.weight", (c_in,), "ones", 0.1
    yield ENC + "layer_norm.bias", (c_in,), "normal", 0.02
    yield ENC + "post_extract_proj.weight", (e.embed_dim, c_in), "normal", 0.02 * math.sqrt(1024 / c_in)
    yield ENC + "post_extract_proj.bias", (e.embed_dim,), "normal", 0.02
    d = e.embed_dim
    wstd = 0.02 * math.sqrt(1024 / d)
    for i in range(e.layers):
        p = f"{ENC}encoder.layers.{i}."
        for n in ("q_proj", "k_proj", "v_proj", "out_proj"):
            yield p + f"self_attn.{n}.weight", (d, d), "normal", wstd * (2.0 if n in ("q_proj", "k_proj") else 1.0)
            yield p + f"self_attn.{n}.bias", (d,), "normal", 0.02
        yield p + "self_attn.rotary_emb.freqs", (e.head_dim // 2,), "rope_freqs", e.rope_theta
        yield p + "self_attn_layer_norm.weight", (d,), "ones", 0.1
        yield p + "self_attn_layer_norm.bias", (d,), "normal", 0.02
        yield p + "fc1.weight", (e.ffn_dim, d), "normal", wstd
        yield p + "fc1.bias", (e.ffn_dim,), "normal", 0.02
        yield p + "fc2.weight", (d, e.ffn_dim), "normal", wstd
        yield p + "fc2.bias", (d,), "normal", 0.02
        yield p + "final_layer_norm.weight", (d,), "ones", 0.1
        yield p + "final_layer_norm.bias", (d,), "normal", 0.02
    yield ENC + "encoder.layer_norm.weight", (d,), "ones", 0.1
    yield ENC + "encoder.layer_norm.bias", (d,), "normal", 0.02
    c_in = d
    for j, (c, k, _s) in enumerate(e.adapter_layers):
        p = f"{SPE}length_shrink.conv_layers.{j}."
        yield p + "0.weight", (c, c_in, k), "normal", math.sqrt(2.0 / (c_in * k))   # kaiming, speech_encoder.py:37
        yield p + "2.1.weight", (c,), "ones", 0.1
        yield p + "2.1.bias", (c,), "normal", 0.02
        c_in = c
    yield SPE + "proj.weight", (e.llm_dim, c_in), "normal", 0.02 * math.sqrt(1024 / c_in)
    yield SPE + "proj.bias", (e.llm_dim,), "normal", 0.02
    h = l.hidden
    lstd = 0.02 * math.sqrt(4096 / h)
    yield "model.embed_tokens.weight", (l.vocab, h), "normal", 1.0
    for i in range(l.layers):
        p = f"model.layers.{i}."
        yield p + "self_attn.q_proj.weight", (l.heads * l.head_dim, h), "normal", lstd
        yield p + "self_attn.k_proj.weight", (l.kv_heads * l.head_dim, h), "normal", lstd
        yield p + "self_attn.v_proj.weight", (l.kv_heads * l.head_dim, h), "normal", lstd
        yield p + "self_attn.o_proj.weight", (h, l.heads * l.head_dim), "normal", lstd
        yield p + "mlp.gate_proj.weight", (l.ffn, h), "normal", lstd
        yield p + "mlp.up_proj.weight", (l.ffn, h), "normal", lstd
        yield p + "mlp.down_proj.weight", (h, l.ffn), "normal", lstd * math.sqrt(h / l.ffn)
        yield p + "input_layernorm.weight", (h,), "ones", 0.1
        yield p + "post_attention_layernorm.weight", (h,), "ones", 0.1
    yield "model.norm.weight", (h,), "ones", 0.1
    yield "lm_head.weight", (l.vocab, h), "normal", lstd * 4.0   # sharpened head: healthier arg-max margins (SURVEY §7 hard part 2)


def make_state_dict(cfg: InfiniSSTConfig, seed: int = 0, device: str = "cpu",
                    dtype: torch.dtype = torch.float32) -> Dict[str, torch.Tensor]:
    """Random-init weights, deterministic in (seed, key order).  On CUDA each tensor is
    sampled on the device (8B parameters would take minutes on the host)."""
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}
    for key, shape, kind, std in weight_specs(cfg):
        if kind == "normal":
            t = torch.empty(shape, device=device, dtype=torch.float32 if device == "cpu" else dtype)
            t.normal_(0.0, std, generator=g)
        elif kind == "ones":
            t = torch.ones(shape, device=device, dtype=torch.float32)
            t += std * torch.randn(shape, device=device, dtype=torch.float32, generator=g)
        elif kind == "zeros":
            t = torch.zeros(shape, device=device, dtype=torch.float32)
        else:  # rope_freqs: rotary_embedding_torch RotaryEmbedding(dim).freqs
            n = shape[0]
            t = 1.0 / (std ** (torch.arange(0, 2 * n, 2, device=device, dtype=torch.float32) / (2 * n)))
        sd[key] = t.to(dtype)
    return sd


def make_audio(seconds: float, seed: int = 998244353, sample_rate: int = 16000) -> torch.Tensor:
    """0.1*N(0,1) + a 220 Hz tone, seeded with the reference's own seed
    (agents/infinisst.py:74); float32 [n] on the host (SimulEval hands the agent host floats)."""
    g = torch.Generator()
    g.manual_seed(seed)
    n = int(round(seconds * sample_rate))
    t = torch.arange(n, dtype=torch.float32) / sample_rate
    return 0.1 * torch.randn(n, generator=g) + 0.05 * torch.sin(2 * math.pi * 220.0 * t)
