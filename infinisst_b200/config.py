"""Configuration of the per-chunk streaming step.

Field names follow the reference's flags and hyper-parameters:
  encoder flags   agents/options.py:1-41  (block_size, max_cache_size, xpos, rope, length_shrink_cfg)
  generation      agents/options.py:43-108, agents/infinisst.py:185-198
  production values scripts/infer/infinisst.sh:41-87 (block 48, cache 576, "[(1024,2,2)] * 2",
                  xpos 0, max_llm_cache_size 1000, always-cache-system-prompt, no-repeat 5/100, penalty 1.2)
  model dims      wav2vec2-large (wav2vec_vox_960h_pl.pt) and Llama-3.1-8B-Instruct
"""
from __future__ import annotations

from dataclasses import dataclass, field, asdict
from typing import Dict, List, Optional, Tuple

W2V2_CONV = [(512, 10, 5)] + [(512, 3, 2)] * 4 + [(512, 2, 2)] * 2
LLAMA3_ROPE = {"factor": 8.0, "low_freq_factor": 1.0, "high_freq_factor": 4.0,
               "original_max_position_embeddings": 8192}


@dataclass
class EncoderConfig:
    conv_layers: List[Tuple[int, int, int]] = field(default_factory=lambda: list(W2V2_CONV))
    embed_dim: int = 1024
    ffn_dim: int = 4096
    heads: int = 16
    layers: int = 24
    block_size: int = 48            # --block-size
    max_cache_size: int = 576       # --max-cache-size
    adapter_layers: List[Tuple[int, int, int]] = field(default_factory=lambda: [(1024, 2, 2)] * 2)  # --length-shrink-cfg
    llm_dim: int = 4096
    rope: bool = True               # --rope 1
    xpos: bool = False              # --xpos 0
    rope_theta: float = 10000.0
    rope_angle_dtype: str = "fp32"  # see SURVEY App. A.2 (dtype hazard); 'fp32' | 'model'

    @property
    def head_dim(self) -> int:
        return self.embed_dim // self.heads

    @property
    def conv_dim(self) -> int:
        return self.conv_layers[-1][0]


@dataclass
class LlmConfig:
    hidden: int = 4096
    layers: int = 32
    heads: int = 32
    kv_heads: int = 8
    head_dim: int = 128
    ffn: int = 14336
    vocab: int = 128256 + 7         # llm.py:150-167 adds 7 special tokens
    rms_eps: float = 1e-5
    rope_theta: float = 500000.0
    rope_scaling: Optional[Dict[str, float]] = field(default_factory=lambda: dict(LLAMA3_ROPE))
    # ids stored on the HF config by `preprocess` (llm.py:176-190)
    user_token_id: int = 882
    assist_token_id: int = 78191
    start_header_id: int = 128006
    sp_patch_token_id: int = 128256


@dataclass
class TemplateConfig:
    """Token layout of one turn (agents/infinisst.py:225-268).  With a real
    tokenizer these come from `apply_chat_template`; without one (this container has no Llama
    tokenizer) the synthetic template below reproduces the structure with the real special ids."""
    start_header_id: int = 128006
    end_header_id: int = 128007
    eot_id: int = 128009
    user_token_id: int = 882
    assist_token_id: int = 78191
    nl_id: int = 271                # "\n\n"
    sp_patch_id: int = 128256
    speech_tokens_per_chunk: int = 12   # block_size // 4
    system_ids: List[int] = field(default_factory=list)


@dataclass
class GenConfig:
    latency_multiplier: int = 1
    max_new_tokens: int = 10              # 10 * multiplier (agents/infinisst.py:125-128)
    beam: int = 1                         # greedy oracle; beam>1 is §8f "next"
    no_repeat_ngram_lookback: int = 100
    no_repeat_ngram_size: int = 5
    repetition_penalty: float = 1.2
    suppress_tokens: List[int] = field(default_factory=list)
    eos_token_ids: List[int] = field(default_factory=lambda: [128001, 128008, 128009])
    pad_token_id: int = 128004            # <|finetune_right_pad_id|> (agents/infinisst.py:140)
    max_llm_cache_size: int = 1000
    always_cache_system_prompt: bool = True


@dataclass
class InfiniSSTConfig:
    enc: EncoderConfig = field(default_factory=EncoderConfig)
    llm: LlmConfig = field(default_factory=LlmConfig)
    tpl: TemplateConfig = field(default_factory=TemplateConfig)
    gen: GenConfig = field(default_factory=GenConfig)
    name: str = "production"

    def to_dict(self) -> dict:
        return asdict(self)


def _system_ids(n: int, bos: int, start_header: int, end_header: int, eot: int, filler0: int) -> List[int]:
    """Synthetic system turn with the real structure: BOS, <start_header>, 'system',
    <end_header>, "\n\n", text..., <eot>.  `n` tokens in total."""
    head = [bos, start_header, filler0, end_header, filler0 + 1]
    body = [filler0 + 2 + (7 * i) % 97 for i in range(n - len(head) - 1)]
    return head + body + [eot]


def production_config() -> InfiniSSTConfig:
    """BASELINE.json configs[1..4]: wav2vec2-large + Llama-3.1-8B, m=1."""
    cfg = InfiniSSTConfig()
    cfg.tpl.system_ids = _system_ids(40, 128000, 128006, 128007, 128009, 1000)
    return cfg


def tiny_config(max_cache_size: int = 576, max_llm_cache_size: int = 1000) -> InfiniSSTConfig:
    """BASELINE.json configs[0]: 2-layer wav2vec2 + 2-layer Llama-style LLM.

    Head sizes are kept at the production values (64 for the encoder, 128 and 4:1 GQA for
    the LLM) so the tiny parity run exercises the same kernels as production; widths,
    depths and the vocabulary are shrunk.  The vocabulary is deliberately odd (like 128263)."""
    C = 64
    enc = EncoderConfig(conv_layers=[(C, 10, 5)] + [(C, 3, 2)] * 4 + [(C, 2, 2)] * 2,
                        embed_dim=128, ffn_dim=256, heads=2, layers=2, block_size=48,
                        max_cache_size=max_cache_size, adapter_layers=[(128, 2, 2)] * 2, llm_dim=512)
    V = 512
    llm = LlmConfig(hidden=512, layers=2, heads=4, kv_heads=1, head_dim=128, ffn=768, vocab=V + 7,
                    user_token_id=300, assist_token_id=301, start_header_id=V - 6, sp_patch_token_id=V)
    tpl = TemplateConfig(start_header_id=V - 6, end_header_id=V - 5, eot_id=V - 3, user_token_id=300,
                         assist_token_id=301, nl_id=271, sp_patch_id=V, speech_tokens_per_chunk=12,
                         system_ids=_system_ids(40, V - 8, V - 6, V - 5, V - 3, 310))
    gen = GenConfig(eos_token_ids=[V - 7, V - 4, V - 3], pad_token_id=V - 2,
                    max_llm_cache_size=max_llm_cache_size)
    return InfiniSSTConfig(enc=enc, llm=llm, tpl=tpl, gen=gen, name="tiny")
