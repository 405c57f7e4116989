"""Stand-ins for the third-party packages the reference imports but this image lacks, so that the
reference's OWN hot-path code (under /root/reference, imported unmodified) can be executed here to
generate fixtures that pin the oracle (tests/golden/make_ref_pins.py).

What is a stand-in and what is the reference's code
---------------------------------------------------
Executed unmodified from /root/reference (this is what the fixtures pin):
  agents/infinisst.py        InfiniSST.policy / _prepare_speech / _prepare_inputs / eviction / drop-last
  model/llm.py               SpeechLlamaModel.forward (encode-once, splice), SpeechLlamaForCausalLM.forward,
                             prepare_inputs_for_generation
  model/speech_encoder.py    SpeechEncoderW2V2RoPE.__init__ / set_blocksize / encode_speech /
                             _get_feat_extract_output_lengths, ConvFeatureExtractionModel (adapter), caches
  model/patches/patch_speech_encoder.py   patch_w2v2, masks, uni_w2v2_forward, encoder extract_features,
                             uni_self_attn_forward, uni_mha_init, uni_mha_forward
  model/patches/patch_llm.py llama_sdpa_attention_new_forward (un-rotated-K cache)
  train/dataset.py           the DEFAULT_* token constants

Stand-ins written here from the published behaviour of the absent packages (NOT
reference code - the same third-party semantics the oracle has to restate anyway):
  fairseq 0.12.2             Wav2Vec2Model / TransformerEncoder / TransformerSentenceEncoderLayer /
                             MultiheadAttention *constructors and containers only* (their forward methods are
                             replaced by the reference's patch_w2v2), ConvFeatureExtractionModel (layer_norm
                             mode), Fp32LayerNorm, TransposeLast, gelu, utils.softmax, pad_to_multiple, ...
  rotary_embedding_torch     RotaryEmbedding(dim, use_xpos).rotate_queries_with_cached_keys (+ xPos get_scale)
  simuleval                  SpeechToTextAgent / AgentStates / ReadAction / WriteAction / entrypoint
  lightning                  LightningModule = nn.Module
  transformers 4.47 names    LlamaSdpaAttention / LlamaFlashAttention2 (placeholders so patch_llm imports),
                             DynamicCache with the 4.47 list API (`key_cache`, `value_cache`), and HF's greedy
                             `_sample` loop (patch_hf.py needs 4.47 internals; greedy per SURVEY App. C).
  The Llama blocks themselves (RMSNorm, MLP, rotary tables, apply_rotary_pos_emb, repeat_kv, logits
  processors) are the real transformers 5.5 modules of this image.
Everything else the reference imports (wandb, jieba, soundfile, deepspeed, ...) is an empty auto-stub.
"""
from __future__ import annotations

import importlib.abc
import importlib.machinery
import math
import sys
import types
from typing import List, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

REFERENCE = "/root/reference"


# ------------------------------------------------------------------------------------------------
# generic auto-stub: any attribute is a dummy class, any submodule imports
# ------------------------------------------------------------------------------------------------
class _Dummy:
    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return None

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Dummy()


class _StubModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        cls = type(name, (_Dummy,), {})
        setattr(self, name, cls)
        return cls


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def __init__(self, roots):
        self.roots = set(roots)

    def find_spec(self, fullname, path, target=None):
        if fullname.split(".")[0] in self.roots and fullname not in sys.modules:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _StubModule(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        pass


def _mod(name: str, **attrs) -> types.ModuleType:
    m = _StubModule(name)
    m.__path__ = []
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    parent, _, leaf = name.rpartition(".")
    if parent and parent in sys.modules:
        setattr(sys.modules[parent], leaf, m)
    return m


# ------------------------------------------------------------------------------------------------
# fairseq 0.12.2 stand-ins
# ------------------------------------------------------------------------------------------------
class TransposeLast(nn.Module):
    def forward(self, x):
        return x.transpose(-2, -1)


class Fp32LayerNorm(nn.LayerNorm):
    def forward(self, x):
        out = F.layer_norm(x.float(), self.normalized_shape,
                           self.weight.float() if self.weight is not None else None,
                           self.bias.float() if self.bias is not None else None, self.eps)
        return out.type_as(x)


class FairseqDropout(nn.Module):
    def __init__(self, p, module_name=None):
        super().__init__()
        self.p = p

    def forward(self, x, inplace: bool = False):
        return F.dropout(x, p=self.p, training=True, inplace=inplace) if (self.p > 0 and self.training) else x


def quant_noise(module, p, block_size):
    assert p <= 0
    return module


class GradMultiply(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, scale):
        ctx.scale = scale
        return x.new(x)

    @staticmethod
    def backward(ctx, grad):
        return grad * ctx.scale, None


def fs_gelu(x):
    return F.gelu(x.float()).type_as(x)


def fs_softmax(x, dim: int, onnx_trace: bool = False):
    return F.softmax(x, dim=dim, dtype=torch.float32)


def fs_index_put(tensor, indices, value):
    tensor[indices] = value
    return tensor


def fs_pad_to_multiple(x, multiple, dim=-1, value=0):
    if x is None:
        return None, 0
    tsz = x.size(dim)
    m = tsz / multiple
    remainder = math.ceil(m) * multiple - tsz
    if m.is_integer():
        return x, 0
    pad_offset = (0,) * (-1 - dim) * 2
    return F.pad(x, (*pad_offset, 0, remainder), value=value), remainder


class FsConvFeatureExtractionModel(nn.Module):
    """fairseq ConvFeatureExtractionModel, mode='layer_norm' (wav2vec2-large)."""

    def __init__(self, conv_layers, dropout=0.0, mode="layer_norm", conv_bias=True):
        super().__init__()
        assert mode == "layer_norm"
        in_d = 1
        self.conv_layers = nn.ModuleList()
        for dim, k, stride in conv_layers:
            self.conv_layers.append(nn.Sequential(
                nn.Conv1d(in_d, dim, k, stride=stride, bias=conv_bias),
                nn.Dropout(p=dropout),
                nn.Sequential(TransposeLast(), Fp32LayerNorm(dim, elementwise_affine=True), TransposeLast()),
                nn.GELU()))
            in_d = dim

    def forward(self, x):
        x = x.unsqueeze(1)
        for conv in self.conv_layers:
            x = conv(x)
        return x


class MultiheadAttention(nn.Module):
    """Container only: __init__ and forward are installed by the reference's patch_w2v2."""

    def reset_parameters(self):
        pass

    def apply_sparse_mask(self, attn_weights, tgt_len: int, src_len: int, bsz: int):
        return attn_weights

    @staticmethod
    def _append_prev_key_padding_mask(key_padding_mask, prev_key_padding_mask, batch_size, src_len, static_kv):
        if prev_key_padding_mask is not None and static_kv:
            return prev_key_padding_mask
        if prev_key_padding_mask is not None and key_padding_mask is not None:
            return torch.cat([prev_key_padding_mask.float(), key_padding_mask.float()], dim=1)
        if prev_key_padding_mask is not None or key_padding_mask is not None:
            raise NotImplementedError("padding masks are not used on the streaming path")
        return None


class TransformerSentenceEncoderLayer(nn.Module):
    def __init__(self, embedding_dim, ffn_embedding_dim, num_attention_heads, layer_norm_first=True):
        super().__init__()
        self.embedding_dim = embedding_dim
        self.activation_fn = fs_gelu
        self.self_attn = MultiheadAttention(embedding_dim, num_attention_heads, dropout=0.0, self_attention=True)
        self.dropout1 = nn.Dropout(0.0)
        self.dropout2 = nn.Dropout(0.0)
        self.dropout3 = nn.Dropout(0.0)
        self.layer_norm_first = layer_norm_first
        self.self_attn_layer_norm = nn.LayerNorm(embedding_dim)
        self.fc1 = nn.Linear(embedding_dim, ffn_embedding_dim)
        self.fc2 = nn.Linear(ffn_embedding_dim, embedding_dim)
        self.final_layer_norm = nn.LayerNorm(embedding_dim)


class TransformerEncoder(nn.Module):
    def __init__(self, embed_dim, ffn_dim, heads, layers):
        super().__init__()
        self.dropout = 0.0
        self.embedding_dim = embed_dim
        self.required_seq_len_multiple = 2
        self.layers = nn.ModuleList(
            [TransformerSentenceEncoderLayer(embed_dim, ffn_dim, heads) for _ in range(layers)])
        self.layer_norm_first = True
        self.layer_norm = nn.LayerNorm(embed_dim)
        self.layerdrop = 0.0


class Wav2Vec2Model(nn.Module):
    def __init__(self, conv_layers, embed_dim, ffn_dim, heads, layers):
        super().__init__()
        self.conv_layers_cfg = list(conv_layers)
        self.embed = conv_layers[-1][0]
        self.feature_extractor = FsConvFeatureExtractionModel(conv_layers, 0.0, "layer_norm", True)
        self.post_extract_proj = nn.Linear(self.embed, embed_dim)
        self.crop_seq_to_multiple = 1
        self.dropout_input = nn.Dropout(0.0)
        self.dropout_features = nn.Dropout(0.0)
        self.feature_grad_mult = 0.0
        self.quantizer = None
        self.input_quantizer = None
        self.encoder = TransformerEncoder(embed_dim, ffn_dim, heads, layers)
        self.layer_norm = nn.LayerNorm(self.embed)

    def _get_feat_extract_output_lengths(self, input_lengths):
        for _c, k, s in self.conv_layers_cfg:
            input_lengths = torch.floor((input_lengths - k) / s + 1)
        return input_lengths.to(torch.long)


class HubertModel(nn.Module):
    pass


# ------------------------------------------------------------------------------------------------
# rotary_embedding_torch stand-in: interleaved pairs, fp32 angles; xPos as in the
# library's 0.8 line: scale_d = (2d + 0.4 dim) / (1.4 dim), power = (t - len(t) // 2) / 512 where `t` is the
# position slice handed to get_scale (so queries are centred on q_len // 2 and keys on k_len // 2),
# scale repeated per interleaved pair, keys take scale ** -1
# ------------------------------------------------------------------------------------------------
class RotaryEmbedding(nn.Module):
    def __init__(self, dim, use_xpos=False, theta=10000, xpos_scale_base=512):
        super().__init__()
        self.use_xpos = bool(use_xpos)
        self.scale_base = xpos_scale_base
        self.freqs = nn.Parameter(1.0 / (theta ** (torch.arange(0, dim, 2)[: dim // 2].float() / dim)))
        self.register_buffer("scale", (torch.arange(0, dim, 2) + 0.4 * dim) / (1.4 * dim), persistent=False)

    @staticmethod
    def _rotate_half(x):
        x = x.reshape(*x.shape[:-1], -1, 2)
        x1, x2 = x.unbind(dim=-1)
        return torch.stack((-x2, x1), dim=-1).reshape(*x.shape[:-2], -1)

    def get_scale(self, t):
        power = (t - len(t) // 2) / self.scale_base
        scale = self.scale.float() ** power[:, None]
        return scale.repeat_interleave(2, dim=-1)

    def _rotate(self, t, offset, scale=1.0):
        n = t.shape[-2]
        pos = torch.arange(n, device=t.device, dtype=torch.float32) + offset
        freqs = torch.einsum("i,j->ij", pos, self.freqs.float())
        freqs = freqs.repeat_interleave(2, dim=-1)
        out = t.float() * freqs.cos() * scale + self._rotate_half(t.float()) * freqs.sin() * scale
        return out.type_as(t)

    def rotate_queries_with_cached_keys(self, q, k, seq_dim=-2):
        q_len, k_len = q.shape[-2], k.shape[-2]
        assert q_len <= k_len
        q_scale = k_scale = 1.0
        if self.use_xpos:
            seq = torch.arange(k_len, device=q.device, dtype=torch.float32)
            q_scale = self.get_scale(seq[-q_len:]).type(q.dtype)
            k_scale = self.get_scale(seq).type(q.dtype)
            k_scale = k_scale ** -1
        return self._rotate(q, k_len - q_len, q_scale), self._rotate(k, 0, k_scale)


# ------------------------------------------------------------------------------------------------
# simuleval stand-ins
# ------------------------------------------------------------------------------------------------
class AgentStates:
    def __init__(self):
        self.reset()

    def reset(self):
        self.source = []
        self.target = []
        self.source_finished = False
        self.target_finished = False
        self.source_sample_rate = 0


class SpeechToTextAgent:
    def __init__(self, args=None):
        self.args = args
        self.states = None


class ReadAction:
    def is_read(self):
        return True


class WriteAction:
    def __init__(self, content, finished):
        self.content, self.finished = content, finished

    def is_read(self):
        return False


def entrypoint(cls):
    return cls


# ------------------------------------------------------------------------------------------------
# transformers 4.47 pieces
# ------------------------------------------------------------------------------------------------
class DynamicCache447:
    """transformers 4.47 DynamicCache surface used by the reference (`key_cache` / `value_cache` lists,
    `update` = append or cat on dim -2, `cache[i] -> (k, v)`), plus the two methods transformers 5.5's
    LlamaModel asks of a cache object."""

    def __init__(self):
        self.key_cache: List[torch.Tensor] = []
        self.value_cache: List[torch.Tensor] = []

    def update(self, key_states, value_states, layer_idx, cache_kwargs=None):
        if len(self.key_cache) <= layer_idx:
            self.key_cache.append(key_states)
            self.value_cache.append(value_states)
        else:
            self.key_cache[layer_idx] = torch.cat([self.key_cache[layer_idx], key_states], dim=-2)
            self.value_cache[layer_idx] = torch.cat([self.value_cache[layer_idx], value_states], dim=-2)
        return self.key_cache[layer_idx], self.value_cache[layer_idx]

    def __getitem__(self, i):
        return self.key_cache[i], self.value_cache[i]

    def __iter__(self):
        for i in range(len(self.key_cache)):
            yield self.key_cache[i], self.value_cache[i]

    def __len__(self):
        return len(self.key_cache)

    def get_seq_length(self, layer_idx: int = 0) -> int:
        return 0 if len(self.key_cache) <= layer_idx else self.key_cache[layer_idx].shape[-2]

    def get_mask_sizes(self, q_length, layer_idx: int = 0):
        if not isinstance(q_length, int):
            q_length = q_length.shape[0]
        return self.get_seq_length(layer_idx) + q_length, 0


class GreedyOut:
    def __init__(self, sequences, past_key_values, step_logits, step_scores):
        self.sequences, self.past_key_values = sequences, past_key_values
        self.step_logits, self.step_scores = step_logits, step_scores


def greedy_generate(model, input_ids, num_beams=1, max_new_tokens=10, encoder_input_ids=None,
                    encoder_no_repeat_ngram_size=0, no_repeat_ngram_size=0, repetition_penalty=1.0,
                    suppress_tokens=None, past_key_values=None, eos_token_ids=(), **model_kwargs):
    """HF 4.47 `GenerationMixin._sample` with do_sample=False (what patch_hf.py:606-624 dispatches to when
    num_beams == 1), driving the REFERENCE's prepare_inputs_for_generation / forward; logits processors are
    the real transformers classes in the order of `_get_logits_processor`."""
    from transformers.generation.logits_process import (EncoderNoRepeatNGramLogitsProcessor,
                                                        NoRepeatNGramLogitsProcessor,
                                                        RepetitionPenaltyLogitsProcessor,
                                                        SuppressTokensLogitsProcessor)
    assert num_beams == 1, "beam search needs transformers 4.47 internals (patch_hf.py); greedy per SURVEY App. C"
    procs = []
    if repetition_penalty is not None and repetition_penalty != 1.0:
        procs.append(RepetitionPenaltyLogitsProcessor(penalty=repetition_penalty))
    if no_repeat_ngram_size:
        procs.append(NoRepeatNGramLogitsProcessor(no_repeat_ngram_size))
    if encoder_no_repeat_ngram_size and encoder_input_ids is not None and encoder_input_ids.numel() > 0:
        procs.append(EncoderNoRepeatNGramLogitsProcessor(encoder_no_repeat_ngram_size, encoder_input_ids))
    if suppress_tokens:
        procs.append(SuppressTokensLogitsProcessor(suppress_tokens, device=input_ids.device))
    for k in ("attention_mask", "do_sample", "top_p", "top_k", "epsilon_cutoff", "temperature",
              "num_return_sequences", "pad_token_id", "return_dict_in_generate", "return_legacy_cache", "use_cache"):
        model_kwargs.pop(k, None)
    if past_key_values is None:
        past_key_values = DynamicCache447()
    step_logits, step_scores = [], []
    eos = set(int(e) for e in eos_token_ids)
    for _ in range(max_new_tokens):
        inputs = model.prepare_inputs_for_generation(input_ids, past_key_values=past_key_values, **model_kwargs)
        out = model(**inputs, return_dict=True)
        logits = out.logits[:, -1, :].float().clone()
        step_logits.append(logits.clone())
        scores = logits
        for p in procs:
            scores = p(input_ids, scores)
        step_scores.append(scores.clone())
        nxt = torch.argmax(scores, dim=-1)
        input_ids = torch.cat([input_ids, nxt[:, None]], dim=-1)
        if int(nxt[0]) in eos:
            break
    return GreedyOut(input_ids, past_key_values, step_logits, step_scores)


# ------------------------------------------------------------------------------------------------
# beam search: the reference's own loop + scorer (model/patches/patch_hf.py:43-302, 687-967) under stand-ins
# for the transformers 4.47 classes it extends (transformers.generation.beam_search is gone in 5.x)
# ------------------------------------------------------------------------------------------------
class BeamHypotheses447:
    """transformers 4.47 BeamHypotheses minus `add` (the reference installs its own, patch_hf.py:278-302)."""

    def __init__(self, num_beams, length_penalty, early_stopping, max_length=None):
        self.length_penalty, self.early_stopping, self.max_length = length_penalty, early_stopping, max_length
        self.num_beams = num_beams
        self.beams = []
        self.worst_score = 1e9

    def __len__(self):
        return len(self.beams)

    def is_done(self, best_sum_logprobs, cur_len, decoder_prompt_len=0):
        if len(self) < self.num_beams:
            return False
        if self.early_stopping is True:
            return True
        if self.early_stopping is False:
            highest_attainable_score = best_sum_logprobs / (cur_len - decoder_prompt_len) ** self.length_penalty
            return self.worst_score >= highest_attainable_score
        raise NotImplementedError("early_stopping='never' is not used by the reference")


class BeamScorer447:
    pass


class BeamSearchScorer447(BeamScorer447):
    """transformers 4.47 BeamSearchScorer constructor / `is_done`; `process` and `finalize` are the reference's
    (patch_hf.py:43-275), attached by load_ref_patch_hf() the way patch_hf() does (:968-972)."""

    def __init__(self, batch_size, num_beams, device, length_penalty=1.0, do_early_stopping=False,
                 num_beam_hyps_to_keep=1, num_beam_groups=1, max_length=None):
        self.num_beams, self.device, self.length_penalty = num_beams, device, length_penalty
        self.do_early_stopping, self.num_beam_hyps_to_keep = do_early_stopping, num_beam_hyps_to_keep
        self.num_beam_groups = num_beam_groups
        self.group_size = num_beams // num_beam_groups
        self._beam_hyps = [BeamHypotheses447(self.group_size, length_penalty, do_early_stopping, max_length)
                           for _ in range(batch_size * num_beam_groups)]
        self._done = torch.tensor([False] * (batch_size * num_beam_groups), dtype=torch.bool, device=device)

    @property
    def is_done(self):
        return self._done.all()


class SizedDynamicCache447(DynamicCache447):
    """`DynamicCache(num_hidden_layers)` as the reference's scorer constructs it (patch_hf.py:114, 192)."""

    def __init__(self, num_hidden_layers=None):
        super().__init__()


def load_ref_patch_hf():
    """Executes /root/reference/model/patches/patch_hf.py itself (not the no-op placeholder install() registers
    for the agent's import) with the 4.47-only names it imports supplied by the stand-ins above."""
    if getattr(load_ref_patch_hf, "mod", None) is not None:
        return load_ref_patch_hf.mod
    import importlib.util
    import transformers.generation.utils as gu
    bs = types.ModuleType("transformers.generation.beam_search")
    bs.BeamSearchScorer, bs.BeamScorer, bs.BeamHypotheses = BeamSearchScorer447, BeamScorer447, BeamHypotheses447
    sys.modules["transformers.generation.beam_search"] = bs
    for name in ("_split_model_inputs", "stack_model_outputs"):      # only used with low_memory=True
        if not hasattr(gu, name):
            setattr(gu, name, None)
    spec = importlib.util.spec_from_file_location("ref_patch_hf", REFERENCE + "/model/patches/patch_hf.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.DynamicCache = SizedDynamicCache447
    BeamSearchScorer447.process = mod.beam_search_process           # patch_hf.py:968-972
    BeamSearchScorer447.finalize = mod.beam_search_finalize
    BeamHypotheses447.add = mod.beam_hypotheses_add
    load_ref_patch_hf.mod = mod
    return mod


class _BeamSelf:
    """The GenerationMixin surface `generation_mixin_beam_search` (patch_hf.py:687-967) touches: HF 4.47 helper
    methods restated minimally; forward / prepare_inputs_for_generation are the reference model's own."""

    def __init__(self, model):
        self.model, self.config = model, model.config

    def __call__(self, **kw):
        return self.model(**kw)

    def prepare_inputs_for_generation(self, input_ids, **kw):
        return self.model.prepare_inputs_for_generation(input_ids, **kw)

    def _get_initial_cache_position(self, input_ids, model_kwargs):
        return model_kwargs

    def _has_unfinished_sequences(self, this_peer_finished, synced_gpus, device=None, **kw):
        return not this_peer_finished

    def _update_model_kwargs_for_generation(self, outputs, model_kwargs, is_encoder_decoder=False, **kw):
        if getattr(outputs, "past_key_values", None) is not None:
            model_kwargs["past_key_values"] = outputs.past_key_values
        return model_kwargs

    def _temporary_reorder_cache(self, past_key_values, beam_idx):
        # 4.47: DynamicCache.reorder_cache = index_select of every layer on the batch dim
        for i in range(len(past_key_values.key_cache)):
            past_key_values.key_cache[i] = past_key_values.key_cache[i].index_select(0, beam_idx)
            past_key_values.value_cache[i] = past_key_values.value_cache[i].index_select(0, beam_idx)
        return past_key_values


def beam_generate(model, input_ids, num_beams=4, max_new_tokens=10, encoder_input_ids=None,
                  encoder_no_repeat_ngram_size=0, no_repeat_ngram_size=0, repetition_penalty=1.0,
                  suppress_tokens=None, past_key_values=None, eos_token_ids=(), pad_token_id=None, record=None,
                  **model_kwargs):
    """The beam branch of the reference's `generate` (patch_hf.py:626-655): HF's set-up steps restated (real
    transformers processors / stopping criteria, 4.47-style scorer), then the REFERENCE's
    `_expand_inputs_for_generation` (:305-342), `_beam_search` (:687-967), `process` / `finalize` / `add`."""
    from transformers.generation.logits_process import (EncoderNoRepeatNGramLogitsProcessor, LogitsProcessorList,
                                                        NoRepeatNGramLogitsProcessor,
                                                        RepetitionPenaltyLogitsProcessor,
                                                        SuppressTokensLogitsProcessor)
    from transformers.generation.stopping_criteria import EosTokenCriteria, MaxLengthCriteria, StoppingCriteriaList
    mod = load_ref_patch_hf()
    procs = LogitsProcessorList()
    if repetition_penalty is not None and repetition_penalty != 1.0:
        procs.append(RepetitionPenaltyLogitsProcessor(penalty=repetition_penalty))
    if no_repeat_ngram_size:
        procs.append(NoRepeatNGramLogitsProcessor(no_repeat_ngram_size))
    if encoder_no_repeat_ngram_size and encoder_input_ids is not None and encoder_input_ids.numel() > 0:
        procs.append(EncoderNoRepeatNGramLogitsProcessor(encoder_no_repeat_ngram_size, encoder_input_ids))
    if suppress_tokens:
        procs.append(SuppressTokensLogitsProcessor(suppress_tokens, device=input_ids.device))
    for k in ("attention_mask", "do_sample", "top_p", "top_k", "epsilon_cutoff", "temperature",
              "num_return_sequences", "return_dict_in_generate", "return_legacy_cache", "use_cache"):
        model_kwargs.pop(k, None)
    if past_key_values is None:
        past_key_values = DynamicCache447()              # HF generate step 7 creates the cache when none is passed
    max_length = input_ids.shape[1] + max_new_tokens
    eos = torch.tensor([int(e) for e in eos_token_ids], dtype=torch.long)
    stop = StoppingCriteriaList([MaxLengthCriteria(max_length=max_length), EosTokenCriteria(eos_token_id=eos)])
    gen_cfg = types.SimpleNamespace(_pad_token_tensor=torch.tensor(pad_token_id), _eos_token_tensor=eos,
                                    output_attentions=False, output_hidden_states=False, output_scores=False,
                                    output_logits=False, return_dict_in_generate=True, low_memory=False,
                                    do_sample=False)
    scorer = BeamSearchScorer447(batch_size=input_ids.shape[0], num_beams=num_beams, device=input_ids.device,
                                 length_penalty=1.0, do_early_stopping=False, num_beam_hyps_to_keep=1,
                                 max_length=max_length)
    if record is not None:                                # per-step candidates as the reference's scorer receives them
        inner = scorer.process

        def process(input_ids_, next_scores, next_tokens, next_indices, **kw):
            out = inner(input_ids_, next_scores, next_tokens, next_indices, **kw)
            record.append({"scores": next_scores[0].clone(), "tokens": next_tokens[0].clone(),
                           "beams": next_indices[0].clone(), "next_scores": out["next_beam_scores"].clone(),
                           "next_tokens": out["next_beam_tokens"].clone(),
                           "next_beams": out["next_beam_indices"].clone(),
                           "n_hyps": len(scorer._beam_hyps[0]), "done": bool(scorer._done[0])})
            return out
        scorer.process = process
    ids_k, kw_k = mod.generation_mixin_expand_inputs_for_generation(
        expand_size=num_beams, is_encoder_decoder=False, input_ids=input_ids, past_key_values=past_key_values,
        **model_kwargs)
    return mod.generation_mixin_beam_search(_BeamSelf(model), ids_k, scorer, logits_processor=procs,
                                            stopping_criteria=stop, generation_config=gen_cfg, synced_gpus=False,
                                            **kw_k)


# ------------------------------------------------------------------------------------------------
def install() -> None:
    """Put the stand-ins into sys.modules, auto-stub the rest, and make /root/reference importable."""
    if getattr(install, "done", False):
        return
    import transformers
    import transformers.models.llama.modeling_llama as ml
    for name in ("LlamaSdpaAttention", "LlamaFlashAttention2"):          # transformers 4.47 names, gone in 5.x
        if not hasattr(ml, name):
            setattr(ml, name, type(name, (ml.LlamaAttention,), {}))
    import transformers.modeling_flash_attention_utils as mf
    if not hasattr(mf, "_flash_attention_forward"):
        mf._flash_attention_forward = None

    _mod("fairseq")
    fs_utils = _mod("fairseq.utils", index_put=fs_index_put, is_xla_tensor=lambda t: False, softmax=fs_softmax,
                    eval_str_dict=lambda x, type=dict: None if x is None else eval(x),
                    get_activation_fn=lambda name: fs_gelu)
    sys.modules["fairseq"].utils = fs_utils
    _mod("fairseq.models")
    _mod("fairseq.models.wav2vec", TransformerEncoder=TransformerEncoder,
         TransformerSentenceEncoderLayer=TransformerSentenceEncoderLayer, Wav2Vec2Model=Wav2Vec2Model)
    _mod("fairseq.models.wav2vec.wav2vec2", Wav2Vec2Model=Wav2Vec2Model, TransformerEncoder=TransformerEncoder,
         TransformerSentenceEncoderLayer=TransformerSentenceEncoderLayer,
         ConvFeatureExtractionModel=FsConvFeatureExtractionModel)
    _mod("fairseq.models.wav2vec.utils", pad_to_multiple=fs_pad_to_multiple)
    _mod("fairseq.models.hubert")
    _mod("fairseq.models.hubert.hubert", HubertModel=HubertModel)
    _mod("fairseq.modules", GradMultiply=GradMultiply, TransposeLast=TransposeLast, Fp32LayerNorm=Fp32LayerNorm,
         MultiheadAttention=MultiheadAttention)
    _mod("fairseq.modules.multihead_attention", MultiheadAttention=MultiheadAttention)
    _mod("fairseq.modules.fairseq_dropout", FairseqDropout=FairseqDropout)
    _mod("fairseq.modules.quant_noise", quant_noise=quant_noise)
    _mod("rotary_embedding_torch", RotaryEmbedding=RotaryEmbedding)
    _mod("simuleval")
    _mod("simuleval.utils", entrypoint=entrypoint)
    _mod("simuleval.agents", SpeechToTextAgent=SpeechToTextAgent)
    _mod("simuleval.agents.states", AgentStates=AgentStates)
    _mod("simuleval.agents.actions", ReadAction=ReadAction, WriteAction=WriteAction)
    _mod("lightning", LightningModule=nn.Module)
    sys.meta_path.insert(0, _StubFinder(["fairseq", "simuleval", "lightning", "wandb", "jieba", "soundfile", "deepspeed",
                                         "sacrebleu", "peft", "accelerate", "torchaudio", "librosa", "textgrid"]))
    # patch_hf.py is a copy of transformers 4.47's generate/_beam_search and imports 4.47 internals: only the
    # beam machinery lives there; greedy_generate above stands in for HF's `_sample`.
    if REFERENCE not in sys.path:
        sys.path.insert(0, REFERENCE)
    for pkg in ("agents", "model", "train"):      # the reference's top-level directories (an unrelated `agents`
        m = types.ModuleType(pkg)                 # package in site-packages would otherwise shadow the first)
        m.__path__ = [REFERENCE + "/" + pkg]
        sys.modules[pkg] = m
    ph = types.ModuleType("model.patches.patch_hf")
    ph.patch_hf = lambda: None
    sys.modules["model.patches.patch_hf"] = ph
    install.done = True
