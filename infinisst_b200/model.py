"""Reference-signature shims over the C-ABI engine.

`SpeechLlamaForCausalLM.generate / forward` and `SpeechEncoderW2V2RoPE.encode_speech` keep the
argument names and meaning of the reference (model/llm.py:192-295, model/speech_encoder.py:202-236,
call site agents/infinisst.py:307-332) so agents/infinisst.py can call them unchanged; the tensors
behind `states.speech_cache` / `states.past_key_values` become opaque stream handles
(only their truthiness and `past_key_values[0][0].size(2)` are observed by the agent).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence

import torch

from .config import InfiniSSTConfig
from .engine import Engine


class _KVSizeView:
    """Stands in for a [B, Hkv, L, hd] key tensor: `.size(2)` is the live KV length."""

    def __init__(self, handle: "StreamHandle"):
        self._h = handle

    def size(self, dim: Optional[int] = None):
        cfg = self._h.engine.cfg.llm
        shape = (1, cfg.kv_heads, self._h.kv_len, cfg.head_dim)
        return shape if dim is None else shape[dim]

    @property
    def shape(self):
        return self.size()


class StreamHandle:
    """Opaque per-stream state: W2V2RoPECache (model/speech_encoder.py:85-97) and the
    DynamicCache of the LLM live inside the library, addressed by a stream id."""

    def __init__(self, engine: Engine, sid: int):
        self.engine, self.sid = engine, sid
        self.closed = False

    @property
    def kv_len(self) -> int:
        return self.engine.kv_len(self.sid)

    @property
    def n_steps(self) -> int:           # W2V2RoPECache.n_steps
        return self.engine.enc_steps(self.sid)

    def get_seq_length(self) -> int:    # DynamicCache API
        return self.kv_len

    def __getitem__(self, layer: int):
        v = _KVSizeView(self)
        return (v, v)

    def __len__(self):
        return self.engine.cfg.llm.layers

    def close(self):
        if not self.closed:
            self.engine.close_stream(self.sid)
            self.closed = True


@dataclass
class GenerateOutput:
    sequences: torch.Tensor            # int64 [B, prompt + generated], right-padded with pad_token_id
    past_key_values: object            # StreamHandle (B == 1) or list of StreamHandle
    generated: Optional[List[List[int]]] = None   # extension: chosen tokens per row without padding
    sequences_scores: Optional[torch.Tensor] = None     # beam search: score of the returned hypothesis per row
    beam_trace: Optional[list] = None                   # beam search, beam_trace=True: candidates / choices per step


@dataclass
class CausalLMOutput:
    logits: torch.Tensor               # [B, 1, vocab]: last position only (the reference computes and
    past_key_values: object            # discards the other T-1 rows)
    loss: Optional[torch.Tensor] = None


class SpeechEncoderW2V2RoPE:
    """encode_speech / _get_feat_extract_output_lengths / set_blocksize of
    model/speech_encoder.py:143-145,202-236."""

    def __init__(self, engine: Engine):
        self.engine = engine
        self.blocksize = engine.cfg.enc.block_size
        self.max_cache_size = engine.cfg.enc.max_cache_size
        self._multiplier = 1

    def set_blocksize(self, multiplier: int) -> None:
        self._multiplier = multiplier

    def _get_feat_extract_output_lengths(self, input_lengths: torch.LongTensor) -> torch.LongTensor:
        n = input_lengths.clone().to(torch.long)
        e = self.engine.cfg.enc
        for (_c, k, s) in list(e.conv_layers) + list(e.adapter_layers):
            n = torch.div(n - k, s, rounding_mode="floor") + 1
        return n

    def encode_speech(self, src_tokens: torch.Tensor, src_lens=None, cache: Optional[Sequence[StreamHandle]] = None):
        """src_tokens [B, n_samples] -> (feature [B, T', D_llm] bf16 on the GPU, cache).
        `cache` is a StreamHandle (B == 1) or a list of B handles; None opens new streams."""
        B = src_tokens.shape[0]
        single = isinstance(cache, StreamHandle)
        handles = [cache] if single else (list(cache) if cache is not None else
                                          [StreamHandle(self.engine, self.engine.open_stream()) for _ in range(B)])
        assert len(handles) == B, "one stream handle per row"
        feats = self.engine.encode_chunk([h.sid for h in handles], src_tokens.float(), self._multiplier,
                                         return_feats=True)
        return feats, (handles[0] if (single or (cache is None and B == 1)) else handles)


class _Inner:
    def __init__(self, engine: Engine):
        self.speech_encoder = SpeechEncoderW2V2RoPE(engine)
        self.speech_features_extracted = False   # kept for API parity; the shim is re-entrant (quirk Q4)
        self.inference = True

        class _Emb:
            embedding_dim = engine.cfg.llm.hidden
        self.embed_tokens = _Emb()


class SpeechLlamaForCausalLM:
    """Drop-in for the object the agent calls `generate` on (agents/infinisst.py:150-181, 307-332)."""

    def __init__(self, cfg: InfiniSSTConfig, engine: Optional[Engine] = None, **engine_kwargs):
        self.cfg = cfg
        self.engine = engine or Engine(cfg, **engine_kwargs)
        self.model = _Inner(self.engine)
        self.dtype = torch.bfloat16
        self.device = torch.device(f"cuda:{self.engine.device}")

    def load_state_dict(self, state_dict) -> None:
        self.engine.load_state_dict(state_dict)

    def eval(self):
        return self

    # -------------------------------------------------------------- helpers
    def _slot_map(self, ids: List[int]) -> List[int]:
        """Prompt positions overwritten by speech features (model/llm.py:86-113): for every
        (`user`, `assistant`) header pair, positions [u+3, a-2) take the next a-u-5 speech vectors."""
        l = self.cfg.llm
        n = len(ids)
        up = [t for t in range(1, n) if ids[t] == l.user_token_id and ids[t - 1] == l.start_header_id]
        ap = [t for t in range(1, n) if ids[t] == l.assist_token_id and ids[t - 1] == l.start_header_id]
        slot, idx = [-1] * n, 0
        for u, a in zip(up, ap):
            for j in range(a - u - 5):
                slot[u + 3 + j] = idx + j
            idx += a - u - 5
        return slot

    def _handles(self, states, past_key_values, B: int) -> List[StreamHandle]:
        sts = states if isinstance(states, (list, tuple)) else [states] * B
        if B > 1 and not isinstance(states, (list, tuple)):
            raise ValueError("B > 1 needs one states object per row (independent streams); the reference's "
                             "pseudo-batch tiles one stream (agents/infinisst.py:291-301)")
        out = []
        for st in sts:
            h = getattr(st, "speech_cache", None) if st is not None else None
            if h is None:
                h = StreamHandle(self.engine, self.engine.open_stream())
                if st is not None:
                    st.speech_cache = h
            out.append(h)
        return out

    # -------------------------------------------------------------- reference API
    @torch.inference_mode()
    def generate(self, attention_mask=None, input_ids=None, speech_batch=None, do_sample=False, top_p=1.0, top_k=0,
                 epsilon_cutoff=0.0, temperature=1.0, num_beams=1, max_new_tokens=10, num_return_sequences=1,
                 encoder_input_ids=None, encoder_no_repeat_ngram_size=0, no_repeat_ngram_size=0,
                 repetition_penalty=1.0, pad_token_id=None, return_dict_in_generate=True, return_legacy_cache=False,
                 use_cache=True, past_key_values=None, suppress_tokens=None, states=None, multiplier=1,
                 forced_tokens=None, pin_prefix=0, length_penalty=1.0, beam_follow=None, beam_trace=False,
                 **_unused) -> GenerateOutput:
        if do_sample:
            raise NotImplementedError("infinisst_b200 implements greedy and beam search (do_sample=False), the "
                                      "modes the shipped scripts use")
        if num_beams > 1 and forced_tokens is not None:
            raise ValueError("forced_tokens is the greedy teacher forcing; beam search takes beam_follow")
        if encoder_no_repeat_ngram_size not in (0, no_repeat_ngram_size):
            raise NotImplementedError("encoder_no_repeat_ngram_size must equal no_repeat_ngram_size "
                                      "(agents/infinisst.py:319-320 passes the same value)")
        B = input_ids.shape[0]
        handles = self._handles(states, past_key_values, B)
        if attention_mask is not None:
            # ragged prompts, right-padded as the agent's tokenizer pads (padding_side="right", agents/infinisst.py:137):
            # a stream at its first chunk (system + turn) next to streams in later turns
            lens = attention_mask.to(torch.long).sum(dim=1).tolist()
            ids = [input_ids[b, :lens[b]].tolist() for b in range(B)]
        else:
            ids = [input_ids[b].tolist() for b in range(B)]
        self.model.speech_encoder.set_blocksize(multiplier)
        self.engine.encode_chunk([h.sid for h in handles], speech_batch.float(), multiplier)
        # model/llm.py:101-110 splices `speech_features[i, index : index + a_p - u_p - 5]` between the headers with
        # torch.cat: when the encoder produced FEWER features than the prompt has <sp_patch> slots (short final chunk at
        # multiplier > 1) the spliced sequence is simply shorter.  Same here: the surplus slots are not fed to the LLM.
        slot_maps = [self._slot_map(r) for r in ids]
        n_feat = self.engine.speech_tokens
        fed = [[t for t, sl in zip(r, sm) if sl < n_feat] for r, sm in zip(ids, slot_maps)]
        fed_slots = [[sl for sl in sm if sl < n_feat] for sm in slot_maps]
        enc = [[] for _ in range(B)]
        if isinstance(encoder_input_ids, (list, tuple)):      # extension: ragged per-stream histories
            enc = [[int(t) for t in row] for row in encoder_input_ids]
        elif encoder_input_ids is not None and encoder_input_ids.numel() > 0:
            enc = [[int(t) for t in encoder_input_ids[b].tolist()] for b in range(B)]

        class _G:
            pass
        g = _G()
        g.max_new_tokens = max_new_tokens
        g.no_repeat_ngram_size = no_repeat_ngram_size
        g.repetition_penalty = float(repetition_penalty)
        g.eos_token_ids = list(self.cfg.gen.eos_token_ids)
        g.suppress_tokens = list(suppress_tokens or [])
        extra = {}
        if num_beams > 1:
            # beam search with KV hand-back (patch_hf.py:626-655, 687-967; agents/infinisst.py:334-336): `sequences`
            # is the best hypothesis (closing EOS appended when it fits), the handle continues from ITS cache
            res = self.engine.generate_beam([h.sid for h in handles], fed, fed_slots, enc, g,
                                            num_beams, pin_prefix=pin_prefix, length_penalty=float(length_penalty),
                                            follow=beam_follow, want_trace=beam_trace)
            toks = res[0]
            extra = {"sequences_scores": torch.tensor(res[1]), "beam_trace": res[2] if beam_trace else None}
        else:
            toks = self.engine.generate([h.sid for h in handles], fed, fed_slots, enc, g,
                                        pin_prefix=pin_prefix, forced=forced_tokens)
        pad = self.cfg.gen.pad_token_id if pad_token_id is None else pad_token_id
        width = max(len(r) + len(t) for r, t in zip(ids, toks))
        seqs = torch.full((B, width), pad, dtype=torch.long)
        for b in range(B):
            row = ids[b] + toks[b]
            seqs[b, :len(row)] = torch.tensor(row, dtype=torch.long)
        pkv = handles[0] if B == 1 else handles
        return GenerateOutput(sequences=seqs, past_key_values=pkv, generated=toks, **extra)

    @torch.inference_mode()
    def forward(self, input_ids=None, text_input_ids=None, attention_mask=None, text_attention_mask=None,
                past_key_values=None, inputs_embeds=None, labels=None, text_labels=None, use_cache=None,
                output_attentions=None, output_hidden_states=None, speech_batch=None, src_lengths=None,
                after_lens=None, return_dict=None, states=None, multiplier=1, pin_prefix=0, last_only=False,
                **_unused) -> CausalLMOutput:
        """model/llm.py:192-270.  `logits` is [B, T, vocab] like the reference's (lm_head over every position, :236-237);
        `last_only=True` (extension) computes the last position only, [B, 1, vocab] - all that generation reads."""
        if labels is not None:
            raise NotImplementedError("training loss (model/llm.py:239-258) is out of scope")
        B = input_ids.shape[0] if input_ids is not None else inputs_embeds.shape[0]
        handles = self._handles(states, past_key_values, B)
        sids = [h.sid for h in handles]
        if speech_batch is not None:
            self.model.speech_encoder.set_blocksize(multiplier)
            self.engine.encode_chunk(sids, speech_batch.float(), multiplier)
        if inputs_embeds is not None:
            T = inputs_embeds.shape[1]
            logits = self.engine.forward(sids, None, embeds=inputs_embeds.reshape(B * T, -1), lens=[T] * B,
                                         pin_prefix=pin_prefix, all_positions=not last_only)
        else:
            T = input_ids.shape[1]
            ids = [input_ids[b].tolist() for b in range(B)]
            slots = [self._slot_map(r) if speech_batch is not None else [-1] * len(r) for r in ids]
            logits = self.engine.forward(sids, ids, slots, pin_prefix=pin_prefix, all_positions=not last_only)
        logits = logits[:, None, :] if last_only else logits.view(B, T, -1)
        return CausalLMOutput(logits=logits, past_key_values=handles[0] if B == 1 else handles)

    __call__ = forward
