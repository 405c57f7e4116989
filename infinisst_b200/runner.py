"""Lock-step driver of many independent streams on one engine (one GPU): the serving-side
equivalent of calling `InfiniSST.policy` once per stream and chunk (agents/infinisst.py:270-394),
with the per-stream agent state (target ids, eviction checkpoints, system prompt size) kept per
stream.  Used by bench.py and the batched parity tests."""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch

from .agent import S2TAgentStates, TemplateTokenizer, evict_plan
from .config import InfiniSSTConfig
from .engine import Engine
from .model import SpeechLlamaForCausalLM, StreamHandle


class LockstepRunner:
    def __init__(self, engine: Engine, cfg: InfiniSSTConfig, n_streams: int, beam: int = 1):
        self.engine, self.cfg, self.beam = engine, cfg, beam
        self.model = SpeechLlamaForCausalLM(cfg, engine=engine)
        self.tok = TemplateTokenizer(cfg)
        self.states: List[S2TAgentStates] = []
        for _ in range(n_streams):
            st = S2TAgentStates()
            st.speech_cache = StreamHandle(engine, engine.open_stream())
            st.past_key_values = st.speech_cache
            st.system_prompt_size = len(cfg.tpl.system_ids)
            self.states.append(st)
        self.sids = [st.speech_cache.sid for st in self.states]
        self.chunk = 0
        n_speech = cfg.tpl.speech_tokens_per_chunk * cfg.gen.latency_multiplier
        self.first_ids = self.tok.system_ids() + self.tok.turn_ids(n_speech)
        self.next_ids = [cfg.tpl.eot_id] + self.tok.turn_ids(n_speech)
        self.first_slots = self.model._slot_map(self.first_ids)
        self.next_slots = self.model._slot_map(self.next_ids)
        self.evictions = 0
        self.evict_log: List[List[Optional[tuple]]] = []

    @property
    def n(self) -> int:
        return len(self.states)

    def prompt(self) -> List[int]:
        return self.first_ids if self.chunk == 0 else self.next_ids

    def _after_generate(self, toks: Sequence[Sequence[int]]) -> List[List[int]]:
        """Drop-last rule (agents/infinisst.py:363) + sliding-window eviction (:337-361) per stream."""
        g = self.cfg.gen
        outs, log = [], []
        for st, t in zip(self.states, toks):
            out_ids = list(t[:-1])
            st.target_ids.extend(out_ids)
            if len(st.target_ids) > 4 * g.no_repeat_ngram_lookback:
                st.target_ids = st.target_ids[-g.no_repeat_ngram_lookback:]
            cur = st.speech_cache.kv_len
            plan = evict_plan(st, cur, g.max_llm_cache_size, g.always_cache_system_prompt)
            if plan is not None:
                self.engine.kv_evict(st.speech_cache.sid, plan[0], plan[1])
                self.evictions += 1
            log.append(None if plan is None else (plan[0], plan[1], cur))   # (keep_prefix, drop_upto, cur)
            outs.append(out_ids)
        self.evict_log.append(log)
        self.chunk += 1
        return outs

    def step_device(self, pcm: torch.Tensor, forced: Optional[Sequence[Sequence[int]]] = None) -> List[List[int]]:
        """One 960 ms chunk for every stream; `pcm` float32 [n, samples] (device or host).  Engine calls only."""
        g = self.cfg.gen
        first = self.chunk == 0
        ids = self.first_ids if first else self.next_ids
        slots = self.first_slots if first else self.next_slots
        self.engine.encode_chunk(self.sids, pcm, g.latency_multiplier)
        enc = [st.target_ids[-g.no_repeat_ngram_lookback:] for st in self.states]
        pin = self.states[0].system_prompt_size if g.always_cache_system_prompt else 0
        if self.beam > 1:      # the reference's shipped decoding: beam search, KV of the best hypothesis handed back
            toks, _ = self.engine.generate_beam(self.sids, [ids] * self.n, [slots] * self.n, enc, g, self.beam, pin_prefix=pin)
        else:
            toks = self.engine.generate(self.sids, [ids] * self.n, [slots] * self.n, enc, g, pin_prefix=pin, forced=forced)
        self.last_tokens = toks
        return self._after_generate(toks)

    def step_api(self, pcm_host: torch.Tensor) -> List[List[int]]:
        """The same step through the reference-facing call: `model.generate` with the kwargs of
        agents/infinisst.py:307-332, host tensors in, host token ids out."""
        g = self.cfg.gen
        ids = torch.tensor([self.prompt()] * self.n, dtype=torch.long)
        enc_t = [st.target_ids[-g.no_repeat_ngram_lookback:] for st in self.states]
        pin = self.states[0].system_prompt_size if g.always_cache_system_prompt else 0
        out = self.model.generate(
            attention_mask=None, input_ids=ids, speech_batch=pcm_host, do_sample=False, num_beams=self.beam,
            max_new_tokens=g.max_new_tokens, num_return_sequences=1, encoder_input_ids=enc_t,
            encoder_no_repeat_ngram_size=g.no_repeat_ngram_size, no_repeat_ngram_size=g.no_repeat_ngram_size,
            repetition_penalty=g.repetition_penalty, pad_token_id=g.pad_token_id, return_dict_in_generate=True,
            use_cache=True, past_key_values=None, suppress_tokens=g.suppress_tokens,
            states=self.states if self.n > 1 else self.states[0], multiplier=g.latency_multiplier, pin_prefix=pin)
        self.last_tokens = out.generated
        return self._after_generate(out.generated)

    def close(self):
        for st in self.states:
            st.speech_cache.close()
