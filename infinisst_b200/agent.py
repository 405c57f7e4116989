"""SimulEval agent mirror of agents/infinisst.py (class InfiniSST, S2TAgentStates) on the B200 engine.

Same class / method / flag names (agents/infinisst.py:69-72,115-128,185-198,270-394;
agents/options.py:1-125).  Differences, each deliberate:
  * `--beam 1` (greedy) is accepted besides the reference's asserted beam > 1; `--beam k` runs the
    reference's beam search with KV hand-back (patch_hf.py:43-302, 687-967) on the CUDA path;
  * `cache_checkpoints` and `system_prompt_size` live on the per-stream states (quirk Q3) so several
    streams can share one agent/engine;
  * `policy_batch` advances many independent streams per call (the reference can only tile one
    stream, agents/infinisst.py:291-301); streams may join a running batch at any call (their first chunk, with the
    long system + turn prompt, shares the batch with later turns of the others: right-padded ids + attention_mask).
SimulEval is imported when present; otherwise light stand-ins keep the agent usable and testable.
"""
from __future__ import annotations

import contextlib
from dataclasses import dataclass, field
from time import perf_counter
from typing import List, Optional, Sequence

import torch

from .config import InfiniSSTConfig, production_config, tiny_config
from .model import SpeechLlamaForCausalLM

try:  # pragma: no cover - simuleval is not installed in the build image
    from simuleval.agents import SpeechToTextAgent
    from simuleval.agents.actions import ReadAction, WriteAction
    from simuleval.agents.states import AgentStates
    from simuleval.utils import entrypoint
except Exception:  # minimal stand-ins with the attributes the agent touches
    class ReadAction:
        def is_read(self):
            return True

    @dataclass
    class WriteAction:
        content: str = ""
        finished: bool = False

        def is_read(self):
            return False

    class AgentStates:
        def __init__(self):
            self.reset()

        def reset(self):
            self.source: list = []
            self.target: list = []
            self.source_finished = False
            self.target_finished = False
            self.source_sample_rate = 0

    class SpeechToTextAgent:
        def __init__(self, args):
            self.args = args
            self.states = self.build_states()

    def entrypoint(cls):
        return cls


DEFAULT_SPEECH_PATCH_TOKEN = "<sp_patch>"      # train/dataset.py:52
DEFAULT_LATENCY_TOKEN = "<latency_{}>"         # train/dataset.py:56


def synchronized_timer(description: str, sink: Optional[list] = None):
    """agents/infinisst.py:37-48: device-synchronised wall time of one policy step."""
    @contextlib.contextmanager
    def timer():
        torch.cuda.synchronize()
        t0 = perf_counter()
        yield
        torch.cuda.synchronize()
        dt = perf_counter() - t0
        if sink is not None:
            sink.append(dt)
    return timer()


class S2TAgentStates(AgentStates):
    """agents/infinisst.py:50-67 plus the per-stream eviction bookkeeping (quirk Q3)."""
    MAX_SRC_LEN = 1600000

    def __init__(self, src_len=0, speech_cache=None, past_key_values=None, target_ids=None, segment_idx=0,
                 translations_list=None):
        super().__init__()
        self.src_len = src_len
        self.speech_cache = speech_cache
        self.past_key_values = past_key_values
        self.target_ids = target_ids if target_ids is not None else []
        self.segment_idx = segment_idx
        self.translations_list = translations_list if translations_list is not None else []
        self.cache_checkpoints: List[int] = []
        self.system_prompt_size = 0

    def reset(self):
        super().reset()
        for m in getattr(self, "_pseudo_mirrors", None) or []:
            m.reset()
        self._pseudo_mirrors = None
        if getattr(self, "speech_cache", None) is not None:
            self.speech_cache.close()        # give the stream slot and its KV pages back
        self.src_len = 0
        self.speech_cache = None
        self.past_key_values = None
        self.target_ids = []
        self.segment_idx = 0
        self.translations_list = []
        self.cache_checkpoints = []
        self.system_prompt_size = 0


class TemplateTokenizer:
    """Stand-in for the Llama-3.1 tokenizer + chat template when no tokenizer files are available:
    reproduces the turn layout of agents/infinisst.py:225-268 from a TemplateConfig."""

    def __init__(self, cfg: InfiniSSTConfig):
        self.tpl = cfg.tpl
        self.pad_token_id = cfg.gen.pad_token_id

    def system_ids(self) -> List[int]:
        return list(self.tpl.system_ids)

    def turn_ids(self, n_speech: int) -> List[int]:
        t = self.tpl
        return [t.start_header_id, t.user_token_id, t.end_header_id, t.nl_id] + [t.sp_patch_id] * n_speech + \
               [t.eot_id, t.start_header_id, t.assist_token_id, t.end_header_id, t.nl_id]

    def decode(self, ids, skip_special_tokens: bool = True) -> str:
        if isinstance(ids, int):
            ids = [ids]
        return " ".join(f"t{i}" for i in ids)


def non_language_token_ids(tokenizer, n_vocab: int) -> List[int]:
    """`--suppress-non-language` (agents/infinisst.py:142-148): ids of every token whose decoded text contains an
    opening parenthesis (ASCII or full-width); they are passed to generate as `suppress_tokens` (:329)."""
    bad_words = ["(", "\uff08"]
    out = []
    for idx in range(n_vocab):
        decoded = tokenizer.decode(idx, skip_special_tokens=True)
        if any(b in decoded for b in bad_words):
            out.append(idx)
    return out


def evict_plan(states: S2TAgentStates, cur: int, max_llm_cache_size: int, always_cache_system_prompt: bool):
    """Integer part of agents/infinisst.py:337-352.  Returns None or (keep_prefix, drop_upto): the
    logical KV tokens [keep_prefix, drop_upto) are evicted."""
    states.cache_checkpoints.append(cur)
    if cur <= max_llm_cache_size:
        return None
    new = 0
    for i, ckpt in enumerate(states.cache_checkpoints):
        new = cur - ckpt
        if new <= max_llm_cache_size:
            rest = states.cache_checkpoints[i + 1:]
            trimmed = ckpt - (states.system_prompt_size if always_cache_system_prompt else 0)
            states.cache_checkpoints = [c - trimmed for c in rest]
            break
    if new == 0:
        # a single turn longer than the window: the reference's `k[:, :, -0:]` keeps everything (:355)
        # (and would duplicate the system prompt); nothing is evicted here.
        return None
    return (states.system_prompt_size if always_cache_system_prompt else 0, cur - new)


@entrypoint
class InfiniSST(SpeechToTextAgent):
    def __init__(self, args):
        self.min_start_sec = args.min_start_sec
        self.latency_multiplier = args.latency_multiplier
        self.source_segment_size = getattr(args, "source_segment_size", 960 * args.latency_multiplier)
        self.max_latency_multiplier = args.max_latency_multiplier
        self.source_lang = args.source_lang
        self.target_lang = args.target_lang
        self.beam = args.beam
        if self.beam < 1:
            raise ValueError("--beam must be >= 1")
        self.no_repeat_ngram_lookback = args.no_repeat_ngram_lookback
        self.no_repeat_ngram_size = args.no_repeat_ngram_size
        self.repetition_penalty = args.repetition_penalty
        self.suppress_non_language = args.suppress_non_language
        self.max_new_tokens = args.max_new_tokens
        # sampling flags are handed to generate like the reference does (agents/infinisst.py:93-97, 311-315); the CUDA
        # path implements the shipped modes (greedy, beam search) and refuses do_sample loudly instead of ignoring it
        self.do_sample = getattr(args, "do_sample", False)
        self.top_p, self.top_k = getattr(args, "top_p", 1.0), getattr(args, "top_k", 0)
        self.epsilon_cutoff, self.temperature = getattr(args, "epsilon_cutoff", 0.0), getattr(args, "temperature", 1.0)
        # --pseudo-batch-size B (agents/infinisst.py:291-301): the reference tiles ONE stream B times (audio, prompt, both
        # caches) to measure batched throughput and reads row 0.  Here B engine streams mirror the stream: they are fed
        # the same audio / prompt through one batched call, so the batched kernels run on B rows and row 0 is returned.
        self.pseudo_batch_size = getattr(args, "pseudo_batch_size", 1)
        if self.pseudo_batch_size < 1:
            raise ValueError("--pseudo-batch-size must be >= 1")
        self.max_llm_cache_size = args.max_llm_cache_size
        self.always_cache_system_prompt = args.always_cache_system_prompt
        self.dpo_sampling = getattr(args, "dpo_sampling", False)              # agents/infinisst.py:100-101
        self.output_file = getattr(args, "output_file", "translations.json")
        self.chunk_latencies: List[float] = []
        self.args = args
        self.load_model(args)
        super().__init__(args)

    def build_states(self):
        return S2TAgentStates()

    def update_multiplier(self, multiplier):
        self.latency_multiplier = multiplier
        self.max_new_tokens = 10 * multiplier

    def load_model(self, args):
        """agents/infinisst.py:130-183.  Architecture: `args.model_config` (an InfiniSSTConfig, 'tiny' or
        'production') when given, otherwise read off the checkpoint itself (`--state-dict-path` shapes +
        `--model-name` config.json / generation_config.json + `--w2v2-path` arguments + `--length-shrink-cfg`,
        checkpoint.py).  Weights: `args.state_dict` or `--state-dict-path`, strict like `load_state_dict` (:180).
        Tokenizer: `args.tokenizer`, else `AutoTokenizer.from_pretrained(--model-name, padding_side="right",
        use_fast=False)` when that is a local directory (:135-140) followed by the reference's `preprocess`
        (:177), else the template stand-in."""
        from . import checkpoint as ck
        cfg = getattr(args, "model_config", None)
        sd = getattr(args, "state_dict", None)
        path = getattr(args, "state_dict_path", None)
        if sd is None and path:
            sd = ck.load_reference_state_dict(path)
        if cfg is None and sd is not None and path:
            hf, gen = ck.read_hf_config(getattr(args, "model_name", None))
            w2v2 = getattr(args, "w2v2_path", None)
            import os
            wa = ck.read_w2v2_args(w2v2) if w2v2 and os.path.isfile(w2v2) else None
            cfg = ck.infer_config(sd, block_size=args.block_size, max_cache_size=args.max_cache_size,
                                  length_shrink_cfg=getattr(args, "length_shrink_cfg", None), xpos=bool(args.xpos),
                                  rope=bool(args.rope), hf_config=hf, w2v2_args=wa, generation_config=gen)
            cfg.tpl.system_ids = production_config().tpl.system_ids if cfg.llm.vocab >= 128256 else []
        cfg = cfg or "production"
        if isinstance(cfg, str):
            cfg = tiny_config() if cfg == "tiny" else production_config()
        cfg.enc.block_size = args.block_size
        cfg.enc.max_cache_size = args.max_cache_size
        cfg.enc.rope, cfg.enc.xpos = bool(args.rope), bool(args.xpos)
        if args.w2v2_type != "w2v2":
            raise ValueError(f"Unsupported type: {args.w2v2_type}")            # agents/infinisst.py:171
        self.cfg = cfg
        self.tokenizer = getattr(args, "tokenizer", None)
        if self.tokenizer is None:
            self.tokenizer = self._hf_tokenizer(getattr(args, "model_name", None))
        if self.tokenizer is None:
            self.tokenizer = TemplateTokenizer(cfg)
        elif not isinstance(self.tokenizer, TemplateTokenizer) and hasattr(self.tokenizer, "add_tokens"):
            ck.preprocess_tokenizer(self.tokenizer, cfg, self.max_latency_multiplier)       # :177
            if hasattr(self.tokenizer, "apply_chat_template"):
                # the real system turn (agents/infinisst.py:229-241) sizes the pinned prefix / KV capacity
                ck.template_from_tokenizer(self.tokenizer, cfg, self.source_lang, self.target_lang, self.latency_multiplier)
        self.bad_words_ids = list(getattr(args, "bad_words_ids", []) or [])
        if self.suppress_non_language and not self.bad_words_ids:
            n_vocab = len(self.tokenizer) if hasattr(self.tokenizer, "__len__") else cfg.llm.vocab
            self.bad_words_ids = non_language_token_ids(self.tokenizer, n_vocab)
        # cfg.gen mirrors the flags the engine sizes itself from (the agent evicts with --max-llm-cache-size, so the KV
        # capacity must follow it, not the config default)
        cfg.gen.max_llm_cache_size = int(args.max_llm_cache_size)
        cfg.gen.always_cache_system_prompt = bool(args.always_cache_system_prompt)
        cfg.gen.no_repeat_ngram_size = int(args.no_repeat_ngram_size)
        cfg.gen.no_repeat_ngram_lookback = int(args.no_repeat_ngram_lookback)
        cfg.gen.repetition_penalty = float(args.repetition_penalty)
        cfg.gen.latency_multiplier = int(self.latency_multiplier)
        cfg.gen.beam = int(self.beam)
        # engine capacities from the flags: first-chunk prompt = system + 9 header tokens + block_size//4 * m speech
        # slots (agents/infinisst.py:225-260; 73 / 85 / 97 tokens at m = 2 / 3 / 4 with a 40-token system turn), longest
        # generation = --max-new-tokens or 10 * m after update_multiplier (:125-128)
        n_speech_max = args.block_size // 4 * max(self.max_latency_multiplier, self.latency_multiplier)
        max_prompt = len(cfg.tpl.system_ids) + 10 + n_speech_max + 8
        max_new = max(int(self.max_new_tokens), 10 * max(self.max_latency_multiplier, self.latency_multiplier))
        self.model = SpeechLlamaForCausalLM(
            cfg, engine=getattr(args, "engine", None),
            max_streams=max(getattr(args, "max_streams", 8), self.pseudo_batch_size),
            max_multiplier=max(self.max_latency_multiplier, self.latency_multiplier), max_beams=self.beam,
            max_prompt=max(64, max_prompt), max_new_tokens=max_new)
        if sd is not None:
            ck.check_state_dict(sd, cfg)
            self.model.load_state_dict(sd)
        self.model.model.inference = True
        self.llama31 = "3.1" in str(getattr(args, "model_name", ""))         # :183

    @staticmethod
    def _hf_tokenizer(model_name):
        """agents/infinisst.py:135-140 for a local model directory (this build never reaches for the hub)."""
        import os
        if not model_name or not os.path.isdir(model_name):
            return None
        import transformers
        tok = transformers.AutoTokenizer.from_pretrained(model_name, padding_side="right", use_fast=False)
        tok.pad_token = "<|finetune_right_pad_id|>"
        return tok

    @staticmethod
    def add_args(parser):
        # agents/options.py:1-125 + agents/infinisst.py:185-198
        parser.add_argument("--source-lang", type=str, default="English")
        parser.add_argument("--target-lang", type=str, default="German")
        parser.add_argument("--min-start-sec", default=0.32, type=float)
        parser.add_argument("--w2v2-path", type=str, default=None)
        parser.add_argument("--w2v2-type", type=str, default=None)
        parser.add_argument("--ctc-finetuned", type=lambda x: (str(x).lower() == "true"), default=False)
        parser.add_argument("--length-shrink-cfg", type=str, default=None)
        parser.add_argument("--block-size", type=int, default=12)
        parser.add_argument("--max-cache-size", type=int, default=125)
        parser.add_argument("--xpos", type=int, default=1)
        parser.add_argument("--rope", type=int, default=1)
        parser.add_argument("--max-len-a", type=int, default=5)
        parser.add_argument("--max-len-b", type=int, default=20)
        parser.add_argument("--beam", type=int, default=1)
        parser.add_argument("--no-repeat-ngram-lookback", type=int, default=100)
        parser.add_argument("--no-repeat-ngram-size", type=int, default=3)
        parser.add_argument("--repetition-penalty", type=float, default=1.2)
        parser.add_argument("--suppress-non-language", action="store_true")
        parser.add_argument("--max-new-tokens", type=int, default=1000)
        parser.add_argument("--do-sample", action="store_true")
        parser.add_argument("--top-p", type=float, default=1.0)
        parser.add_argument("--top-k", type=int, default=0)
        parser.add_argument("--epsilon-cutoff", type=float, default=0.0)
        parser.add_argument("--temperature", type=float, default=1.0)
        parser.add_argument("--model-name", type=str, default="facebook/opt-350m")
        parser.add_argument("--state-dict-path", type=str, default=None)
        parser.add_argument("--latency-multiplier", type=int, default=4)
        parser.add_argument("--max-latency-multiplier", type=int, default=4)
        parser.add_argument("--max-llm-cache-size", type=int, default=10000)
        parser.add_argument("--always-cache-system-prompt", action="store_true")
        parser.add_argument("--dpo-sampling", action="store_true")
        parser.add_argument("--output-file", type=str, default="translations.json")
        parser.add_argument("--pseudo-batch-size", type=int, default=1)

    # ---------------------------------------------------------------- per-chunk pieces
    def _prepare_speech(self, states, explicit_offset: bool = True) -> torch.Tensor:
        """agents/infinisst.py:200-223; returns float32 [1, n] on the host (the bf16 cast of :222
        happens on the device inside conv0)."""
        sp_seg_frame = int(self.args.block_size // 4 * 0.08 * 16000)
        if len(states.source) > states.MAX_SRC_LEN:
            states.src_len -= len(states.source) - states.MAX_SRC_LEN
            states.source = states.source[-states.MAX_SRC_LEN:]
        src = states.source[states.src_len:]
        source = src.float() if isinstance(src, torch.Tensor) else torch.tensor(src, dtype=torch.float32)
        if source.size(0) % sp_seg_frame != 0:                                # :211-213: ONE segment, whatever m is
            n_pad = sp_seg_frame - source.size(0) % sp_seg_frame
            source = torch.cat([source, torch.zeros(n_pad)], dim=0)
        if source.size(0) > sp_seg_frame * self.latency_multiplier:
            # more than m segments pending (e.g. --min-start-sec beyond one chunk): the reference would encode them all
            # and splice only the first 12 * m features (model/llm.py:101-110); that silent truncation is refused here
            raise ValueError(f"{source.size(0) // sp_seg_frame} audio segments are pending but one policy call takes at "
                             f"most latency_multiplier = {self.latency_multiplier}; call policy once per source segment")
        if states.src_len == 0 and explicit_offset:
            source = torch.cat([torch.zeros(79 + 320), source], dim=0)
        states.src_len = len(states.source)
        return source.unsqueeze(0)

    def _prepare_inputs(self, states) -> torch.Tensor:
        """agents/infinisst.py:225-268.  With a HF tokenizer the chat template is applied exactly as
        the reference does; with the TemplateTokenizer the same layout is built from the template ids."""
        n_speech = self.args.block_size // 4 * self.latency_multiplier
        tok = self.tokenizer
        if isinstance(tok, TemplateTokenizer):
            ids: List[int] = []
            if states.speech_cache is None:
                ids += tok.system_ids()
                states.system_prompt_size = len(ids)
                ids += tok.turn_ids(n_speech)
            else:
                ids = [self.cfg.tpl.eot_id] + tok.turn_ids(n_speech)
            return torch.tensor([ids], dtype=torch.long)
        def chat_ids(msgs):
            # transformers >= 5 returns a BatchEncoding here, 4.47 (the reference's pin) the id tensor itself
            out = tok.apply_chat_template([msgs], return_tensors="pt", padding=True, truncation=False, add_special_tokens=False)
            return out["input_ids"] if hasattr(out, "keys") else out
        messages = []
        if states.speech_cache is None:
            latency_token = DEFAULT_LATENCY_TOKEN.format(self.latency_multiplier)
            messages.append({"role": "system", "content": f"Translate the following speech from {self.source_lang} "
                                                          f"to {self.target_lang} with latency {latency_token}."})
            states.system_prompt_size = chat_ids(messages).size(1)
        messages.append({"role": "user", "content": n_speech * DEFAULT_SPEECH_PATCH_TOKEN})
        messages.append({"role": "assistant", "content": ""})
        input_ids = chat_ids(messages)[:, :-1]
        if states.speech_cache is not None:
            if self.llama31:
                input_ids = input_ids[:, 25:]      # Llama-3.1 default system header (agents/infinisst.py:262-264)
            else:
                input_ids[:, 0] = tok.eos_token_id  # llama-3-8B-instruct (:265-266)
        return input_ids

    def _evict(self, states) -> None:
        cur = states.past_key_values[0][0].size(2)
        plan = evict_plan(states, cur, self.max_llm_cache_size, self.always_cache_system_prompt)
        if plan is not None:
            states.past_key_values.engine.kv_evict(states.past_key_values.sid, plan[0], plan[1])

    def _finish(self, states, input_len: int, sequence: List[int]):
        output_ids = sequence[input_len:-1]                                   # agents/infinisst.py:363
        states.target_ids.extend(output_ids)
        translation = self.tokenizer.decode(output_ids, skip_special_tokens=True).strip().replace("�", "")
        if getattr(self, "dpo_sampling", False):                              # agents/infinisst.py:369-382
            states.translations_list.append(f"'{translation}'" if translation else "''")
            if states.source_finished:
                try:
                    with open(self.output_file, "a", encoding="utf-8") as f:
                        f.write(f"[{', '.join(states.translations_list)}]" + "\n")
                    states.translations_list = []
                except Exception as e:  # noqa: BLE001 - the reference reports and carries on
                    print(f"Error writing translations to file: {e}")
        # the reference's per-chunk log line (agents/infinisst.py:384): KV length and the words emitted so far
        if getattr(getattr(self, "args", None), "log_chunks", True) and states.past_key_values is not None:
            print(states.past_key_values[0][0].size(2), (" " if self.target_lang != "Chinese" else "").join(states.target))
        states.segment_idx += 1
        if translation != "" or states.source_finished:
            return WriteAction(content=translation, finished=states.source_finished)
        return ReadAction()

    @torch.inference_mode()
    def policy(self, states: Optional[S2TAgentStates] = None):
        if states is None:
            states = self.states
        if self.pseudo_batch_size > 1:
            # agents/infinisst.py:291-301: B - 1 mirror streams receive the same audio and prompts, row 0 is the stream
            mirrors = getattr(states, "_pseudo_mirrors", None)
            if mirrors is None:
                mirrors = states._pseudo_mirrors = [self.build_states() for _ in range(self.pseudo_batch_size - 1)]
            for m in mirrors:
                m.source, m.source_sample_rate, m.source_finished = states.source, states.source_sample_rate, states.source_finished
            return self.policy_batch([states] + mirrors)[0]
        return self.policy_batch([states])[0]

    @torch.inference_mode()
    def policy_batch(self, states_list: Sequence[S2TAgentStates]):
        """agents/infinisst.py:270-394 for several independent streams in lock-step."""
        actions: List[object] = [None] * len(states_list)
        run = []
        for i, st in enumerate(states_list):
            secs = float(len(st.source)) / st.source_sample_rate if st.source_sample_rate else 0.0
            if not st.source_finished and secs < self.min_start_sec:
                actions[i] = ReadAction()
            elif st.source_finished and secs < 0.32:
                actions[i] = WriteAction(content="", finished=True)
            else:
                run.append(i)
        if not run:
            return actions
        with synchronized_timer("generate", self.chunk_latencies):
            sts = [states_list[i] for i in run]
            for st in sts:
                if st.source_finished:
                    st.segment_idx = -1                                       # agents/infinisst.py:303-304
            first = [st.speech_cache is None for st in sts]
            mixed = any(first) != all(first)
            # streams may join a running batch: a joining stream brings the long first-chunk prompt (system + turn)
            # and its 79+320 zero offset is the library's zero carried tail, so every row has the same sample count
            speech = torch.cat([self._prepare_speech(st, explicit_offset=not mixed) for st in sts], dim=0)
            rows = [self._prepare_inputs(st)[0] for st in sts]
            lens = [int(r.numel()) for r in rows]
            ids = torch.full((len(rows), max(lens)), int(self.tokenizer.pad_token_id), dtype=torch.long)
            mask = torch.zeros(len(rows), max(lens), dtype=torch.long)
            for j, r in enumerate(rows):
                ids[j, :lens[j]] = r
                mask[j, :lens[j]] = 1
            enc = [st.target_ids[-self.no_repeat_ngram_lookback:] for st in sts]
            sizes = {st.system_prompt_size for st in sts}
            if self.always_cache_system_prompt and len(sizes) != 1:
                raise ValueError("streams of one batch must share the system prompt length")
            pin = sts[0].system_prompt_size if self.always_cache_system_prompt else 0
            outputs = self._generate(sts, ids, speech, enc, pin, mask if min(lens) != max(lens) else None)
            for st in sts:
                st.past_key_values = st.speech_cache
                self._evict(st)
        for j, i in enumerate(run):
            seq = [t for t in outputs.sequences[j].tolist()]
            n_gen = self._n_generated[j]
            seq = seq[: lens[j] + n_gen]
            actions[i] = self._finish(states_list[i], lens[j], seq)
        return actions

    def _generate(self, sts, ids, speech, enc, pin, attention_mask=None):
        """model.generate with the kwargs of agents/infinisst.py:307-332 (ragged `encoder_input_ids`
        are passed per stream instead of one padded tensor)."""
        enc_t = enc if any(len(e) for e in enc) else None
        out = self.model.generate(
            attention_mask=attention_mask, input_ids=ids, speech_batch=speech, do_sample=self.do_sample, top_p=self.top_p,
            top_k=self.top_k, epsilon_cutoff=self.epsilon_cutoff, temperature=self.temperature, num_beams=self.beam, max_new_tokens=self.max_new_tokens,
            num_return_sequences=1, encoder_input_ids=enc_t, encoder_no_repeat_ngram_size=self.no_repeat_ngram_size,
            no_repeat_ngram_size=self.no_repeat_ngram_size, repetition_penalty=self.repetition_penalty,
            pad_token_id=self.tokenizer.pad_token_id, return_dict_in_generate=True, return_legacy_cache=False,
            use_cache=True, past_key_values=None, suppress_tokens=self.bad_words_ids,
            states=sts if len(sts) > 1 else sts[0], multiplier=self.latency_multiplier, pin_prefix=pin)
        self._n_generated = [len(t) for t in out.generated]
        return out

    def _gen_cfg(self):
        class _G:
            pass
        g = _G()
        g.max_new_tokens = self.max_new_tokens
        g.no_repeat_ngram_size = self.no_repeat_ngram_size
        g.repetition_penalty = float(self.repetition_penalty)
        g.eos_token_ids = list(self.cfg.gen.eos_token_ids)
        g.suppress_tokens = list(self.bad_words_ids)
        return g
