"""Host-side engine: owns one `isst_ctx` (one per GPU / process) and exposes the per-chunk
step for batches of independent streams.  PyTorch is used for tensors and CUDA streams only;
all compute goes through the C-ABI (include/infinisst_b200.h)."""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, List, Optional, Sequence

import torch

from . import _lib
from .config import InfiniSSTConfig

ENC = "model.speech_encoder.speech_encoder."


def llama_inv_freq(llm_cfg) -> torch.Tensor:
    """HF llama3 RoPE frequencies (rope_type='llama3'; patch_llm.py:287-299 calls
    LlamaRotaryEmbedding; constants SURVEY App. A.3)."""
    hd = llm_cfg.head_dim
    inv = 1.0 / (llm_cfg.rope_theta ** (torch.arange(0, hd, 2, dtype=torch.int64).float() / hd))
    sc = llm_cfg.rope_scaling
    if not sc:
        return inv
    factor, lo, hi, old = sc["factor"], sc["low_freq_factor"], sc["high_freq_factor"], \
        sc["original_max_position_embeddings"]
    wl = 2 * math.pi / inv
    scaled = torch.where(wl > old / lo, inv / factor, inv)
    smooth = (old / wl - lo) / (hi - lo)
    mid = (1 - smooth) * scaled / factor + smooth * scaled
    medium = ~(wl < old / hi) & ~(wl > old / lo)
    return torch.where(medium, mid, scaled)


def _ints(vals: Sequence[int], ctype=C.c_int):
    return (ctype * max(1, len(vals)))(*vals)


class Engine:
    def __init__(self, cfg: InfiniSSTConfig, device: int = 0, max_streams: int = 8, max_batch: Optional[int] = None,
                 max_multiplier: int = 1, kv_pages: Optional[int] = None, max_kv_len: Optional[int] = None,
                 max_prompt: int = 64, max_new_tokens: Optional[int] = None, max_beams: int = 1):
        if not torch.cuda.is_available():
            raise RuntimeError("infinisst_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.lib = _lib.load()
        self.cfg = cfg
        self.device = device
        e, l, g = cfg.enc, cfg.llm, cfg.gen
        max_batch = max_batch or max_streams * max_beams        # beam search advances streams x beams rows per step
        max_new = max_new_tokens or max(g.max_new_tokens, 10 * max_multiplier)
        if max_kv_len is None:
            # window + pinned system prompt + one full turn of slack (the cache is trimmed after the turn)
            max_kv_len = g.max_llm_cache_size + len(cfg.tpl.system_ids) + 2 * (max_prompt + max_new) + 64
        if kv_pages is None:
            kv_pages = max_streams * (max_kv_len // 16 + 3)
            if max_beams > 1:       # private tail page sets: 2 per beam + the kept hypotheses' snapshots (+1 in flight)
                kv_pages += max_streams * (3 * max_beams + 1) * ((15 + max_new) // 16 + 1)
        c = _lib.IsstConfig()
        c.n_conv = len(e.conv_layers)
        for j, (d, k, s) in enumerate(e.conv_layers):
            c.conv_dim[j], c.conv_k[j], c.conv_s[j] = d, k, s
        c.enc_dim, c.enc_ffn, c.enc_heads, c.enc_layers = e.embed_dim, e.ffn_dim, e.heads, e.layers
        c.block_size, c.max_cache_size = e.block_size, e.max_cache_size
        c.n_adapter = len(e.adapter_layers)
        for j, (d, k, s) in enumerate(e.adapter_layers):
            c.adapter_dim[j], c.adapter_k[j], c.adapter_s[j] = d, k, s
        c.hidden, c.layers, c.heads, c.kv_heads = l.hidden, l.layers, l.heads, l.kv_heads
        c.head_dim, c.ffn, c.vocab, c.rms_eps = l.head_dim, l.ffn, l.vocab, l.rms_eps
        c.max_streams, c.max_batch, c.max_multiplier = max_streams, max_batch, max_multiplier
        c.kv_pages, c.max_kv_len, c.max_prompt, c.max_new_tokens = kv_pages, max_kv_len, max_prompt, max_new
        # with --rope 0 the rotary embedding (and its xPos scaling) is never applied: patch_speech_encoder.py:823-824
        c.enc_xpos, c.enc_no_rope = int(bool(e.xpos and e.rope)), int(not e.rope)
        self._c = c
        self.max_kv_len = max_kv_len
        self.max_multiplier = max_multiplier
        h = C.c_void_p()
        _lib.check(self.lib.isst_create(C.byref(c), device, C.byref(h)))
        self.h = h
        self.total_stride = 1
        for (_d, _k, s) in e.conv_layers:
            self.total_stride *= s
        self.chunk_samples = e.block_size * self.total_stride          # 15360 at block 48
        self.first_offset = 79 + 320 if self.total_stride == 320 else None
        self.speech_tokens = None

    # ------------------------------------------------------------------ lifetime
    def close(self):
        if getattr(self, "h", None):
            self.lib.isst_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ weights
    def _load(self, name: str, t: torch.Tensor):
        if t.dtype == torch.float32:
            dt = _lib.DTYPE_F32
        elif t.dtype == torch.bfloat16:
            dt = _lib.DTYPE_BF16
        else:
            t, dt = t.float(), _lib.DTYPE_F32
        t = t.contiguous()
        shape = (C.c_int64 * max(1, t.dim()))(*t.shape)
        _lib.check(self.lib.isst_load_weight(self.h, name.encode(), C.c_void_p(t.data_ptr()), shape, t.dim(), dt))

    def load_state_dict(self, sd: Dict[str, torch.Tensor]) -> None:
        """Ingest a reference-layout state dict (agents/infinisst.py:179-180) and the RoPE tables."""
        for k, v in sd.items():
            self._load(k, v)
        e, l = self.cfg.enc, self.cfg.llm
        if e.rope_angle_dtype != "fp32":
            raise NotImplementedError("encoder RoPE angles are formed in fp32/fp64 (rope_angle_dtype='fp32'); the "
                                      "bf16-position variant of some rotary_embedding_torch versions is not built")
        # RoPE frequencies: rotary_embedding_torch's `freqs` parameter (one copy per layer, identical; layer 0 is
        # used) and HF's llama3-scaled inv_freq.  The library forms the angles itself, at absolute positions.
        freqs = sd[ENC + "encoder.layers.0.self_attn.rotary_emb.freqs"].detach().float().cpu().contiguous()
        self._load("rope.enc.inv_freq", freqs)
        self._load("rope.llm.inv_freq", llama_inv_freq(l).float().contiguous())
        _lib.check(self.lib.isst_finalize_weights(self.h))

    # ------------------------------------------------------------------ streams
    def open_stream(self) -> int:
        sid = C.c_int()
        _lib.check(self.lib.isst_stream_open(self.h, C.byref(sid)))
        return sid.value

    def close_stream(self, sid: int) -> None:
        _lib.check(self.lib.isst_stream_close(self.h, sid))

    def kv_len(self, sid: int) -> int:
        v = C.c_int()
        _lib.check(self.lib.isst_kv_len(self.h, sid, C.byref(v)))
        return v.value

    def enc_steps(self, sid: int) -> int:
        v = C.c_int()
        _lib.check(self.lib.isst_enc_steps(self.h, sid, C.byref(v)))
        return v.value

    def kv_evict(self, sid: int, keep_prefix: int, drop_upto: int) -> None:
        _lib.check(self.lib.isst_kv_evict(self.h, sid, keep_prefix, drop_upto))

    def debug_shift_positions(self, sid: int, delta: int) -> None:
        _lib.check(self.lib.isst_debug_shift_positions(self.h, sid, delta))

    def pages_free(self) -> int:
        return self.lib.isst_pages_free(self.h)

    def launch_count(self) -> int:
        return self.lib.isst_launch_count(self.h)

    def path_count(self, name: str) -> int:
        """Launches so far of one kernel variant (isst_path_count)."""
        v = C.c_int64()
        _lib.check(self.lib.isst_path_count(self.h, name.encode(), C.byref(v)))
        return v.value

    def option(self, key: str, value: int) -> None:
        _lib.check(self.lib.isst_debug_option(self.h, key.encode(), int(value)))

    # ------------------------------------------------------------------ the per-chunk step
    def _stream_ptr(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def encode_chunk(self, sids: Sequence[int], pcm: torch.Tensor, multiplier: int = 1,
                     return_feats: bool = False) -> Optional[torch.Tensor]:
        """pcm: float32 [n, n_samples], host (ideally pinned) or CUDA."""
        assert pcm.dtype == torch.float32 and pcm.dim() == 2 and pcm.shape[0] == len(sids)
        pcm = pcm.contiguous()
        n, ns = pcm.shape
        out, out_ptr = None, None
        shrink = 1
        for (_d, _k, st) in self.cfg.enc.adapter_layers:
            shrink *= st
        n_tok = ((ns - ns % self.chunk_samples) // self.total_stride) // shrink
        if return_feats:
            out = torch.empty(n, n_tok, self.cfg.llm.hidden, dtype=torch.bfloat16, device=f"cuda:{self.device}")
            out_ptr = C.c_void_p(out.data_ptr())
        _lib.check(self.lib.isst_encode_chunk(self.h, n, _ints(sids), C.c_void_p(pcm.data_ptr()), ns, multiplier,
                                              out_ptr, self._stream_ptr()))
        self.speech_tokens = n_tok
        return out

    def _gen_params(self, gen, max_new: Optional[int] = None, pin_prefix: int = 0):
        gp = _lib.IsstGenParams()
        gp.max_new_tokens = max_new or gen.max_new_tokens
        gp.no_repeat_ngram_size = gen.no_repeat_ngram_size
        gp.repetition_penalty = gen.repetition_penalty
        gp.n_eos = len(gen.eos_token_ids)
        for i, t in enumerate(gen.eos_token_ids):
            gp.eos_token_ids[i] = t
        sup = list(gen.suppress_tokens or [])
        self._sup_keep = _ints(sup, C.c_int32)
        gp.n_suppress = len(sup)
        gp.suppress_tokens = C.cast(self._sup_keep, C.POINTER(C.c_int32))
        gp.pin_prefix = pin_prefix
        return gp

    def generate(self, sids: Sequence[int], ids: Sequence[Sequence[int]], speech_slots: Sequence[Sequence[int]],
                 enc_ids: Sequence[Sequence[int]], gen, pin_prefix: int = 0, max_new: Optional[int] = None,
                 forced: Optional[Sequence[Sequence[int]]] = None) -> List[List[int]]:
        """Prefill each stream's turn prompt and greedy-decode.  Returns all chosen tokens per stream
        (the last one is never forwarded: drop-last rule)."""
        n = len(sids)
        gp = self._gen_params(gen, max_new, pin_prefix)
        mn = gp.max_new_tokens
        flat = [t for row in ids for t in row]
        slots = [t for row in speech_slots for t in row]
        eflat = [t for row in enc_ids for t in row]
        out = (C.c_int32 * (n * mn))()
        cnt = (C.c_int * n)()
        fptr = None
        if forced is not None:
            ff = [int(forced[b][s]) if s < len(forced[b]) else 0 for b in range(n) for s in range(mn)]
            fptr = _ints(ff, C.c_int32)
        _lib.check(self.lib.isst_generate(self.h, n, _ints(sids), _ints(flat, C.c_int32), _ints([len(r) for r in ids]),
                                          _ints(slots, C.c_int32), _ints(eflat, C.c_int32),
                                          _ints([len(r) for r in enc_ids]), C.byref(gp), fptr, out, cnt,
                                          self._stream_ptr()))
        return [[out[b * mn + s] for s in range(cnt[b])] for b in range(n)]

    def generate_beam(self, sids: Sequence[int], ids: Sequence[Sequence[int]], speech_slots: Sequence[Sequence[int]],
                      enc_ids: Sequence[Sequence[int]], gen, num_beams: int, pin_prefix: int = 0,
                      max_new: Optional[int] = None, length_penalty: float = 1.0, follow: Optional[Sequence[dict]] = None,
                      want_trace: bool = False):
        """Prefill each stream's turn prompt once and beam-search `num_beams` continuations (patch_hf.py:687-967);
        the stream's KV continues from the best hypothesis (agents/infinisst.py:334-336).  Returns per stream the
        generated part of `sequences` (closing EOS appended when it fits) and the sequence scores; with
        `want_trace` also every step's candidates / chosen beams.  `follow`: per stream {"steps": [{"closed":
        [(parent, tok)], "next": [(parent, tok)]}], "done": bool} teacher-forces the discrete choices."""
        n, k = len(sids), num_beams
        gp = self._gen_params(gen, max_new, pin_prefix)
        mn = gp.max_new_tokens
        n_keep = max(2, 1 + len(gen.eos_token_ids)) * k
        flat = [t for row in ids for t in row]
        slots = [t for row in speech_slots for t in row]
        eflat = [t for row in enc_ids for t in row]
        out = (C.c_int32 * (n * (mn + 1)))()
        cnt = (C.c_int * n)()
        scores = (C.c_float * n)()
        fptr = None
        if follow is not None:
            closed = [-1] * (n * mn * k * 2)
            nxt = [0] * (n * mn * k * 2)
            for b, f in enumerate(follow):
                for s_, st in enumerate(f["steps"]):
                    base = ((b * mn + s_) * k) * 2
                    for j, (par, tok) in enumerate(st["closed"]):
                        closed[base + 2 * j], closed[base + 2 * j + 1] = par, tok
                    for j, (par, tok) in enumerate(st["next"]):
                        nxt[base + 2 * j], nxt[base + 2 * j + 1] = par, tok
            keep = (_ints(closed, C.c_int32), _ints(nxt, C.c_int32), _ints([len(f["steps"]) for f in follow], C.c_int32),
                    _ints([int(f["done"]) for f in follow], C.c_int32))
            fs = _lib.IsstBeamFollow(*[C.cast(x, C.POINTER(C.c_int32)) for x in keep])
            fptr = C.byref(fs)
        tptr, tr = None, None
        if want_trace:
            tr = ((C.c_float * (n * mn * n_keep))(), (C.c_int32 * (n * mn * n_keep))(), (C.c_int32 * (n * mn * k * 2))(),
                  (C.c_float * (n * mn * k))(), (C.c_int32 * n)())
            ts = _lib.IsstBeamTrace(C.cast(tr[0], C.POINTER(C.c_float)), C.cast(tr[1], C.POINTER(C.c_int32)),
                                    C.cast(tr[2], C.POINTER(C.c_int32)), C.cast(tr[3], C.POINTER(C.c_float)),
                                    C.cast(tr[4], C.POINTER(C.c_int32)))
            tptr = C.byref(ts)
        _lib.check(self.lib.isst_generate_beam(self.h, n, _ints(sids), _ints(flat, C.c_int32),
                                               _ints([len(r) for r in ids]), _ints(slots, C.c_int32),
                                               _ints(eflat, C.c_int32), _ints([len(r) for r in enc_ids]), C.byref(gp),
                                               k, length_penalty, fptr, out, cnt, scores, tptr, self._stream_ptr()))
        toks = [[out[b * (mn + 1) + s_] for s_ in range(cnt[b])] for b in range(n)]
        sc = [scores[b] for b in range(n)]
        if not want_trace:
            return toks, sc
        V = self.cfg.llm.vocab
        trace = []
        for b in range(n):
            steps = []
            for s_ in range(tr[4][b]):
                o = (b * mn + s_) * n_keep
                cand = [(tr[0][o + j], tr[1][o + j] // V, tr[1][o + j] % V) for j in range(n_keep)]
                o2 = (b * mn + s_) * k
                steps.append({"cand": cand,
                              "next": [(tr[2][(o2 + j) * 2], tr[2][(o2 + j) * 2 + 1]) for j in range(k)],
                              "scores": [tr[3][o2 + j] for j in range(k)]})
            trace.append(steps)
        return toks, sc, trace

    def forward(self, sids: Sequence[int], ids: Optional[Sequence[Sequence[int]]], speech_slots=None,
                embeds: Optional[torch.Tensor] = None, lens: Optional[Sequence[int]] = None,
                pin_prefix: int = 0, all_positions: bool = False) -> torch.Tensor:
        """Append tokens (or given embeddings) to the stream caches; last-position logits [n, vocab] f32, or with
        `all_positions` the logits of every position, packed [sum(lens), vocab] (the reference's forward shape)."""
        n = len(sids)
        if ids is not None:
            lens = [len(r) for r in ids]
            flat = _ints([t for row in ids for t in row], C.c_int32)
        else:
            flat = None
        slots = None
        if speech_slots is not None:
            slots = _ints([t for row in speech_slots for t in row], C.c_int32)
        eptr = None
        if embeds is not None:
            embeds = embeds.to(device=f"cuda:{self.device}", dtype=torch.bfloat16).contiguous()
            eptr = C.c_void_p(embeds.data_ptr())
        rows = sum(lens) if all_positions else n
        out = torch.empty(rows, self.cfg.llm.vocab, dtype=torch.float32, device=f"cuda:{self.device}")
        fn = self.lib.isst_forward_all if all_positions else self.lib.isst_forward
        _lib.check(fn(self.h, n, _ints(sids), flat, _ints(lens), slots, eptr, pin_prefix,
                      C.c_void_p(out.data_ptr()), self._stream_ptr()))
        return out

    # ------------------------------------------------------------------ per-kernel-class timing
    def profile(self, on: bool = True) -> None:
        _lib.check(self.lib.isst_profile_enable(self.h, int(on)))

    def profile_reset(self) -> None:
        _lib.check(self.lib.isst_profile_reset(self.h))

    def profile_read(self) -> Dict[str, dict]:
        """{class: {launches, ms, flops, bytes}} - event time and algorithmic work per kernel class."""
        out, i = {}, 0
        name = C.create_string_buffer(64)
        n, ms, fl, by = C.c_int64(), C.c_double(), C.c_double(), C.c_double()
        while True:
            st = self.lib.isst_profile_read(self.h, i, name, 64, C.byref(n), C.byref(ms), C.byref(fl), C.byref(by))
            if st == 1:
                break
            _lib.check(st)
            out[name.value.decode()] = {"launches": n.value, "ms": ms.value, "flops": fl.value, "bytes": by.value}
            i += 1
        return out

    # ------------------------------------------------------------------ debug taps
    def debug(self, on=True) -> None:
        """bit 0: keep taps (per-step logits, encoder stages); bit 1: per-CTA phase stamps of GEMM launches."""
        _lib.check(self.lib.isst_debug_enable(self.h, int(on)))

    def read_tap(self, name: str, dtype=torch.bfloat16) -> torch.Tensor:
        nb = C.c_int64()
        _lib.check(self.lib.isst_debug_read(self.h, name.encode(), None, 0, C.byref(nb)))
        buf = torch.empty(nb.value, dtype=torch.uint8)
        _lib.check(self.lib.isst_debug_read(self.h, name.encode(), C.c_void_p(buf.data_ptr()), nb.value, C.byref(nb)))
        return buf.view(dtype)

    # ------------------------------------------------------------------ stand-alone operators
    def op_gemm(self, act: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None, gelu: bool = False,
                resid: Optional[torch.Tensor] = None, dual: bool = False, out_f32: bool = False,
                force_swap: int = -1, force_splits: int = 0) -> torch.Tensor:
        M, K = act.shape
        N = w.shape[0] // (2 if dual else 1)
        out = torch.empty(M, N, dtype=torch.float32 if out_f32 else torch.bfloat16, device=act.device)
        _lib.check(self.lib.isst_op_gemm(
            self.h, C.c_void_p(act.data_ptr()), C.c_void_p(w.data_ptr()), M, N, K,
            C.c_void_p(bias.data_ptr()) if bias is not None else None, int(gelu),
            C.c_void_p(resid.data_ptr()) if resid is not None else None, int(dual), C.c_void_p(out.data_ptr()),
            int(out_f32), force_swap, force_splits, self._stream_ptr()))
        return out

    def decode_attention_bench(self, n: int, L: int, iters: int = 20) -> float:
        ms = C.c_float()
        _lib.check(self.lib.isst_op_decode_attention_bench(self.h, n, L, iters, C.byref(ms), self._stream_ptr()))
        return ms.value
