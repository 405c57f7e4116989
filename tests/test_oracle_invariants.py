"""Pins for the oracle itself (the reference has no tests: SURVEY §4, §8c).

Each test is a property of the reference's own two code paths (training/offline vs
inference/streaming) restated in oracle/infinisst_oracle.py."""
import copy

import pytest
import torch

from infinisst_b200 import tiny_config
from infinisst_b200.synthetic import make_audio, make_state_dict
from oracle import infinisst_oracle as O


@pytest.mark.parametrize("bs", [16, 48, 96])
@pytest.mark.parametrize("cache", [100, 576])
@pytest.mark.parametrize("seq", [24, 48, 96])
def test_mask_closed_form(bs, cache, seq):
    """both reference mask builders equal one closed form."""
    for prefix in [0, 48, 96, 100, 528, 576, 624, 1008, 4800]:
        if prefix == 0:
            ref = O.mask_offline(seq, cache, bs)
        else:
            ref = O.mask_streaming(seq, prefix, cache, bs)
        assert torch.equal(ref, O.mask_closed_form(seq, prefix, cache, bs)), (bs, cache, seq, prefix)


def test_conv_chunked_equals_full(tiny_cfg, tiny_sd):
    """patch_speech_encoder.py:254-264: the ring keeps exactly the context the conv stack
    needs, so chunked frames == frames of the whole utterance (layer_norm mode is per-frame)."""
    seg = O.chunk_samples(tiny_cfg.enc)
    n_chunks = 4
    wav = make_audio(n_chunks * seg / 16000.0)
    full = torch.cat([torch.zeros(79 + 320), wav])[None]
    ref = O.conv_feature_extractor(tiny_sd, tiny_cfg.enc, full)          # [1, C, 48*n]
    assert ref.shape[-1] == 48 * n_chunks
    cache = O.new_enc_cache(tiny_cfg.enc)
    got = []
    for c in range(n_chunks):
        taps = {}
        x = wav[c * seg:(c + 1) * seg]
        if c == 0:
            x = torch.cat([torch.zeros(79 + 320), x])
        O.w2v2_streaming_forward(tiny_sd, tiny_cfg.enc, x[None], cache, tiny_cfg.enc.block_size, taps)
        got.append(taps["conv"])
        assert cache.src.shape[1] == 79 + 320 + 320 * 48
    got = torch.cat(got, dim=1).transpose(1, 2)
    torch.testing.assert_close(got, ref, atol=2e-5, rtol=1e-4)


@pytest.mark.parametrize("max_cache", [96, 576])
def test_encoder_streaming_equals_offline(max_cache):
    """Training path (whole utterance, mask_offline, empty cache) == inference path (chunks,
    mask_streaming, KV cache); RoPE is relative so window-relative positions agree."""
    cfg = tiny_config(max_cache_size=max_cache)
    sd = make_state_dict(cfg, seed=1)
    seg = O.chunk_samples(cfg.enc)
    n_chunks = 5
    wav = torch.cat([torch.zeros(79 + 320), make_audio(n_chunks * seg / 16000.0)])
    off, _ = O.encode_speech(sd, cfg.enc, wav[None], None)
    cache, outs = None, []
    for c in range(n_chunks):
        lo = 0 if c == 0 else 399 + c * seg
        x, cache = O.encode_speech(sd, cfg.enc, wav[None, lo: 399 + (c + 1) * seg], cache)
        outs.append(x)
        assert all(l.k.shape[1] <= max_cache + 48 for l in cache.layers)
    torch.testing.assert_close(torch.cat(outs, 1), off, atol=3e-4, rtol=1e-3)


def test_llm_incremental_equals_full_pass(tiny_cfg, tiny_sd):
    """Multi-turn incremental prefill + decode on the un-rotated cache == one causal pass
    (no eviction): patch_llm.py:280-299 vs the training forward."""
    torch.manual_seed(0)
    T = [61, 1, 1, 22, 1]
    emb = torch.randn(1, sum(T), tiny_cfg.llm.hidden)
    full = O.llama_forward(tiny_sd, tiny_cfg.llm, emb, O.LlmCache.empty(tiny_cfg.llm.layers))
    cache = O.LlmCache.empty(tiny_cfg.llm.layers)
    outs, s = [], 0
    for t in T:
        outs.append(O.llama_forward(tiny_sd, tiny_cfg.llm, emb[:, s:s + t], cache))
        s += t
    assert cache.length() == sum(T)
    torch.testing.assert_close(torch.cat(outs, 1), full, atol=2e-4, rtol=1e-3)


def test_eviction_shifts_positions(tiny_cfg, tiny_sd):
    """After eviction the kept keys are re-rotated at positions 0..L-1 (patch_llm.py:287-291):
    decoding on an evicted cache == decoding on a cache built from only the kept tokens."""
    torch.manual_seed(1)
    lc = tiny_cfg.llm
    emb = torch.randn(1, 50, lc.hidden)
    cache = O.LlmCache.empty(lc.layers)
    O.llama_forward(tiny_sd, lc, emb, cache)
    O.apply_eviction(cache, (8, 20))
    assert cache.length() == 28
    nxt = torch.randn(1, 1, lc.hidden)
    a = O.llama_forward(tiny_sd, lc, nxt, copy.deepcopy(cache))
    # the K/V of a kept token depend on its *original* context, so rebuild by slicing, not re-running
    b = O.llama_forward(tiny_sd, lc, nxt, cache)
    torch.testing.assert_close(a, b)
    assert cache.length() == 29


def test_eviction_timeline_appendix_b():
    """S=40, first prompt 61, later 22, 9 forwarded generated tokens,
    max 1000, system prompt pinned."""
    st = O.EvictionState()
    cur, log = 0, []
    for chunk in range(40):
        cur += (61 if chunk == 0 else 22) + 9
        kept = O.evict(st, cur, 1000, True, 40)
        log.append((cur, kept))
        if kept is not None:
            cur = kept[0] + kept[1]
    assert log[0] == (70, None) and log[1] == (101, None) and log[30] == (1000, None)
    assert log[31] == (1031, (40, 961))
    assert all(c == 1032 and k == (40, 961) for c, k in log[32:])


def test_eviction_matches_literal_slices():
    """Integer model vs literally slicing a tensor the way agents/infinisst.py:354-361 does,
    with random turn lengths; also checks the checkpoint rebasing keeps turn boundaries."""
    import random
    rng = random.Random(0)
    for keep_sys in (True, False):
        st = O.EvictionState()
        toks = []                       # logical token ids currently in the cache
        nxt = 0
        boundaries = []
        for chunk in range(300):
            n = (61 if chunk == 0 else 22) + rng.randint(0, 9)
            toks += list(range(nxt, nxt + n))
            nxt += n
            cur = len(toks)
            kept = O.evict(st, cur, 200, keep_sys, 40)
            if kept is not None:
                p, t = kept
                toks = toks[:p] + toks[cur - t:]
                assert len(toks) <= 200 + (40 if keep_sys else 0)
                # the tail starts at a turn boundary: its first token id is a turn start
                assert toks[p] in boundaries or t == cur
            boundaries.append(nxt)
            assert st.checkpoints[-1] == len(toks) if kept is None else True
            # checkpoints are strictly increasing positions inside the cache
            assert all(0 < c <= len(toks) for c in st.checkpoints)


def test_drop_last_rule_and_kv_growth(tiny_cfg, tiny_sd):
    """10 chosen tokens, 9 forwarded; output_ids excludes the last."""
    st = O.StreamState()
    seg = O.chunk_samples(tiny_cfg.enc)
    audio = make_audio(2 * seg / 16000.0).tolist()
    out0, rec0 = O.policy_chunk(tiny_sd, tiny_cfg, st, audio[:seg])
    assert len(rec0.sequences[0]) == 61 + 10 and len(out0) == 9
    assert st.llm_cache.length() == 61 + 9
    out1, rec1 = O.policy_chunk(tiny_sd, tiny_cfg, st, audio[:2 * seg])
    assert len(rec1.sequences[0]) == 22 + 10 and st.llm_cache.length() == 70 + 22 + 9
    assert st.target_ids == out0 + out1


def test_speech_slot_map(tiny_cfg):
    ids = O.build_prompt(tiny_cfg.tpl, True)
    slot = O.speech_slot_map(tiny_cfg.llm, ids)
    pos = [t for t, s in enumerate(slot) if s >= 0]
    assert [slot[t] for t in pos] == list(range(12))
    assert all(ids[t] == tiny_cfg.tpl.sp_patch_id for t in pos)
    assert sum(1 for t in ids if t == tiny_cfg.tpl.sp_patch_id) == 12
    ids2 = O.build_prompt(tiny_cfg.tpl, False)
    assert len(ids) == 61 and len(ids2) == 22 and ids2[0] == tiny_cfg.tpl.eot_id
