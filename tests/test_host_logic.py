"""Host-side logic of the product package against the oracle's restatement of the reference
(no GPU): eviction integers, prompt layout, speech-slot map, speech preparation, RoPE tables,
and the stream-parallel plumbing (world_size-2 gloo)."""
import argparse
import os
import random

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from infinisst_b200 import production_config, tiny_config
from infinisst_b200 import stream_parallel as sp
from infinisst_b200.agent import InfiniSST, S2TAgentStates, TemplateTokenizer, evict_plan, non_language_token_ids
from infinisst_b200.engine import llama_inv_freq
from infinisst_b200.model import SpeechLlamaForCausalLM
from oracle import infinisst_oracle as O


@pytest.mark.parametrize("keep_sys", [True, False])
@pytest.mark.parametrize("window", [150, 200, 1000])
def test_evict_plan_equals_reference_integers(keep_sys, window):
    """agents/infinisst.py:337-352: the product's plan (drop [keep_prefix, drop_upto)) and the oracle's
    literal (prefix_kept, tail_kept) describe the same kept index set, chunk after chunk."""
    rng = random.Random(window)
    st = S2TAgentStates()
    st.system_prompt_size = 40
    ost = O.EvictionState()
    cur_p = cur_o = 0
    for chunk in range(400):
        n = (61 if chunk == 0 else 22) + rng.randint(0, 9)
        cur_p += n
        cur_o += n
        plan = evict_plan(st, cur_p, window, keep_sys)
        kept = O.evict(ost, cur_o, window, keep_sys, 40)
        if kept is None or kept[1] == cur_o:
            assert plan is None
        else:
            prefix, tail = kept
            assert plan == (prefix, cur_o - tail)
            cur_o = prefix + tail
            cur_p -= plan[1] - plan[0]
        assert cur_p == cur_o and st.cache_checkpoints == ost.checkpoints


def test_template_prompt_layout_matches_oracle():
    for cfg in (tiny_config(), production_config()):
        tok = TemplateTokenizer(cfg)
        assert tok.system_ids() + tok.turn_ids(12) == O.build_prompt(cfg.tpl, True)
        assert [cfg.tpl.eot_id] + tok.turn_ids(12) == O.build_prompt(cfg.tpl, False)
        assert len(O.build_prompt(cfg.tpl, True)) == 61 and len(O.build_prompt(cfg.tpl, False)) == 22


def test_slot_map_matches_oracle():
    cfg = production_config()
    m = object.__new__(SpeechLlamaForCausalLM)
    m.cfg = cfg
    for first in (True, False):
        ids = O.build_prompt(cfg.tpl, first)
        assert m._slot_map(ids) == O.speech_slot_map(cfg.llm, ids)
    two_turns = O.build_prompt(cfg.tpl, True) + O.build_prompt(cfg.tpl, False)
    assert m._slot_map(two_turns) == O.speech_slot_map(cfg.llm, two_turns)
    assert sorted(s for s in m._slot_map(two_turns) if s >= 0) == list(range(24))


def test_prepare_speech_matches_reference_semantics():
    """agents/infinisst.py:200-223: zero-pad to a multiple of 15360, 79+320 zeros before the first chunk."""
    p = argparse.ArgumentParser()
    InfiniSST.add_args(p)
    agent = object.__new__(InfiniSST)
    agent.args = p.parse_args(["--block-size", "48"])
    agent.latency_multiplier = 1
    cfg = tiny_config()
    g = torch.Generator().manual_seed(0)
    src = torch.randn(15360 + 7000, generator=g).tolist()
    st = S2TAgentStates()
    st.source = src[:15360]
    a = agent._prepare_speech(st)
    b, n = O.prepare_speech(cfg.enc, src[:15360], 0, torch.float32)
    assert torch.equal(a, b) and st.src_len == n == 15360 and a.shape == (1, 15360 + 399)
    st.source = src                                   # second call: 7000 new samples -> padded to 15360
    a = agent._prepare_speech(st)
    b, n = O.prepare_speech(cfg.enc, src, 15360, torch.float32)
    assert torch.equal(a, b) and a.shape == (1, 15360) and st.src_len == n == len(src)
    assert float(a[0, 7000:].abs().sum()) == 0.0


def test_llama3_rope_frequencies_match_oracle():
    for cfg in (tiny_config(), production_config()):
        torch.testing.assert_close(llama_inv_freq(cfg.llm), O.llama_inv_freq(cfg.llm), rtol=0, atol=0)


def test_feat_extract_output_lengths():
    from infinisst_b200.model import SpeechEncoderW2V2RoPE

    class _E:
        cfg = production_config()
    enc = object.__new__(SpeechEncoderW2V2RoPE)
    enc.engine = _E()
    n = torch.tensor([15759, 15360 * 2 + 399, 400, 16000])
    got = enc._get_feat_extract_output_lengths(n)
    want = [O.feat_extract_output_length(_E.cfg.enc, int(x)) for x in n]
    assert got.tolist() == want and want[0] == 12 and want[1] == 24


def test_shard_streams_partition():
    for world in (1, 2, 4, 8):
        parts = [sp.shard_streams(512, world, r) for r in range(world)]
        assert sorted(s for p in parts for s in p) == list(range(512))
        assert max(map(len, parts)) - min(map(len, parts)) <= 1
    assert sp.shard_streams(5, 4, 3) == [3] and sp.shard_streams(2, 4, 3) == []
    with pytest.raises(ValueError):
        sp.shard_streams(4, 2, 2)


def _gloo_worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    sp.init("gloo")
    sp.barrier()
    mine = sp.shard_streams(9, world, rank)
    red = sp.reduce_stats(local_ms=100.0 * (rank + 1), local_units=len(mine) * 0.96)
    lat = sp.gather_floats([float(rank), float(rank) + 0.5])
    q.put((rank, red, lat, mine))
    dist.destroy_process_group()


def test_stream_parallel_world2_gloo():
    """N>1 path: time = max over ranks, units = sum over ranks, latencies gathered from all ranks."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, red, lat, mine in res:
        assert red["ms"] == 200.0 and abs(red["units"] - 9 * 0.96) < 1e-9 and red["world"] == 2
        assert lat == [0.0, 0.5, 1.0, 1.5]
    assert res[0][3] == [0, 2, 4, 6, 8] and res[1][3] == [1, 3, 5, 7]


def test_suppress_non_language_token_ids():
    """agents/infinisst.py:142-148: tokens whose text contains '(' or the full-width '\uff08' are suppressed."""
    class Tok:
        words = ["hello", " (", "a(b", "\uff08", "\uff09", ")", "x"]

        def decode(self, idx, skip_special_tokens=True):
            return self.words[idx]
    assert non_language_token_ids(Tok(), 7) == [1, 2, 3]
    assert non_language_token_ids(TemplateTokenizer(tiny_config()), 16) == []


def test_dpo_sampling_translation_log(tmp_path):
    """agents/infinisst.py:363-394: drop-last output slice, quoted per-segment translations appended to
    --output-file as one bracketed line when the source finishes, Read / Write action rule."""
    from infinisst_b200.agent import InfiniSST, ReadAction, S2TAgentStates, WriteAction

    class Tok:
        def decode(self, ids, skip_special_tokens=True):
            return " ".join(f"w{i}" for i in ids)

    agent = InfiniSST.__new__(InfiniSST)                 # host logic only: no engine, no GPU
    agent.tokenizer, agent.dpo_sampling, agent.output_file = Tok(), True, str(tmp_path / "translations.json")
    st = S2TAgentStates()
    a = agent._finish(st, 3, [9, 9, 9, 5, 6, 7])          # prompt of 3, generated 5 6 7: the last token is dropped
    assert isinstance(a, WriteAction) and a.content == "w5 w6" and not a.finished and st.target_ids == [5, 6]
    a = agent._finish(st, 3, [9, 9, 9, 7])                # a single generated token: nothing is emitted
    assert isinstance(a, ReadAction) and st.translations_list == ["'w5 w6'", "''"] and st.segment_idx == 2
    st.source_finished = True
    a = agent._finish(st, 3, [9, 9, 9, 8, 1])
    assert isinstance(a, WriteAction) and a.finished and a.content == "w8"
    assert open(agent.output_file, encoding="utf-8").read() == "['w5 w6', '', 'w8']\n"
    assert st.translations_list == []


def test_sampling_is_refused_not_ignored():
    """`--do-sample` reaches `model.generate` (agents/infinisst.py:311-315) and is refused there: the CUDA path
    implements the shipped decoding modes (greedy, beam search) and must not silently fall back to them."""
    import inspect
    from infinisst_b200.agent import InfiniSST
    from infinisst_b200.model import SpeechLlamaForCausalLM
    src = inspect.getsource(InfiniSST._generate)
    assert "do_sample=self.do_sample" in src and "temperature=self.temperature" in src
    model = SpeechLlamaForCausalLM.__new__(SpeechLlamaForCausalLM)       # no engine needed: the check comes first
    with pytest.raises(NotImplementedError, match="do_sample"):
        model.generate(input_ids=torch.zeros(1, 3, dtype=torch.long), do_sample=True)
