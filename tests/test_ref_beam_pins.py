"""Beam search (the reference's shipped decoding, `--beam 4`) pinned on the REFERENCE'S OWN loop and scorer:
tests/golden/make_ref_beam_pins.py executes model/patches/patch_hf.py (`generation_mixin_beam_search`,
`beam_search_process`, `beam_search_finalize`, `beam_hypotheses_add`, `_expand_inputs_for_generation`) and the
agent's beam branch (agents/infinisst.py:334-336) from /root/reference on the tiny configuration; the stand-ins
for the transformers 4.47 classes they extend are listed in tests/golden/ref_standins.py.

CPU (`-m "not gpu"`): the oracle's `generate_beam` must reproduce every step's 2k candidates (scores, tokens,
parent beams), the beams kept, the hypothesis counts, the early-`done` step, the returned sequence, the emitted
ids and the KV lengths (hand-back of the best hypothesis + eviction).
GPU (`-m gpu`): the CUDA path must reproduce the same stream (tests/test_gpu_beam.py).
"""
import os

import numpy as np
import pytest
import torch

from infinisst_b200 import tiny_config
from infinisst_b200.synthetic import make_audio, make_state_dict
from oracle import infinisst_oracle as O
from parity_utils import bf16_weights

PINS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_tiny_beam.npz")
SEG = 15360


@pytest.fixture(scope="module")
def pins():
    return np.load(PINS)


def beam_weights(cfg, eos_scale: float):
    """Scenario weights of make_ref_beam_pins.py: EOS rows of lm_head scaled so EOS candidates appear."""
    sd = bf16_weights(make_state_dict(cfg, seed=0))
    if eos_scale != 1.0:
        w = sd["lm_head.weight"].clone()
        w[cfg.gen.eos_token_ids] = (w[cfg.gen.eos_token_ids].float() * eos_scale).to(w.dtype)
        sd["lm_head.weight"] = w
    return sd


@pytest.mark.parametrize("name", ["plain", "eos"])
def test_oracle_reproduces_reference_beam_search(pins, name):
    n, k = int(pins[f"{name}_n_chunks"]), int(pins["beam"])
    cfg = tiny_config(max_cache_size=int(pins["max_cache"]), max_llm_cache_size=int(pins["max_llm"]))
    cfg.gen.beam = k
    sd = beam_weights(cfg, float(pins[f"{name}_eos_scale"]))
    audio = make_audio(n * SEG / 16000.0)
    st = O.StreamState()
    closed_by_eos = early_done = evictions = 0
    for c in range(n):
        p = f"{name}_c{c}_"
        out_ids, rec = O.policy_chunk(sd, cfg, st, audio[: (c + 1) * SEG].tolist())
        steps = pins[p + "cand_scores"].shape[0]
        assert len(rec.trace) == steps                                       # same stopping step
        for s, tr in enumerate(rec.trace):
            np.testing.assert_allclose([x[0] for x in tr["cand"]], pins[p + "cand_scores"][s], atol=2e-4, rtol=1e-4)
            assert [x[1] for x in tr["cand"]] == pins[p + "cand_beams"][s].tolist()
            assert [x[2] for x in tr["cand"]] == pins[p + "cand_tokens"][s].tolist()
            assert [x[0] for x in tr["next"]] == pins[p + "next_beams"][s].tolist()
            assert [x[1] for x in tr["next"]] == pins[p + "next_tokens"][s].tolist()
            np.testing.assert_allclose(tr["scores"], pins[p + "next_scores"][s], atol=2e-4, rtol=1e-4)
            assert tr["done"] == bool(pins[p + "done"][s])
            closed_by_eos += len(tr["closed"])
        early_done += rec.trace[-1]["done"]
        assert rec.sequences[0] == pins[p + "sequence"].tolist()
        assert out_ids == pins[p + "output_ids"].tolist()
        log = st.kv_log[-1]
        _, hyp_kv, after = pins[p + "kv"].tolist()
        assert (log["cur"], log["after"]) == (hyp_kv, after)                 # KV hand-back + eviction
        evictions += log["kept"] is not None
    assert evictions >= 2
    if name == "eos":
        assert closed_by_eos >= 10 and early_done >= 2                       # the EOS paths were exercised


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="the reference tree only exists in the build container")
def test_beam_pins_regenerate_from_reference(pins, tmp_path):
    """In the build container: re-run the reference's beam search and require the committed fixture to be what it
    produces."""
    import subprocess
    import sys
    script = os.path.join(os.path.dirname(PINS), "make_ref_beam_pins.py")
    env = dict(os.environ, REF_PINS_OUT=str(tmp_path / "beam.npz"))
    r = subprocess.run([sys.executable, script], env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    fresh = np.load(tmp_path / "beam.npz")
    for name in ("plain", "eos"):
        for c in range(int(pins[f"{name}_n_chunks"])):
            p = f"{name}_c{c}_"
            assert fresh[p + "sequence"].tolist() == pins[p + "sequence"].tolist()
            assert fresh[p + "cand_tokens"].tolist() == pins[p + "cand_tokens"].tolist()
            np.testing.assert_allclose(fresh[p + "cand_scores"], pins[p + "cand_scores"], atol=1e-4)
            assert fresh[p + "kv"].tolist() == pins[p + "kv"].tolist()
