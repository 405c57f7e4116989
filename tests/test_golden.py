"""The oracle against its committed golden vectors (tests/golden/make_golden.py): any edit to
oracle/infinisst_oracle.py that changes results is caught here, on CPU."""
import os

import numpy as np
import pytest
import torch

from infinisst_b200 import tiny_config
from infinisst_b200.synthetic import make_audio, make_state_dict
from parity_utils import OracleStream, bf16_weights

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tiny_stream.npz")


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def test_oracle_reproduces_golden_stream(gold):
    n = int(gold["n_chunks"])
    cfg = tiny_config(max_cache_size=int(gold["max_cache"]), max_llm_cache_size=int(gold["max_llm"]))
    sd = bf16_weights(make_state_dict(cfg, seed=0))
    audio = make_audio(n * 15360 / 16000.0)
    orc = OracleStream(cfg, sd)
    evictions = 0
    for c in range(n):
        out_ids, rec, taps = orc.chunk(audio[: (c + 1) * 15360].tolist())
        np.testing.assert_allclose(taps["speech_feats"][0].numpy(), gold[f"c{c}_speech_feats"], atol=2e-4, rtol=1e-3)
        got = torch.stack([l[0] for l in rec.step_logits]).numpy()
        np.testing.assert_allclose(got, gold[f"c{c}_step_logits"], atol=2e-3, rtol=1e-3)
        # integer results are bit-exact
        assert rec.sequences[0] == gold[f"c{c}_sequence"].tolist()
        assert out_ids == gold[f"c{c}_output_ids"].tolist()
        log = orc.st.kv_log[-1]
        kept = log["kept"] if log["kept"] is not None else (-1, -1)
        assert [log["cur"], kept[0], kept[1], log["after"]] == gold[f"c{c}_kv"].tolist()
        evictions += log["kept"] is not None
    assert evictions >= 3, "the golden stream must exercise sliding-window eviction"
