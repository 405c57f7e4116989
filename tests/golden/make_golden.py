"""Generates tests/golden/tiny_stream.npz from the oracle (SURVEY §8c pin (iii)).

The reference ships no golden vectors and cannot be imported here, so these are
oracle outputs, frozen so that (a) later edits to the oracle are caught by
tests/test_golden.py and (b) the CUDA path is checked against fixed numbers on the GPU box.
Run:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from infinisst_b200 import tiny_config                      # noqa: E402
from infinisst_b200.synthetic import make_audio, make_state_dict   # noqa: E402
from parity_utils import OracleStream, bf16_weights          # noqa: E402

N_CHUNKS = 8
MAX_CACHE, MAX_LLM = 96, 150          # stress variant: both windows slide inside 8 chunks


def main():
    torch.set_num_threads(4)
    cfg = tiny_config(max_cache_size=MAX_CACHE, max_llm_cache_size=MAX_LLM)
    sd = bf16_weights(make_state_dict(cfg, seed=0))
    seg = 15360
    audio = make_audio(N_CHUNKS * seg / 16000.0)
    orc = OracleStream(cfg, sd)
    out = {"n_chunks": np.int32(N_CHUNKS), "max_cache": np.int32(MAX_CACHE), "max_llm": np.int32(MAX_LLM)}
    for c in range(N_CHUNKS):
        out_ids, rec, taps = orc.chunk(audio[: (c + 1) * seg].tolist())
        out[f"c{c}_speech_feats"] = taps["speech_feats"][0].numpy().astype(np.float32)
        out[f"c{c}_enc_out"] = taps["enc_out"][0].numpy().astype(np.float16)
        out[f"c{c}_step_logits"] = torch.stack([l[0] for l in rec.step_logits]).numpy().astype(np.float32)
        out[f"c{c}_sequence"] = np.array(rec.sequences[0], dtype=np.int32)
        out[f"c{c}_output_ids"] = np.array(out_ids, dtype=np.int32)
        log = orc.st.kv_log[-1]
        kept = log["kept"] if log["kept"] is not None else (-1, -1)
        out[f"c{c}_kv"] = np.array([log["cur"], kept[0], kept[1], log["after"]], dtype=np.int32)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tiny_stream.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
