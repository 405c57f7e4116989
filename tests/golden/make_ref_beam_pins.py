"""Generates tests/golden/ref_tiny_beam.npz by EXECUTING THE REFERENCE'S OWN beam search from /root/reference.

Same set-up as make_ref_pins.py (the reference agent, model, encoder and attention patches run unmodified on the
tiny configuration), but with the agent's shipped decoding mode, `--beam 4`:

    InfiniSST.policy (agents/infinisst.py:270-394), beam branch (:334-336)
      -> model.generate(num_beams=4)  [HF generate set-up: stand-in, tests/golden/ref_standins.py::beam_generate]
           -> generation_mixin_expand_inputs_for_generation   (model/patches/patch_hf.py:305-342)   reference
           -> generation_mixin_beam_search                    (:687-967)                             reference
                -> beam_search_process / beam_hypotheses_add  (:43-157, :278-302)                    reference
                -> beam_search_finalize                       (:159-275)                             reference
      -> KV hand-back of the best hypothesis, eviction, drop-last output slicing (:334-363)          reference

Two scenarios: the plain synthetic weights (hypotheses are closed at max length), and the same weights with the
EOS rows of lm_head scaled so that EOS candidates appear (hypotheses closed by EOS, early `done`).  Recorded per
chunk: the 2k candidates every step as the scorer receives them (scores, tokens, parent beams), the beams it
chose, the returned sequence, emitted ids, KV length before / after eviction.

The fixture cannot be regenerated on the GPU box (/root/reference does not exist there): it is committed.
Run:  python tests/golden/make_ref_beam_pins.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, HERE)

import ref_standins as RS                                             # noqa: E402
import make_ref_pins as MP                                            # noqa: E402
from infinisst_b200 import tiny_config                                 # noqa: E402
from infinisst_b200.synthetic import make_audio, make_state_dict      # noqa: E402
from parity_utils import bf16_weights                                  # noqa: E402

BEAM = 4
SCENARIOS = [("plain", 1.0, 8), ("eos", 3.0, 6)]      # (name, scale of the EOS rows of lm_head, chunks)


def scenario_weights(cfg, eos_scale: float):
    sd = bf16_weights(make_state_dict(cfg, seed=0))
    if eos_scale != 1.0:
        w = sd["lm_head.weight"].clone()
        w[cfg.gen.eos_token_ids] = (w[cfg.gen.eos_token_ids].float() * eos_scale).to(w.dtype)
        sd["lm_head.weight"] = w
    return sd


def run_scenario(name, eos_scale, n_chunks, out):
    cfg = tiny_config(max_cache_size=MP.MAX_CACHE, max_llm_cache_size=MP.MAX_LLM)
    g = cfg.gen
    sd = scenario_weights(cfg, eos_scale)
    agent, taps = MP.build_reference_agent(cfg, sd)
    agent.beam = BEAM
    agent.cache_checkpoints = []
    rec_steps = []

    def generate(**kw):
        rec_steps.clear()
        res = RS.beam_generate(agent.model, eos_token_ids=g.eos_token_ids, record=rec_steps, **kw)
        taps["gen"] = res
        taps["hyp_kv"] = res.past_key_values[0][0][0].size(2)     # best hypothesis' KV length, before the agent evicts
        return res
    agent.model.generate = generate
    seg = 15360
    audio = make_audio(n_chunks * seg / 16000.0)
    states = agent.build_states()
    states.reset()
    states.source_sample_rate = 16000
    out[f"{name}_n_chunks"] = np.int32(n_chunks)
    out[f"{name}_eos_scale"] = np.float32(eos_scale)
    for c in range(n_chunks):
        states.source = audio[: (c + 1) * seg].tolist()
        kv_before = 0 if states.past_key_values is None else states.past_key_values[0][0].size(2)
        n_target = len(states.target_ids)
        agent.policy(states)
        res = taps["gen"]
        seq = res.sequences[0]
        hyp_kv = taps["hyp_kv"]
        after = states.past_key_values[0][0].size(2)
        p = f"{name}_c{c}_"
        out[p + "sequence"] = seq.numpy().astype(np.int32)
        out[p + "output_ids"] = np.array(states.target_ids[n_target:], dtype=np.int32)
        out[p + "kv"] = np.array([kv_before, hyp_kv, after], dtype=np.int32)
        out[p + "cand_scores"] = torch.stack([r["scores"] for r in rec_steps]).numpy().astype(np.float32)
        out[p + "cand_tokens"] = torch.stack([r["tokens"] for r in rec_steps]).numpy().astype(np.int32)
        out[p + "cand_beams"] = torch.stack([r["beams"] for r in rec_steps]).numpy().astype(np.int32)
        out[p + "next_scores"] = torch.stack([r["next_scores"] for r in rec_steps]).numpy().astype(np.float32)
        out[p + "next_tokens"] = torch.stack([r["next_tokens"] for r in rec_steps]).numpy().astype(np.int32)
        out[p + "next_beams"] = torch.stack([r["next_beams"] for r in rec_steps]).numpy().astype(np.int32)
        out[p + "n_hyps"] = np.array([r["n_hyps"] for r in rec_steps], dtype=np.int32)
        out[p + "done"] = np.array([r["done"] for r in rec_steps], dtype=np.int32)
        out[p + "speech_feats"] = taps["speech_feats"][0].numpy().astype(np.float32)
        print(f"[{name}] chunk {c}: kv {kv_before} -> {hyp_kv} -> {after}, steps {len(rec_steps)}, "
              f"hyps/step {out[p + 'n_hyps'].tolist()}, emitted {out[p + 'output_ids'].tolist()}")


def main():
    torch.set_num_threads(4)
    out = {"beam": np.int32(BEAM), "max_cache": np.int32(MP.MAX_CACHE), "max_llm": np.int32(MP.MAX_LLM)}
    for name, scale, n in SCENARIOS:
        run_scenario(name, scale, n, out)
    path = os.environ.get("REF_PINS_OUT") or os.path.join(HERE, "ref_tiny_beam.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
