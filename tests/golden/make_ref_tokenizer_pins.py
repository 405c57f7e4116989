"""Generates tests/golden/ref_tokenizer_pins.npz: the reference agent (agents/infinisst.py, executed from /root/reference
under the stand-ins of ref_standins.py, as make_ref_pins.py does) driven with a REAL tokenizers-backed HF tokenizer and
Jinja chat template (tests/golden/llama3_style_tokenizer, make_tokenizer_fixture.py) instead of the synthetic template:

  * the reference's OWN `SpeechLlamaForCausalLM.preprocess` (model/llm.py:149-190) adds the 7 speech / latency tokens to
    the tokenizer and records the ids the splice scans for;
  * `_prepare_inputs` (agents/infinisst.py:225-268) builds every prompt through `tokenizer.apply_chat_template`, the
    `[:, :-1]` cut and the Llama-3.1 `[:, 25:]` strip - and, for a model name without "3.1", the `input_ids[:, 0] = eos`
    branch (:265-266);
  * the stream runs 4 chunks: prompts, per-step logits, sequences, KV lengths.
tests/test_tokenizer_pins.py holds the product (agent host logic on the CPU, CUDA path on the GPU) against this."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, HERE)

from make_ref_pins import build_reference_agent                      # noqa: E402
from infinisst_b200 import tiny_config                                # noqa: E402
from infinisst_b200.synthetic import make_audio, make_state_dict     # noqa: E402
from parity_utils import bf16_weights                                 # noqa: E402

TOK_DIR = os.path.join(HERE, "llama3_style_tokenizer")
N_CHUNKS = 4


def load_tokenizer():
    """The fixture through transformers' own fast-tokenizer class.  One version stand-in: transformers 4.47 (what the
    reference pins, README.md:32) returns the id tensor from `apply_chat_template(..., return_tensors="pt")`;
    transformers 5.x returns a BatchEncoding unless `return_dict=False` - the reference calls `.size(1)` / slices on it."""
    import transformers

    class Tokenizer447(transformers.PreTrainedTokenizerFast):
        def apply_chat_template(self, *a, **k):
            k.setdefault("return_dict", False)
            return super().apply_chat_template(*a, **k)
    return Tokenizer447.from_pretrained(TOK_DIR, padding_side="right")


def eos_ids(tok):
    return [int(tok.convert_tokens_to_ids(t)) for t in ("<|end_of_text|>", "<|eom_id|>", "<|eot_id|>")]


def main():
    torch.set_num_threads(4)
    out = {}
    for tag, model_name in (("l31", "synthetic-llama-3.1"), ("l3", "synthetic-llama-3")):
        cfg = tiny_config(max_cache_size=96, max_llm_cache_size=150)
        tok = load_tokenizer()
        cfg.gen.eos_token_ids = eos_ids(tok)
        sd = bf16_weights(make_state_dict(cfg, seed=0))
        agent, taps = build_reference_agent(cfg, sd, tokenizer=tok, model_name=model_name)
        mc = agent.model.config
        out[f"{tag}_config_ids"] = np.array([mc.sp_patch_token_id, mc.user_token_id, mc.assist_token_id, mc.start_header_id,
                                             len(tok), tok.pad_token_id], dtype=np.int32)
        out[f"{tag}_llama31"] = np.int32(agent.llama31)
        if tag == "l3":
            # prompts only: the non-3.1 branch keeps the whole default system header and overwrites position 0
            st = agent.build_states()
            st.reset()
            out["l3_first"] = agent._prepare_inputs(st)[0].numpy().astype(np.int32)
            st.speech_cache = object()
            out["l3_later"] = agent._prepare_inputs(st)[0].numpy().astype(np.int32)
            continue
        seg = 15360
        audio = make_audio(N_CHUNKS * seg / 16000.0)
        st = agent.build_states()
        st.reset()
        st.source_sample_rate = 16000
        out["n_chunks"] = np.int32(N_CHUNKS)
        for c in range(N_CHUNKS):
            st.source = audio[: (c + 1) * seg].tolist()
            kv_before = 0 if st.past_key_values is None else st.past_key_values[0][0].size(2)
            agent.policy(st)
            gen = taps["gen"]
            n_prompt = gen.sequences.size(1) - len(gen.step_logits)
            out[f"c{c}_prompt"] = gen.sequences[0, :n_prompt].numpy().astype(np.int32)
            out[f"c{c}_sequence"] = gen.sequences[0].numpy().astype(np.int32)
            out[f"c{c}_step_logits"] = torch.stack([x[0] for x in gen.step_logits]).numpy().astype(np.float32)
            out[f"c{c}_speech_feats"] = taps["speech_feats"][0].numpy().astype(np.float32)
            out[f"c{c}_kv"] = np.array([kv_before + gen.sequences.size(1) - 1, st.past_key_values[0][0].size(2)], dtype=np.int32)
            print(f"chunk {c}: prompt {n_prompt} tokens, kv {out[f'c{c}_kv'].tolist()}, emitted {gen.sequences[0, n_prompt:-1].tolist()}")
        out["system_prompt_size"] = np.int32(agent.system_prompt_size)
        # latency multiplier 2: the prompt carries 24 <sp_patch> slots and <latency_2> in the system turn
        agent.update_multiplier(2)
        st2 = agent.build_states()
        st2.reset()
        out["m2_first"] = agent._prepare_inputs(st2)[0].numpy().astype(np.int32)
        agent.update_multiplier(1)
    path = os.environ.get("REF_PINS_OUT") or os.path.join(HERE, "ref_tokenizer_pins.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
