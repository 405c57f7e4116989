"""Generates tests/golden/ref_tiny_enc_variants.npz by EXECUTING THE REFERENCE'S OWN encoder code from
/root/reference with the non-production encoder flags (agents/options.py:32-41):

    --xpos 1 --rope 1   RotaryEmbedding(use_xpos=True)              patch_speech_encoder.py:631, 823-824
    --xpos 0 --rope 0   sinusoidal_positional_embedding at n_steps  patch_speech_encoder.py:448-461, 488-493

The agent's own `load_model` (agents/infinisst.py:130-183) builds the model exactly as in make_ref_pins.py
(`patch_w2v2(args.xpos, args.rope)` sets the module globals); `SpeechEncoderW2V2RoPE.encode_speech`
(model/speech_encoder.py:219-236) is then driven chunk by chunk with the reference's `W2V2RoPECache`, far enough
for the 96-frame window to slide.  Reference code: everything under model/ and agents/.  Stand-ins
(tests/golden/ref_standins.py): the fairseq containers and, for xPos, the rotary_embedding_torch arithmetic
(`get_scale`; the package is absent and its version unpinned: that part of the pin is the stand-in's restatement,
the sinusoidal path is the reference's own function).

Run:  python tests/golden/make_ref_variant_pins.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_ref_pins as MP                                             # noqa: E402
from infinisst_b200 import tiny_config                                 # noqa: E402
from infinisst_b200.synthetic import make_audio                       # noqa: E402
from parity_utils import ENC_VARIANTS, variant_state_dict            # noqa: E402

N_CHUNKS = 5
MAX_CACHE = 96


def run_variant(name):
    xpos, rope, _ = ENC_VARIANTS[name]
    xpos, rope = int(xpos), int(rope)
    cfg = tiny_config(max_cache_size=MAX_CACHE)
    sd = variant_state_dict(cfg, name)
    agent, _ = MP.build_reference_agent(cfg, sd, xpos=xpos, rope=rope)
    import model.patches.patch_speech_encoder as pse
    assert (bool(pse.XPOS), bool(pse.ROPE)) == (bool(xpos), bool(rope))
    se = agent.model.model.speech_encoder
    se.set_blocksize(1)
    seg = 15360
    audio = make_audio(N_CHUNKS * seg / 16000.0)
    cache, feats = None, []
    for c in range(N_CHUNKS):
        pcm = audio[c * seg:(c + 1) * seg][None]
        if c == 0:
            pcm = torch.cat([torch.zeros(1, 79 + 320), pcm], dim=1)      # agents/infinisst.py:216-218
        with torch.no_grad():
            f, cache = se.encode_speech(pcm, cache=cache)
        feats.append(f[0].numpy().astype(np.float32))
        print(f"xpos={xpos} rope={rope} chunk {c}: n_steps {cache.n_steps}, |f| {float(f.norm()):.4f}")
    return feats


def main():
    torch.set_num_threads(4)
    out = {"n_chunks": np.int32(N_CHUNKS), "max_cache": np.int32(MAX_CACHE)}
    for name in ENC_VARIANTS:
        for c, f in enumerate(run_variant(name)):
            out[f"{name}_c{c}_speech_feats"] = f
    path = os.environ.get("REF_PINS_OUT") or os.path.join(HERE, "ref_tiny_enc_variants.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
