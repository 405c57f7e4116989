"""Cross-checks of the oracle against the HF 5.5 modules in this image (SURVEY §4 item 5,
§8c pin (ii)).  The Llama / wav2vec2 block arithmetic is unchanged between 4.47 and 5.5."""
import pytest
import torch

from oracle import infinisst_oracle as O

transformers = pytest.importorskip("transformers")


def _hf_llama(cfg, sd):
    from transformers import LlamaConfig, LlamaForCausalLM
    l = cfg.llm
    hc = LlamaConfig(vocab_size=l.vocab, hidden_size=l.hidden, intermediate_size=l.ffn,
                     num_hidden_layers=l.layers, num_attention_heads=l.heads,
                     num_key_value_heads=l.kv_heads, head_dim=l.head_dim, rms_norm_eps=l.rms_eps,
                     rope_theta=l.rope_theta, max_position_embeddings=131072,
                     rope_scaling=dict(rope_type="llama3", **l.rope_scaling),
                     attention_bias=False, tie_word_embeddings=False)
    m = LlamaForCausalLM(hc).eval()
    own = {k: v for k, v in sd.items() if k.startswith("model.layers") or k in
           ("model.embed_tokens.weight", "model.norm.weight", "lm_head.weight")}
    missing, unexpected = m.load_state_dict(own, strict=False)
    assert not unexpected and all("rotary" in k or "inv_freq" in k for k in missing), (missing, unexpected)
    return m


def test_llama3_inv_freq_matches_hf(tiny_cfg, tiny_sd):
    m = _hf_llama(tiny_cfg, tiny_sd)
    torch.testing.assert_close(O.llama_inv_freq(tiny_cfg.llm), m.model.rotary_emb.inv_freq.float())


def test_llama_full_pass_matches_hf(tiny_cfg, tiny_sd):
    m = _hf_llama(tiny_cfg, tiny_sd)
    torch.manual_seed(0)
    emb = torch.randn(2, 37, tiny_cfg.llm.hidden)
    with torch.no_grad():
        ref = m(inputs_embeds=emb).logits
    got = O.llama_forward(tiny_sd, tiny_cfg.llm, emb, O.LlmCache.empty(tiny_cfg.llm.layers))
    torch.testing.assert_close(got, ref, atol=3e-4, rtol=1e-3)


def test_llama_incremental_matches_hf_cache(tiny_cfg, tiny_sd):
    """Un-rotated cache + re-rotation at 0..L-1 == HF's rotated DynamicCache when nothing is evicted."""
    m = _hf_llama(tiny_cfg, tiny_sd)
    torch.manual_seed(1)
    emb = torch.randn(1, 30, tiny_cfg.llm.hidden)
    with torch.no_grad():
        o1 = m(inputs_embeds=emb[:, :22], use_cache=True)
        o2 = m(inputs_embeds=emb[:, 22:23], past_key_values=o1.past_key_values, use_cache=True)
    cache = O.LlmCache.empty(tiny_cfg.llm.layers)
    O.llama_forward(tiny_sd, tiny_cfg.llm, emb[:, :22], cache)
    got = O.llama_forward(tiny_sd, tiny_cfg.llm, emb[:, 22:23], cache)
    torch.testing.assert_close(got, o2.logits, atol=3e-4, rtol=1e-3)


def test_conv_extractor_matches_hf_layernorm_mode(tiny_cfg, tiny_sd):
    from transformers import Wav2Vec2Config
    from transformers.models.wav2vec2.modeling_wav2vec2 import Wav2Vec2FeatureEncoder
    e = tiny_cfg.enc
    hc = Wav2Vec2Config(feat_extract_norm="layer", conv_dim=[c for c, _, _ in e.conv_layers],
                        conv_kernel=[k for _, k, _ in e.conv_layers], conv_stride=[s for _, _, s in e.conv_layers],
                        conv_bias=True, feat_extract_activation="gelu",
                        num_feat_extract_layers=len(e.conv_layers))
    fe = Wav2Vec2FeatureEncoder(hc).eval()
    for j, layer in enumerate(fe.conv_layers):
        p = f"{O.ENC}feature_extractor.conv_layers.{j}."
        layer.conv.weight.data.copy_(tiny_sd[p + "0.weight"])
        layer.conv.bias.data.copy_(tiny_sd[p + "0.bias"])
        layer.layer_norm.weight.data.copy_(tiny_sd[p + "2.1.weight"])
        layer.layer_norm.bias.data.copy_(tiny_sd[p + "2.1.bias"])
    torch.manual_seed(0)
    wav = 0.1 * torch.randn(2, 15759)
    with torch.no_grad():
        ref = fe(wav)
    got = O.conv_feature_extractor(tiny_sd, e, wav)
    assert got.shape == ref.shape == (2, e.conv_dim, 48)
    torch.testing.assert_close(got, ref, atol=1e-5, rtol=1e-4)
    assert O.feat_extract_output_length(e, 15759, with_adapter=False) == 48
    assert O.feat_extract_output_length(e, 15759) == 12


def test_logits_processors_match_hf():
    from transformers.generation.logits_process import (
        EncoderNoRepeatNGramLogitsProcessor, NoRepeatNGramLogitsProcessor,
        RepetitionPenaltyLogitsProcessor, SuppressTokensLogitsProcessor)

    class G:
        repetition_penalty = 1.2
        no_repeat_ngram_size = 3
        suppress_tokens = [5, 17]

    g = torch.Generator().manual_seed(0)
    V = 50
    for trial in range(30):
        ids = torch.randint(0, 8, (1, 25), generator=g)
        enc = torch.randint(0, 8, (1, 40), generator=g)
        scores = torch.randn(1, V, generator=g)
        ref = scores.clone()
        for proc in (RepetitionPenaltyLogitsProcessor(1.2), NoRepeatNGramLogitsProcessor(3),
                     EncoderNoRepeatNGramLogitsProcessor(3, enc),
                     SuppressTokensLogitsProcessor([5, 17], device="cpu")):
            ref = proc(ids, ref)
        got = O.process_logits(scores[0], ids[0].tolist(), enc[0].tolist(), G)
        assert torch.equal(got, ref[0]), trial
    # empty encoder ids (first chunk: agents/infinisst.py:298-301)
    got = O.process_logits(scores[0], ids[0].tolist(), [], G)
    assert torch.isfinite(got).sum() > 0
