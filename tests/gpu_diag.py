"""Verbose GPU diagnostics (not a pytest file): prints error metrics for every stage so one
gpurun call yields as much information as possible.  Usage: python tests/gpu_diag.py <stage>"""
import os
import sys
import time
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch

from infinisst_b200 import tiny_config
from infinisst_b200.engine import Engine
from infinisst_b200.synthetic import make_audio, make_state_dict
from oracle import infinisst_oracle as O
from parity_utils import OracleStream, bf16_weights, max_abs, rel_l2, slot_map


def stage_gemm(impl):
    cfg = tiny_config()
    eng = Engine(cfg, max_streams=2)
    torch.manual_seed(0)
    dev = "cuda:0"
    cases = [
        # M, N, K, kwargs
        (128, 128, 64, {}), (128, 128, 256, {}), (256, 384, 512, {}), (100, 200, 192, {}),
        (3072, 1024, 1024, {}), (1408, 4096, 4096, {}),
        (1, 128, 64, {}), (1, 4096, 4096, {}), (16, 256, 512, {}), (22, 6144, 4096, {}), (48, 3072, 1024, {}),
        (64, 1024, 4096, {}), (12, 519, 512, {"out_f32": True}), (1, 128263, 4096, {"out_f32": True}),
        (22, 4096, 4096, {"force_splits": 4}), (1, 4096, 14336, {"force_splits": 8}),
        (300, 512, 1024, {"force_splits": 3}),
        (48, 1024, 1024, {"bias": True, "resid": True}), (48, 4096, 1024, {"bias": True, "gelu": True}),
        (3072, 4096, 1024, {"bias": True, "gelu": True}), (3072, 1024, 4096, {"bias": True, "resid": True}),
        (22, 768, 512, {"dual": True}), (1, 14336, 4096, {"dual": True}), (1408, 14336, 4096, {"dual": True}),
        (64, 14336, 4096, {"dual": True, "force_splits": 2}),
        (100, 256, 512, {"force_swap": 1}), (30, 256, 512, {"force_swap": 0}),
    ]
    worst = 0.0
    for (M, N, K, kw) in cases:
        try:
            a = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
            dual = kw.get("dual", False)
            w = (torch.randn(N * (2 if dual else 1), K, device=dev) * (K ** -0.5)).bfloat16()
            bias = torch.randn(N, device=dev) if kw.get("bias") else None
            resid = torch.randn(M, N, device=dev).bfloat16() if kw.get("resid") else None
            ref = a.float() @ w.float().t()
            if dual:
                ref = torch.nn.functional.silu(ref[:, :N]) * ref[:, N:]
            if bias is not None:
                ref = ref + bias
            if kw.get("gelu"):
                ref = torch.nn.functional.gelu(ref)
            if resid is not None:
                ref = ref + resid.float()
            out = eng.op_gemm(a, w, bias=bias, gelu=kw.get("gelu", False), resid=resid, dual=dual,
                              out_f32=kw.get("out_f32", False), impl=impl, force_swap=kw.get("force_swap", -1),
                              force_splits=kw.get("force_splits", 0))
            torch.cuda.synchronize()
            err = rel_l2(out, ref)
            worst = max(worst, err)
            flag = "OK " if err < 1e-2 else "BAD"
            print(f"[gemm impl={impl}] {flag} M={M} N={N} K={K} {kw} rel_l2={err:.3e} max_abs={max_abs(out, ref):.3e}", flush=True)
        except Exception as ex:
            print(f"[gemm impl={impl}] EXC M={M} N={N} K={K} {kw}: {ex}", flush=True)
            worst = 1.0
    print(f"[gemm impl={impl}] worst rel_l2 {worst:.3e}")
    return worst < 1e-2


def stage_stream(n_chunks=4, max_cache=576, max_llm=1000):
    cfg = tiny_config(max_cache_size=max_cache, max_llm_cache_size=max_llm)
    sd = bf16_weights(make_state_dict(cfg, seed=0))
    eng = Engine(cfg, max_streams=2)
    eng.load_state_dict(sd)
    eng.debug(True)
    seg = eng.chunk_samples
    audio = make_audio(n_chunks * seg / 16000.0)
    orc = OracleStream(cfg, sd)
    sid = eng.open_stream()
    ok = True
    target_ids = []
    ck = O.EvictionState()
    for c in range(n_chunks):
        out_o, rec, taps = orc.chunk(audio[: (c + 1) * seg].tolist())
        pcm = audio[c * seg:(c + 1) * seg][None].clone()
        if c == 0:
            pcm = torch.cat([torch.zeros(1, 399), pcm], 1)
        feats = eng.encode_chunk([sid], pcm, 1, return_feats=True)
        torch.cuda.synchronize()
        for name, okey in [("enc_conv", "conv"), ("enc_post_proj", "post_proj"), ("enc_layer_0", "enc_layer_0"),
                           ("enc_layer_1", "enc_layer_1"), ("enc_out", "enc_out"), ("speech_feats", "speech_feats")]:
            got = eng.read_tap(name).float()
            ref = taps[okey].flatten()
            e = rel_l2(got[: ref.numel()], ref)
            print(f"[stream c={c}] {name:14s} rel_l2={e:.3e} max_abs={max_abs(got[:ref.numel()], ref):.3e}", flush=True)
            ok &= e < 3e-2
        ids = O.build_prompt(cfg.tpl, c == 0)
        forced = rec.sequences[0][len(ids):]
        toks = eng.generate([sid], [ids], [slot_map(cfg, ids)], [target_ids[-100:]], cfg.gen,
                            pin_prefix=len(cfg.tpl.system_ids), forced=[forced])[0]
        logits = eng.read_tap("step_logits", torch.float32).view(cfg.gen.max_new_tokens, 1, cfg.llm.vocab)
        for s in range(len(rec.step_logits)):
            e = rel_l2(logits[s, 0], rec.step_logits[s][0])
            am_g, am_o = int(logits[s, 0].argmax()), int(rec.step_logits[s][0].argmax())
            print(f"[stream c={c}] step {s} logits rel_l2={e:.3e} max_abs={max_abs(logits[s, 0], rec.step_logits[s][0]):.3e} "
                  f"argmax gpu={am_g} oracle={am_o}", flush=True)
            ok &= e < 5e-2
        print(f"[stream c={c}] tokens gpu={toks} oracle={forced} kv gpu={eng.kv_len(sid)} oracle_cur={orc.st.kv_log[-1]}", flush=True)
        target_ids.extend(out_o)
        cur = eng.kv_len(sid)
        kept = O.evict(ck, cur, cfg.gen.max_llm_cache_size, True, len(cfg.tpl.system_ids))
        if kept is not None:
            eng.kv_evict(sid, kept[0], cur - kept[1])
        ok &= eng.kv_len(sid) == orc.st.llm_cache.length()
    return ok


def main():
    stage = sys.argv[1]
    t0 = time.time()
    try:
        if stage == "gemm_simple":
            ok = stage_gemm(1)
        elif stage == "gemm_tc":
            ok = stage_gemm(0)
        elif stage == "stream":
            ok = stage_stream()
        elif stage == "stream_evict":
            ok = stage_stream(n_chunks=8, max_cache=96, max_llm=150)
        else:
            raise SystemExit("unknown stage")
    except Exception:
        traceback.print_exc()
        ok = False
    print(f"[{stage}] {'PASS' if ok else 'FAIL'} in {time.time() - t0:.1f}s  ISST_GEMM={os.environ.get('ISST_GEMM', 'tc')}", flush=True)
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
