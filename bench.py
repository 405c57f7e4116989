#!/usr/bin/env python
"""bench.py - the per-chunk streaming step on B200 (BASELINE.json metric: speech-seconds translated
per wall-second; p50/p99 per-chunk latency).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU path (oracle port)

One "step" = one 960 ms chunk of EVERY stream owned by the rank: chunked wav2vec2-large encoder ->
length adapter -> Llama-3.1-8B chunk-prefill (22 tokens) -> greedy decode (<= 10 tokens) ->
sliding-window KV eviction.  Streams are independent, so ranks share nothing (model replicated,
no collective on the data path): weak scaling, `--streams` per GPU fixed.
The timed region starts at steady state (both sliding windows full: encoder 576 frames, LLM
~1000 tokens), reached by running real chunks first ("prime", untimed).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "speech_seconds_translated_per_second"
UNIT = "speech-s/s"
CHUNK_S = 0.96
CHUNK = 15360


def peaks():
    """Roofline denominators: MEASURED_PEAKS.json (driver-written) else the documented fallback."""
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"hbm_gbs": p["hbm_gbs"], "tflops": p["bf16_tflops"], "tflops_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "tflops": 1590.0, "tflops_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


# --------------------------------------------------------------------------------------------
# clocks: nvidia-smi sampled in the background during the timed region
# --------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu: int):
        self.gpu, self.proc, self.path = gpu, None, None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "100"], stdout=open(self.path, "w"),
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for nm, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        os.unlink(self.path)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "power_w_max": max(pw), "samples": len(sm),
                "reasons": sorted(reasons)}


def pct(vals, q):
    v = sorted(vals)
    if not v:
        return None
    i = min(len(v) - 1, max(0, int(round(q * (len(v) - 1)))))
    return v[i]


# --------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the oracle (plain-PyTorch restatement of the reference's operator
# sequence, oracle/infinisst_oracle.py) on the host cores, production dimensions, steady state
# --------------------------------------------------------------------------------------------
def _aliased_state_dict(cfg):
    """Production-shaped fp32 random weights.  Layers of one stack alias the tensors of layer 0 (and
    embed_tokens aliases lm_head): the arithmetic and the bytes each layer streams are unchanged
    (one Llama layer is 0.87 GB fp32, far beyond any cache), but the host only has to sample and
    hold 3.3 GB instead of 30 GB.  Timing-only weights: parity runs never use this."""
    from infinisst_b200.synthetic import weight_specs
    import re
    g = torch.Generator().manual_seed(0)
    sd, base = {}, {}
    for key, shape, kind, std in weight_specs(cfg):
        canon = re.sub(r"(encoder\.layers|model\.layers)\.\d+\.", r"\1.0.", key)
        if canon == "model.embed_tokens.weight":
            canon = "lm_head.weight"
        if canon in base:
            sd[key] = base[canon]
            continue
        if kind == "normal":
            n = 1
            for d in shape:
                n *= d
            if n > (1 << 22):      # big matrices: tile a 4M-sample block (sampling 0.75 G normals serially is slow)
                blk = torch.randn(1 << 22, generator=g) * std
                t = blk.repeat((n + blk.numel() - 1) // blk.numel())[:n].reshape(shape).contiguous()
            else:
                t = torch.randn(shape, generator=g) * std
        elif kind == "ones":
            t = torch.ones(shape) + std * torch.randn(shape, generator=g)
        elif kind == "zeros":
            t = torch.zeros(shape)
        else:
            n = shape[0]
            t = 1.0 / (std ** (torch.arange(0, 2 * n, 2, dtype=torch.float32) / (2 * n)))
        base[canon] = t
        sd[key] = t
    if "lm_head.weight" in base and "model.embed_tokens.weight" not in sd:
        sd["model.embed_tokens.weight"] = base["lm_head.weight"]
    return sd


def _steady_oracle_stream(cfg, sd, O, device="cpu", dtype=torch.float32, seed=1):
    """A StreamState whose caches sit at the steady-state lengths (encoder window 576 frames + audio
    ring, LLM window just below the eviction threshold) with random contents: what a stream looks
    like after ~35 chunks, without paying 35 CPU chunks to get there."""
    g = torch.Generator().manual_seed(seed)
    e, l, gen = cfg.enc, cfg.llm, cfg.gen

    def rnd(*shape, scale=0.5):
        return (torch.randn(*shape, generator=g) * scale).to(device=device, dtype=dtype)
    st = O.StreamState()
    sys_n = len(cfg.tpl.system_ids)
    cur, chunks = 0, 0
    while True:                                   # integer timeline of agents/infinisst.py:337-352
        cur += (len(cfg.tpl.system_ids) + 21 if chunks == 0 else 22) + gen.max_new_tokens - 1
        kept = O.evict(st.evict_state, cur, gen.max_llm_cache_size, gen.always_cache_system_prompt, sys_n)
        chunks += 1
        if kept is not None:
            cur = kept[0] + kept[1]
            break
    st.system_size = sys_n
    st.llm_cache = O.LlmCache([rnd(1, l.kv_heads, cur, l.head_dim) for _ in range(l.layers)],
                              [rnd(1, l.kv_heads, cur, l.head_dim) for _ in range(l.layers)])
    st.enc_cache = O.new_enc_cache(e)
    st.enc_cache.n_steps = e.block_size * chunks
    st.enc_cache.src = rnd(1, 79 + 320 + 320 * e.block_size, scale=0.1)
    st.enc_cache.src_len = e.block_size
    for lc in st.enc_cache.layers:
        lc.k = rnd(e.heads, e.max_cache_size, e.head_dim)
        lc.v = rnd(e.heads, e.max_cache_size, e.head_dim)
    st.src_len = CHUNK * chunks
    st.target_ids = [1000 + (7 * i) % 5000 for i in range(9 * chunks)]
    return st, chunks, cur


def run_oracle_cpu(n_chunks: int, warm: int, threads: int):
    """Times `n_chunks` steady-state chunks of ONE stream through the oracle on `threads` host cores.
    Returns (speech-s/s, per-chunk seconds, description)."""
    from infinisst_b200 import production_config
    from oracle import infinisst_oracle as O       # bench.py's reference / cpu_baseline leg may run the oracle
    torch.set_num_threads(threads)
    cfg = production_config()
    t0 = time.perf_counter()
    sd = _aliased_state_dict(cfg)
    st, chunks0, cur = _steady_oracle_stream(cfg, sd, O)
    setup_s = time.perf_counter() - t0
    g = torch.Generator().manual_seed(2)
    times = []
    with torch.inference_mode():
        for c in range(warm + n_chunks):
            new = (0.1 * torch.randn(CHUNK, generator=g)).tolist()
            source = [0.0] * st.src_len + new      # only source[src_len:] is read (agents/infinisst.py:208)
            t1 = time.perf_counter()
            O.policy_chunk(sd, cfg, st, source, torch.float32)
            dt = time.perf_counter() - t1
            if c >= warm:
                times.append(dt)
    total = sum(times)
    desc = (f"oracle port (PyTorch eager fp32, reference operator sequence incl. cat-grown caches and per-call key "
            f"re-rotation), 1 stream x {n_chunks} steady-state chunks (LLM KV {cur} tokens, encoder window "
            f"{cfg.enc.max_cache_size} frames; timing-only weights: layers alias layer 0's tensors, cache contents random), "
            f"{threads} threads, setup {setup_s:.1f}s untimed")
    return CHUNK_S * n_chunks / total, times, desc


def run_oracle_gpu_eager(dev: int, n_chunks: int = 4, n_sub: int = 4, streams: int = 64):
    """The "reference GPU path" stand-in of SURVEY §2.1 / §8d / BASELINE.md §3: the oracle - the reference's own operator
    sequence in plain PyTorch (cat-grown caches, per-call re-rotation of every cached key, repeat_kv, lm_head over all
    prompt positions) - executed in bf16 eager ON THE SAME B200, production dimensions, steady state.  It is a
    PyTorch-eager restatement, NOT the reference (whose dependencies are absent); cuBLAS / ATen kernels
    run here, none of this repo's.  Timed: one stream for `n_chunks` chunks, and `n_sub` distinct streams advanced one
    after the other per chunk round (the reference has no batching across streams: SURVEY §0), scaled to `streams`."""
    from infinisst_b200 import production_config
    from infinisst_b200.synthetic import make_state_dict
    from oracle import infinisst_oracle as O       # bench.py's reference legs may run the oracle
    device = f"cuda:{dev}"
    cfg = production_config()
    sd = make_state_dict(cfg, seed=0, device=device, dtype=torch.bfloat16)
    sts = []
    for i in range(n_sub):
        st, chunks0, cur = _steady_oracle_stream(cfg, sd, O, device=device, dtype=torch.bfloat16, seed=10 + i)
        st.src_len = CHUNK                           # only source[src_len:] is read (agents/infinisst.py:208)
        sts.append(st)
    g = torch.Generator().manual_seed(3)

    def chunk(st):
        source = [0.0] * st.src_len + (0.1 * torch.randn(CHUNK, generator=g)).tolist()
        O.policy_chunk(sd, cfg, st, source, torch.bfloat16)
        st.src_len = CHUNK

    with torch.inference_mode():
        chunk(sts[0])                                 # warm-up (cuBLAS handles, allocator)
        torch.cuda.synchronize()
        one = []
        for _ in range(n_chunks):
            t0 = time.perf_counter()
            chunk(sts[0])
            torch.cuda.synchronize()
            one.append(time.perf_counter() - t0)
        for st in sts[1:]:
            chunk(st)
        torch.cuda.synchronize()
        rounds = []
        for _ in range(max(2, n_chunks // 2)):
            t0 = time.perf_counter()
            for st in sts:
                chunk(st)
            torch.cuda.synchronize()
            rounds.append(time.perf_counter() - t0)
    kv = sts[0].llm_cache.length()
    del sd, sts
    torch.cuda.empty_cache()
    per_round = sum(rounds) / len(rounds)
    return {"label": "PyTorch-eager restatement of the reference (oracle) in bf16 on this B200 - not the reference itself",
            "one_stream": {"speech_s_per_s": CHUNK_S * len(one) / sum(one), "p50_ms": 1e3 * pct(one, 0.5), "chunks": len(one)},
            "looped_streams": {"streams_timed": n_sub, "ms_per_round_of_timed_streams": 1e3 * per_round,
                               "speech_s_per_s": CHUNK_S * n_sub / per_round,
                               "note": f"streams advance one after the other (no cross-stream batching in the reference); "
                                       f"{streams} streams would take {1e3 * per_round * streams / n_sub:.0f} ms per chunk round at "
                                       f"the same rate"},
            "kv_len": kv, "dtype": "bf16"}


def main_reference(args, env):
    if env["rank"] != 0:
        return 0
    threads = os.cpu_count() or 1
    val, times, desc = run_oracle_cpu(n_chunks=max(1, args.steps), warm=min(args.warmup, 1), threads=threads)
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": min(args.warmup, 1), "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "wav2vec2-large + Llama-3.1-8B, 960 ms chunks, greedy <=10 tokens, steady state; "
                               "reference CPU path = oracle port, each step = 1 stream x 1 chunk (bounded sample of "
                               "the 64-stream step)", "streams": 1},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": desc},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "latency": {"p50_ms": 1e3 * pct(times, 0.5), "p99_ms": 1e3 * pct(times, 0.99), "streams": 1},
    }
    emit(line)
    return 0


# --------------------------------------------------------------------------------------------
# native arm
# --------------------------------------------------------------------------------------------
def main_native(args, env):
    from infinisst_b200 import production_config, stream_parallel as sp
    from infinisst_b200.engine import Engine
    from infinisst_b200.runner import LockstepRunner
    from infinisst_b200.synthetic import make_state_dict

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: infinisst_b200 has no CPU fallback "
                         "(use --impl reference for the CPU path)")
    rank, world, dev = env["rank"], env["world"], env["local_rank"]
    torch.cuda.set_device(dev)
    sp.init("nccl", device=dev)
    S = args.streams
    if args.total_streams:
        # BASELINE.json configs[4] literally: a fixed population of streams partitioned over the GPUs (strong scaling)
        if args.total_streams % world:
            raise SystemExit("--total-streams must be a multiple of the GPU count")
        S = args.total_streams // world
    cfg = production_config()
    log = (lambda *a: print(*a, file=sys.stderr, flush=True)) if rank == 0 else (lambda *a: None)

    t0 = time.perf_counter()
    K = args.beam
    # + scratch streams of the stand-alone kernel bench (skipped above 256 streams: state for 2 x 512 streams does
    # not fit 180 GB) and the latency stream
    S_alone = S if S <= 256 else 0
    eng = Engine(cfg, device=dev, max_streams=S + S_alone + 2, max_batch=S * K, max_prompt=64, max_beams=K)
    sd = make_state_dict(cfg, seed=0, device=f"cuda:{dev}", dtype=torch.bfloat16)
    eng.load_state_dict(sd)
    del sd
    torch.cuda.empty_cache()
    for kv in args.engine_opt:                       # A/B aid: per-context options of the library (isst_debug_option)
        k, v = kv.split("=")
        eng.option(k, int(v))
    log(f"[bench] model ready in {time.perf_counter() - t0:.1f}s")

    run = LockstepRunner(eng, cfg, S, beam=K)
    n_prime = args.prime
    n_total = n_prime + args.warmup + args.steps + 1 + args.steps + 2 + 2
    gen = torch.Generator(device=f"cuda:{dev}").manual_seed(998244353 + rank)
    pcm_dev = 0.1 * torch.randn(S, n_total * CHUNK, device=f"cuda:{dev}", generator=gen)
    pcm_host = torch.empty(S, CHUNK, dtype=torch.float32).pin_memory()
    chunk_idx = [0]

    def next_pcm():
        c = chunk_idx[0]
        chunk_idx[0] += 1
        x = pcm_dev[:, c * CHUNK:(c + 1) * CHUNK]
        if c == 0:
            x = torch.cat([torch.zeros(S, 399, device=x.device), x], 1)
        return x.contiguous()

    # ---- prime to steady state (untimed) ----
    t0 = time.perf_counter()
    for _ in range(n_prime):
        run.step_device(next_pcm())
    torch.cuda.synchronize()
    kv0 = eng.kv_len(run.sids[0])
    log(f"[bench] primed {n_prime} chunks in {time.perf_counter() - t0:.1f}s: kv_len={kv0}, evictions={run.evictions}, "
        f"enc_steps={eng.enc_steps(run.sids[0])}")

    # ---- value: inputs resident in HBM ----
    for _ in range(args.warmup):
        run.step_device(next_pcm())
    if args.ncu_step:
        # profiling aid (never a bench value): one steady-state step inside a cudaProfilerStart/Stop range,
        # for `ncu --profile-from-start off ... python bench.py --ncu-step`
        x = next_pcm()
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        run.step_device(x)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        log(f"[bench] ncu step done, launches so far {eng.launch_count()}")
        eng.close()
        return 0
    if args.timeline:
        # profiling aid (never a bench value): CUPTI kernel timeline of one steady-state step, with the real
        # overlaps of programmatic dependent launch (ncu serialises kernels and cannot show them)
        from torch.profiler import ProfilerActivity, profile
        x = next_pcm()
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            run.step_device(x)
            torch.cuda.synchronize()
        evs = sorted((e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA),
                     key=lambda e: e.time_range.start)
        t0 = evs[0].time_range.start
        with open(args.timeline, "w") as f:
            f.write("# start_us dur_us name   (one steady-state step, 64 streams; CUPTI activity records)\n")
            for e in evs:
                f.write(f"{e.time_range.start - t0:.2f} {e.time_range.end - e.time_range.start:.2f} {e.name[:60]}\n")
        log(f"[bench] timeline of {len(evs)} kernels written to {args.timeline}")
        eng.close()
        return 0
    inputs = [next_pcm() for _ in range(args.steps)]
    clocks = ClockSampler(dev)
    torch.cuda.synchronize()
    sp.barrier()
    clocks.start()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    launches0 = eng.launch_count()
    n_tok = 0
    ev[0].record()
    for i in range(args.steps):
        outs = run.step_device(inputs[i])
        n_tok += sum(len(t) for t in run.last_tokens)
        ev[i + 1].record()
    torch.cuda.synchronize()
    sp.barrier()
    clk = clocks.stop()
    launches = eng.launch_count() - launches0
    step_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]
    total_ms = ev[0].elapsed_time(ev[args.steps])
    red = sp.reduce_stats(total_ms, S * args.steps * CHUNK_S, device=f"cuda:{dev}")
    value = red["units"] / (red["ms"] / 1e3)

    # ---- e2e: the reference-facing call (model.generate), pinned host audio in, host token ids out ----
    pcm_host.copy_(next_pcm())
    run.step_api(pcm_host)                                   # warm the API path
    e2e_inputs = [next_pcm().cpu() for _ in range(args.steps)]
    torch.cuda.synchronize()
    sp.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2e_lat = []
    e0.record()
    for i in range(args.steps):
        pcm_host.copy_(e2e_inputs[i])                        # the caller's audio arrives in pinned host memory
        t1 = time.perf_counter()
        run.step_api(pcm_host)                               # H2D of the audio + D2H of the tokens inside
        e2e_lat.append(time.perf_counter() - t1)
    e1.record()
    torch.cuda.synchronize()
    sp.barrier()
    e2e_red = sp.reduce_stats(e0.elapsed_time(e1), S * args.steps * CHUNK_S, device=f"cuda:{dev}")
    e2e_value = e2e_red["units"] / (e2e_red["ms"] / 1e3)
    h2d = S * CHUNK * 4 + S * 22 * 4 * 2 + S * 100 * 4       # audio + prompt ids/slots + n-gram history
    d2h = S * cfg.gen.max_new_tokens * 4 + S * 4

    # ---- roofline: per-kernel-class CUDA-event timing over 2 more steps ----
    eng.profile(True)
    eng.profile_reset()
    for _ in range(2):
        run.step_device(next_pcm())
    torch.cuda.synchronize()
    prof = eng.profile_read()
    eng.profile(False)
    pk = peaks()
    classes = {}
    tot_ms = sum(v["ms"] for v in prof.values()) or 1.0
    for name, v in prof.items():
        if v["launches"] == 0:
            continue
        # the roofline that binds a class is the one its algorithmic work needs longer for at the measured peaks:
        # flops / tensor peak vs bytes / HBM peak (the chunk attentions move 48-88 flop per KV byte: HBM-bound)
        sec = v["ms"] / 1e3
        t_tensor = v["flops"] / (pk["tflops_sustained"] * 1e12)
        t_hbm = v["bytes"] / (pk["hbm_gbs"] * 1e9)
        b = "tensor" if t_tensor > t_hbm else "hbm"
        ach = (v["flops"] / sec / 1e12) if b == "tensor" else (v["bytes"] / sec / 1e9)
        peak = pk["tflops_sustained"] if b == "tensor" else pk["hbm_gbs"]
        classes[name] = {"bound": b, "achieved": ach, "peak": peak, "unit": "TFLOP/s" if b == "tensor" else "GB/s",
                         "frac": ach / peak, "launches_per_step": v["launches"] / 2, "ms_per_step": v["ms"] / 2,
                         "share": v["ms"] / tot_ms}
    dom = max(classes, key=lambda k: classes[k]["ms_per_step"])
    roof = {k: classes[dom][k] for k in ("bound", "achieved", "peak", "unit", "frac")}
    # measured DRAM traffic per launch of the dominant class, from the committed `ncu --set full` capture
    # (profiles/traffic.json, written by tools/make_profiles.py; ncu cannot run inside a bench)
    traffic, traffic_src, traffic_alg = None, None, None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            tj = json.load(f)
        if dom in tj:
            traffic, traffic_src = tj[dom]["dram_bytes_per_launch"], tj[dom]["source"]
            traffic_alg = tj[dom].get("algorithmic_bytes_per_launch_same_launches")
    except (OSError, ValueError, KeyError):
        pass
    v_dom = prof[dom]
    roof.update({"kernel": dom, "traffic": traffic, "traffic_unit": "bytes per launch (mean over the captured launches)",
                 "traffic_source": traffic_src, "traffic_algorithmic_same_launches": traffic_alg,
                 "algorithmic_per_launch": (v_dom["flops"] if roof["bound"] == "tensor" else v_dom["bytes"]) / max(v_dom["launches"], 1),
                 "peaks": pk["source"],
                 "note": "algorithmic bytes/flops per launch (DESIGN.md §4) / CUDA-event time of the launches of this "
                         "class inside the step; sustained bf16 peak for tensor-bound classes (timed inside a long step)"})

    # ---- stand-alone decode attention (the HBM-bound kernel the north star names) ----
    dec = {}
    try:
        if not S_alone:
            raise RuntimeError("skipped: no scratch streams at this batch size")
        ms = eng.decode_attention_bench(S, 1000, 32)
        by = S * 1001 * cfg.llm.kv_heads * cfg.llm.head_dim * 2 * 2
        dec = {"streams": S, "kv_len": 1000, "ms": ms, "achieved_gbs": by / ms / 1e6, "frac": by / ms / 1e6 / pk["hbm_gbs"]}
    except Exception as ex:          # noqa: BLE001
        dec = {"error": str(ex)}

    # ---- single-stream per-chunk latency (BASELINE.json configs[1]) through the same API ----
    lat = {}
    if args.latency_chunks > 0:
        one = LockstepRunner(eng, cfg, 1, beam=K)
        g1 = torch.Generator().manual_seed(7)
        buf = torch.empty(1, CHUNK, dtype=torch.float32).pin_memory()
        first = torch.cat([torch.zeros(1, 399), 0.1 * torch.randn(1, CHUNK, generator=g1)], 1).pin_memory()
        one.step_api(first)
        for _ in range(n_prime):
            buf.copy_(0.1 * torch.randn(1, CHUNK, generator=g1))
            one.step_api(buf)
        ts = []
        for _ in range(args.latency_chunks):
            buf.copy_(0.1 * torch.randn(1, CHUNK, generator=g1))
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            one.step_api(buf)
            torch.cuda.synchronize()
            ts.append(1e3 * (time.perf_counter() - t1))
        lat = {"streams": 1, "chunks": len(ts), "p50_ms": pct(ts, 0.5), "p99_ms": pct(ts, 0.99),
               "kv_len": eng.kv_len(one.sids[0]), "speech_s_per_s": CHUNK_S / (sum(ts) / len(ts) / 1e3)}
        one.close()
    all_e2e = sp.gather_floats([1e3 * t for t in e2e_lat])
    eager = None
    if rank == 0 and world == 1 and args.eager_chunks > 0:
        try:
            eager = run_oracle_gpu_eager(dev, n_chunks=args.eager_chunks)
            eager["native_over_eager_one_stream"] = lat.get("speech_s_per_s", 0.0) / eager["one_stream"]["speech_s_per_s"] if lat else None
            eager["native_64_streams_over_eager_looped"] = e2e_value / eager["looped_streams"]["speech_s_per_s"]
        except Exception as ex:          # noqa: BLE001
            eager = {"error": str(ex)}

    if rank != 0:
        eng.close()
        sp.shutdown()
        return 0
    cpu = None
    if world == 1 and args.cpu_baseline_chunks > 0:
        threads = os.cpu_count() or 1
        v, times, desc = run_oracle_cpu(args.cpu_baseline_chunks, 1, threads)
        cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": desc}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": red["ms"] / args.steps, "higher_is_better": True, "scaling": "strong" if args.total_streams else "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": {"workload": f"BASELINE.json configs[2] per GPU ({S} concurrent streams, batched chunk-prefill + decode; "
                               f"x{world} GPUs = configs[4] partition): wav2vec2-large + Llama-3.1-8B bf16 random-init, "
                               "960 ms chunks, 22-token turn prompt, " + (f"beam search ({K} beams, KV hand-back)" if K > 1 else "greedy") +
                               " <= 10 tokens, max_llm_cache_size 1000 + "
                               "pinned 40-token system prompt, steady state",
                   "streams_per_gpu": S, "streams_total": S * world, "parallelism": f"stream-parallel replicas x{world}, "
                   "no collective on the data path", "kv_len_at_start": kv0, "prime_chunks": n_prime,
                   "l2": "each step streams 15 GB of weights (>> 126 MB L2): inputs larger than L2, no flush needed",
                   "tokens_generated_per_step": n_tok / args.steps},
        "clocks": clk,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "api": "SpeechLlamaForCausalLM.generate (agents/infinisst.py:307-332 kwargs) + kv_evict",
                "batch_chunk_latency_ms": {"p50": pct(all_e2e, 0.5), "p99": pct(all_e2e, 0.99)}},
        "gpu_launches": launches,
        "roofline": roof,
        "kernel_classes": classes,
        "decode_attention_standalone": dec,
        "latency": lat,
        "step_ms": step_ms,
    }
    if cpu is not None:
        line["cpu_baseline"] = cpu
    if eager is not None:
        line["extra"] = {"reference_gpu_eager": eager}
    emit(line)
    eng.close()
    sp.shutdown()
    return 0


_JSON_OUT = None


def _claim_stdout():
    """The contract is ONE JSON line on stdout: keep a private handle on it and point fd 1 at stderr so that
    library chatter (NCCL version banners, warnings) cannot get in front of the line."""
    global _JSON_OUT
    if _JSON_OUT is None:
        sys.stdout.flush()
        _JSON_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line: dict):
    _JSON_OUT.write(json.dumps(line) + "\n")
    _JSON_OUT.flush()


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--streams", type=int, default=64, help="concurrent streams per GPU")
    ap.add_argument("--total-streams", type=int, default=0, help="fixed stream population split over the GPUs (configs[4]: 512); overrides --streams")
    ap.add_argument("--beam", type=int, default=1, help="1 = greedy (the north-star workload); k > 1 = the reference's beam search")
    ap.add_argument("--prime", type=int, default=34, help="untimed chunks run first so both sliding windows are full")
    ap.add_argument("--latency-chunks", type=int, default=20, help="single-stream latency sample (0 = skip)")
    ap.add_argument("--timeline", default="", help="write the CUPTI kernel timeline of one step to this file and exit")
    ap.add_argument("--ncu-step", action="store_true", help="run ONE profiled step after priming and exit (for ncu)")
    ap.add_argument("--cpu-baseline-chunks", type=int, default=3, help="oracle chunks timed on the host at N=1 (0 = skip)")
    ap.add_argument("--engine-opt", action="append", default=[], help="key=value per-context option (A/B runs), e.g. decode_splits=4")
    ap.add_argument("--eager-chunks", type=int, default=4, help="oracle chunks timed in bf16 eager on the GPU at N=1 (0 = skip)")
    args = ap.parse_args()
    from infinisst_b200 import stream_parallel as sp
    env = sp.env_world()
    if env["world"] != args.gpus and env["world"] > 1:
        print(f"[bench] warning: --gpus {args.gpus} but WORLD_SIZE={env['world']}", file=sys.stderr)
    if args.impl == "reference":
        return main_reference(args, env)
    return main_native(args, env)


if __name__ == "__main__":
    sys.exit(main())
