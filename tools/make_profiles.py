"""Digest one `tests/run_gpu_round.sh` result (gpurun_out/) into the tracked artefacts under profiles/:
  <tag>_bench.json              the bench line
  <tag>_launch_list.txt         ncu launch list of one step, grouped by (kernel, grid)
  <tag>_ncu_full_<name>.txt     per-launch summary of the `ncu --set full` captures (time, DRAM bytes, throughput, occupancy limits)
  traffic.json                  measured DRAM bytes per launch of each captured kernel class (read by bench.py -> roofline.traffic)
Usage: python tools/make_profiles.py r01h          (PROF_OUT=gpurun_out/profiles to digest on the GPU box: the raw
                                                   .ncu-rep files of a full round exceed what gpurun copies back)
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.environ.get("PROF_OUT") or os.path.join(ROOT, "profiles")     # PROF_OUT: digest on the GPU box into gpurun_out/
WANT = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_uniform.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "launch__shared_mem_per_block", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "launch__waves_per_multiprocessor", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__pipe_tensor_subpipe_dense_cycles_active.avg.pct_of_peak_sustained_active"]


def to_bytes(val, unit):
    v = float(val.replace(",", ""))
    u = unit.lower()
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)


def full_summary(rep, title):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    lines = [f"# {title}", f"# source: ncu --set full --clock-control none --import-source on ({os.path.basename(rep)}; cold-cache, serialised launches)"]
    recs = []
    for r in data:
        name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "").replace("isst::", "").replace("tc::", "")
        rec = {"kernel": name, "grid": r[idx["Grid Size"]], "block": r[idx["Block Size"]]}
        lines.append(f"{name}  grid {rec['grid']} block {rec['block']}")
        for w in WANT:
            if w in idx and r[idx[w]] not in ("", "n/a"):
                lines.append(f"    {w:75s} {r[idx[w]]:>14s} {units[idx[w]]}")
                rec[w] = (r[idx[w]], units[idx[w]])
        recs.append(rec)
    return "\n".join(lines) + "\n", recs


def main(tag):
    os.makedirs(PROF, exist_ok=True)
    bench = os.path.join(OUT, "bench.json")
    tl = os.path.join(OUT, "timeline.txt")
    if os.path.exists(tl):
        with open(tl) as f, open(os.path.join(PROF, f"{tag}_timeline_step64.txt"), "w") as g:
            g.write(f.read())
    if os.path.exists(bench):
        with open(bench) as f, open(os.path.join(PROF, f"{tag}_bench.json"), "w") as g:
            g.write(f.read())
    beam = os.path.join(OUT, "bench_beam4.json")
    if os.path.exists(beam):
        with open(beam) as f, open(os.path.join(PROF, f"{tag}_bench_beam4_64streams.json"), "w") as g:
            g.write(f.read())
    ll = os.path.join(OUT, "launches.csv")
    if os.path.exists(ll):
        hdr = ("# ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off python bench.py --ncu-step --warmup 1\n"
               "# one steady-state step: 64 streams x 960 ms chunk (encoder + 22-token prefill + 9 decode forwards), kv_len 1001\n")
        txt = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summarize.py"), ll, hdr.replace("\n", "\\n")],
                             capture_output=True, text=True).stdout
        open(os.path.join(PROF, f"{tag}_launch_list.txt"), "w").write(txt)
    traffic = {}
    for rep, name, title, cls in [("prof_gemm_decode.ncu-rep", "decode_chain", "fused decode-layer chain, 64 token rows: one launch = o_proj -> RMSNorm -> gate/up -> down -> RMSNorm -> QKV of the next layer (436 MB of weights)", "gemm_stream"),
                                  ("prof_gemm_prefill.ncu-rep", "gemm_prefill", "tensor-bound GEMMs of one prefill layer (64 streams x 22 = 1408 tokens: o_proj / gate-up / down / qkv order as captured)", "gemm_tensor"),
                                  ("prof_decode_attn.ncu-rep", "decode_attention", "decode attention, 64 streams x kv_len ~1001, layers 0-1 of one decode forward", "attn_decode"),
                                  ("prof_prefill_attn.ncu-rep", "prefill_attention", "tcgen05 chunk-prefill attention, 64 streams x 22 tokens over kv_len ~1023", "attn_prefill"),
                                  ("prof_group_attn.ncu-rep", "beam_group_attention", "beam search shared-prefix decode attention, 64 sentences x 4 beams, prefix ~1010 keys + 4 private tails", "attn_decode_beam4")]:
        p = os.path.join(OUT, rep)
        if not os.path.exists(p):
            continue
        txt, recs = full_summary(p, title)
        open(os.path.join(PROF, f"{tag}_ncu_full_{name}.txt"), "w").write(txt)
        tot = [to_bytes(*r["dram__bytes_read.sum"]) + to_bytes(*r["dram__bytes_write.sum"]) for r in recs
               if "dram__bytes_read.sum" in r and "dram__bytes_write.sum" in r]
        if tot:
            # algorithmic bytes of the SAME captured launches (production dims), so that traffic can be compared like with like
            alg = None
            if cls == "gemm_stream":      # o_proj, gate/up, down, qkv weights of a decode layer (bf16) per chain launch
                alg = (4096 * 4096 + 2 * 14336 * 4096 + 4096 * 14336 + 6144 * 4096) * 2
            elif cls == "gemm_tensor":    # the same four GEMMs at 1408 tokens: weights + activations in + outputs (+ residual reads)
                M = 1408
                w = (6144 * 4096 + 4096 * 4096 + 2 * 14336 * 4096 + 4096 * 14336) * 2
                io = M * 2 * ((4096 + 6144) + (4096 + 2 * 4096) + (4096 + 14336) + (14336 + 2 * 4096))
                alg = (w + io) / 4.0
            elif cls == "attn_decode_beam4":  # shared prefix once per sentence + one private page pair per beam (<= 32 keys)
                alg = 64 * (1008 + 4 * 20) * 8 * 128 * 2 * 2
            elif cls in ("attn_decode", "attn_prefill"):   # K and V of 64 streams x ~1001-1023 tokens x 8 kv heads x 128 x bf16
                alg = 64 * (1001 if cls == "attn_decode" else 1023) * 8 * 128 * 2 * 2
            traffic[cls] = {"dram_bytes_per_launch": sum(tot) / len(tot), "launches_captured": len(tot),
                            "algorithmic_bytes_per_launch_same_launches": alg,
                            "per_launch": [round(t) for t in tot],
                            "source": f"profiles/{tag}_ncu_full_{name}.txt (dram__bytes_read.sum + dram__bytes_write.sum, ncu --set full)"}
    if traffic:
        json.dump(traffic, open(os.path.join(PROF, "traffic.json"), "w"), indent=1)
    print("wrote", sorted(os.listdir(PROF)))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "r01")
