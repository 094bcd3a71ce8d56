#!/usr/bin/env python3
"""Turns one tools/round_gpu.sh output directory (gpurun_out/<tag>) into the tracked summaries under profiles/:
bench lines, per-launch tables, the ncu launch list (kernel shares of one step) and the ncu --set full summaries of the
attention / GEMM / memory-bound kernels, plus profiles/traffic.json (measured DRAM traffic of the dominant kernel
family, read by bench.py). usage: summarise_profiles.py gpurun_out/<tag> [round]"""
import csv
import json
import os
import shutil
import sys

src = sys.argv[1]
rnd = sys.argv[2] if len(sys.argv) > 2 else "r1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
dst = os.path.join(ROOT, "profiles")
os.makedirs(dst, exist_ok=True)


def copy(a, b):
    if os.path.exists(os.path.join(src, a)):
        shutil.copy(os.path.join(src, a), os.path.join(dst, b))
        print("wrote", b)


for a, b in [("bench_vitl_b32.json", f"{rnd}_bench_vitl_b32_504.json"), ("bench_vitl_b4.json", f"{rnd}_bench_vitl_b4_504.json"),
             ("bench_vitb.json", f"{rnd}_bench_vitb.json"), ("bench_beit_large_384.json", f"{rnd}_bench_beit_large_384.json"),
             ("bench_swinv2_large_384.json", f"{rnd}_bench_swinv2_large_384.json"), ("bench_prepost.json", f"{rnd}_bench_prepost.json"),
             ("launch_table_vitl_b32.csv", f"{rnd}_launch_table_vitl_b32_504.csv"), ("launch_table_vitb.csv", f"{rnd}_launch_table_vitb.csv"),
             ("launch_table_beit_large_384.csv", f"{rnd}_launch_table_beit_large_384.csv"),
             ("launch_table_swinv2_large_384.csv", f"{rnd}_launch_table_swinv2_large_384.csv"),
             ("bench_reference_cpu.json", f"{rnd}_bench_reference_cpu_arm.json"),
             ("bench_reference_gpu_eager.json", f"{rnd}_bench_reference_gpu_eager_vitl_b32.json"),
             ("parity_report.json", f"{rnd}_parity_report_five_configs.json"),
             ("bench_vitl_b1.json", f"{rnd}_bench_vitl_b1_504.json"), ("launch_table_vitl_b1.csv", f"{rnd}_launch_table_vitl_b1_504.csv"),
             ("launch_table_vitl_b4.csv", f"{rnd}_launch_table_vitl_b4_504.csv"),
             ("bench_vits_b1.json", f"{rnd}_bench_vits_b1_504.json"), ("bench_vits_b32.json", f"{rnd}_bench_vits_b32_504.json"),
             ("bench_vitg_b16.json", f"{rnd}_bench_vitg_B16.json"), ("launch_table_vitg_b16.csv", f"{rnd}_launch_table_vitg_B16.csv")]:
    copy(a, b)


def read_ncu_csv(path):
    rows = [r for r in csv.reader(open(path)) if r]
    # skip ncu's ==PROF== preamble lines
    start = next(i for i, r in enumerate(rows) if r and r[0] in ("ID", '"ID"'))
    return rows[start], rows[start + 1:]


# ---- launch list: kernel family shares of the captured steps
p = os.path.join(src, "ncu_launches.csv")
if os.path.exists(p):
    hdr, rows = read_ncu_csv(p)
    ik, im, iv = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
    iu = hdr.index("Metric Unit")
    agg = {}
    for r in rows:
        if r[im] != "gpu__time_duration.sum":
            continue
        v = float(r[iv].replace(",", ""))
        ms = v / 1e6 if r[iu] in ("ns", "nsecond") else (v / 1e3 if r[iu] in ("us", "usecond") else v)
        name = r[ik].split("(")[0].replace("void ", "").replace("dpt::", "")
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += ms
    tot = sum(a[1] for a in agg.values())
    with open(os.path.join(dst, f"{rnd}_ncu_launch_list_vitl_b32_504.csv"), "w") as f:
        f.write("# ncu launch list of `python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --profile-steps 0` (ViT-L, B=32, 504^2, bf16), the two timed steps\n")
        f.write("# command: ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip <3 steps> -c <2 steps> --csv  (cold-cache, serialised: compare SHARES, not absolutes)\n")
        f.write(f"# {sum(a[0] for a in agg.values())} launches captured; total {tot:.2f} ms\n")
        f.write("kernel,launches,total_ms,share\n")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"\"{k}\",{a[0]},{a[1]:.3f},{a[1] / tot:.4f}\n")
    print("wrote launch list;", len(agg), "kernels,", f"{tot:.2f} ms")

# ---- full captures
WANT = [("duration_us", "gpu__time_duration.sum"), ("dram_read_MB", "dram__bytes_read.sum"), ("dram_write_MB", "dram__bytes_write.sum"),
        ("tensor_pipe_active_pct", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
        ("dram_throughput_pct", "dram__throughput.avg.pct_of_peak_sustained_elapsed"),
        ("sm_throughput_pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
        ("xu_pipe_pct", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
        ("fma_pipe_pct", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
        ("alu_pipe_pct", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active"),
        ("issue_active_pct", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
        ("sm_clock_ghz", "sm__cycles_elapsed.avg.per_second"),
        ("l2_hit_pct", "lts__t_sector_hit_rate.pct"), ("regs", "launch__registers_per_thread"), ("grid", "launch__grid_size")]
traffic = {}
for tag, title in [("gemm", "the four encoder GEMM launches of block 10 (qkv, proj, fc1, fc2)"), ("attn", "one attention launch"),
                   ("misc", "memory-bound kernels (row_stats, outnorm LayerNorm, bilinear resizes) and the head's halo convolution"),
                   ("attn_beit", "BEiT-L B=16 384^2 bf16: one bias-attention launch (bias tile staged by TMA)"),
                   ("swin", "SwinV2-L B=16 384^2 fp16: seven consecutive launches of the block loop (qkv GEMM with q/k "
                            "normalise epilogue, window attention d=32, proj, post-norm kernels, fc1, fc2)")]:
    p = os.path.join(src, f"ncu_{tag}_raw.csv")
    if not os.path.exists(p):
        continue
    hdr, rows = read_ncu_csv(p)
    units, rows = rows[0], rows[1:]
    ik = hdr.index("Kernel Name")
    cols = [(n, hdr.index(m)) for n, m in WANT if m in hdr]
    out = os.path.join(dst, f"{rnd}_ncu_full_{tag}_in_model.csv")
    with open(out, "w") as f:
        model_txt = "" if tag in ("attn_beit", "swin") else "ViT-L B=32 504^2 bf16, "
        f.write(f"# ncu --set full --clock-control none, {model_txt}{title} inside one forward (bench.py --steps 1 --warmup 3, DPT_GRAPH=0)\n")
        f.write("kernel," + ",".join(n for n, _ in cols) + "\n")
        for r in rows:
            vals = []
            for n, i in cols:
                v = float(r[i].replace(",", "")) if r[i] not in ("", "n/a") else float("nan")
                u = units[i]
                if n == "duration_us":
                    v = v / 1e3 if u in ("ns", "nsecond") else (v * 1e3 if u in ("ms", "msecond") else v)
                if n.endswith("_MB"):
                    v = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1.0) * v
                vals.append(f"{v:.1f}")
            f.write("\"" + r[ik].replace("void ", "").replace("dpt::", "") + "\"," + ",".join(vals) + "\n")
            if tag == "gemm":
                rd, wr = float(vals[1]), float(vals[2])
                traffic.setdefault("gemm256x2", []).append((rd + wr) * 1e6)
    print("wrote", os.path.basename(out), len(rows), "launches")
if traffic.get("gemm256x2"):
    t = traffic["gemm256x2"]
    json.dump({"gemm256x2": {"traffic_bytes_per_launch": sum(t) / len(t),
                             "source": f"profiles/{rnd}_ncu_full_gemm_in_model.csv: mean dram read+write of the four encoder GEMM shapes "
                                       f"(qkv, proj, fc1, fc2: {', '.join(f'{v / 1e6:.1f}' for v in t)} MB), ncu --set full, ViT-L B=32 504^2"}},
              open(os.path.join(dst, "traffic.json"), "w"), indent=1)
    print("wrote traffic.json")
