"""Turns the ncu artefacts that tools/profile_r02.sh leaves in gpurun_out/ into the committed summaries under profiles/.
    python tools/summarize_ncu_r02.py [r02]
"""
import collections
import csv
import gzip
import re
import shutil
import subprocess
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r02"


def launch_table(path, title, how):
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    agg = collections.OrderedDict()
    tot = 0.0
    for row in csv.DictReader(lines):
        name = row.get("Kernel Name")
        if not name:
            continue
        v = float(row["Metric Value"].replace(",", ""))
        v = v / 1e3 if row["Metric Unit"] == "ns" else (v * 1e3 if row["Metric Unit"] == "ms" else v)
        short = re.sub(r"^void ", "", re.sub(r"\(.*", "", name)).replace("mb::", "")[:80]
        a = agg.setdefault(short, [0, 0.0])
        a[0] += 1
        a[1] += v
        tot += v
    n = sum(a[0] for a in agg.values())
    out = [f"# {title}", "", how, "", f"{n} launches, {tot / 1e3:.2f} ms of device time in total.", "",
           "| kernel | launches | total µs | share | avg µs |", "|---|---:|---:|---:|---:|"]
    for k, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        out.append(f"| `{k}` | {c} | {t:.1f} | {100 * t / tot:.1f} % | {t / c:.2f} |")
    return "\n".join(out) + "\n", agg, tot


KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct",
        "sm__inst_executed_pipe_tensor.sum", "smsp__inst_executed.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__pcsamp_warps_issue_stalled_barrier",
        "smsp__pcsamp_warps_issue_stalled_long_scoreboard", "smsp__pcsamp_warps_issue_stalled_wait",
        "smsp__pcsamp_warps_issue_stalled_short_scoreboard", "smsp__pcsamp_warps_issue_stalled_selected"]


def full_table(rep, title, how, max_rows=8):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:2 + max_rows]
    idx = {h: i for i, h in enumerate(hdr)}
    names = [re.sub(r"\(.*", "", r[idx["Kernel Name"]]).replace("void ", "").replace("mb::", "")[:44] + f" #{i}"
             for i, r in enumerate(data)]
    out = [f"# {title}", "", how, "", "| metric | " + " | ".join(f"`{n}`" for n in names) + " |",
           "|---|" + "---:|" * len(names)]
    for k in KEYS:
        if k in idx:
            out.append(f"| `{k}` [{units[idx[k]]}] | " + " | ".join(r[idx[k]] for r in data) + " |")
    return "\n".join(out) + "\n"


md, agg, tot = launch_table(
    f"gpurun_out/launches_{tag}.csv", f"ncu launch list of bench.py's timed region — {tag}",
    "`MB_NCU_RANGE=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 4000 python "
    "bench.py --steps 1 --warmup 3 --no-cpu` (tools/profile_r02.sh): the first 4000 launches of the timed region — one edit "
    "round of 2 requests (encoder, 208-row prefill, then AR token steps; the kernels inside the CUDA-graph replays are listed "
    "individually).  Per-launch times are serialised / cold-cache: compare SHARES, not absolutes.")
open(f"profiles/{tag}_launches.md", "w").write(md)
with open(f"gpurun_out/launches_{tag}.csv", "rb") as fi, gzip.open(f"profiles/{tag}_launches.csv.gz", "wb") as fo:
    shutil.copyfileobj(fi, fo)
md2, _, _ = launch_table(
    "gpurun_out/r02_ar_step_launches.csv", "ncu launch list of one edit round on the eager path (4 LLM layers, 2 token steps)",
    "`ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none python tools/profile_ar_step.py` — "
    "B = 3 CFG rows, AR_LAYERS = 4 of the 28 true-width MoE layers, full RF head / MingTok; the region holds the encoder, the "
    "104-token prefill, TWO token steps and the pixel decoder.")
open(f"profiles/{tag}_ar_step_launches.md", "w").write(md2)
open(f"profiles/{tag}_rf_fused_ncu.md", "w").write(full_table(
    f"gpurun_out/prof_rf_fused_{tag}.ncu-rep", "ncu --set full — the persistent RF sampler kernel (6 rows = 2 requests x 3 CFG rows)",
    "`RF_ROWS=6 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:rf_sample_fused python "
    "tools/profile_rf.py`: ONE launch = 16 Euler steps x 12 residual blocks of the default-size head (28.99 GB of bf16 weights)."
    "  The 3-row launch of the first version of the kernel: `r02_rf_fused_3rows_ncu.md`.", 1))
try:
    open(f"profiles/{tag}_rf_fused_3rows_ncu.md", "w").write(full_table(
        "gpurun_out/r02_rf_fused.ncu-rep", "ncu --set full — the persistent RF sampler kernel, first version (3 rows, 5 ring stages)",
        "`ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:rf_sample_fused python "
        "tools/profile_rf.py` at the commit that introduced the kernel (before the parameter preloading / lean inner loop: "
        "6.89 ms under ncu; the source of profiles/rf_fused_traffic.json).", 1))
except Exception as e:  # noqa: BLE001
    print("3-row capture missing:", e)
open(f"profiles/{tag}_ar_moe_ncu.md", "w").write(full_table(
    f"gpurun_out/prof_ar_moe_{tag}.ncu-rep", "ncu --set full — routed-expert streaming kernels and GQA decode attention of the AR step",
    "`AR_LAYERS=4 ncu --profile-from-start off --set full -k regex:\"moe_expert_kernel|attn_decode_gqa128\" -s 8 -c 6 python "
    "tools/profile_ar_step.py`: decode-step launches (B = 3 rows; `moe_expert_kernel<0, 1>` = gate/up + SwiGLU, `<1, 1>` = down).", 6))
open(f"profiles/{tag}_ar_gemv_ncu.md", "w").write(full_table(
    f"gpurun_out/prof_ar_gemv_{tag}.ncu-rep", "ncu --set full — the per-layer weight-streaming kernel (gemv) inside the AR step",
    "`AR_LAYERS=4 ncu --profile-from-start off --set full -k regex:gemv_bf16_kernel -s 10 -c 8 python tools/profile_ar_step.py`: "
    "LLM decode-step launches (qkv 3072x2048, dense 2048x2048 + residual, router gates 64x2048, shared experts 5632x2048 SwiGLU "
    "and 2048x2816).", 8))
print(md[:3000])
