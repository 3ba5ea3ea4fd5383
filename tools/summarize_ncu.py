"""Turns the ncu artefacts that tools/profile.sh leaves in gpurun_out/ into the committed summaries under profiles/.
    python tools/summarize_ncu.py r01
"""
import collections
import csv
import re
import subprocess
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r01"

# ---- launch list
with open(f"gpurun_out/launches_{tag}.csv") as f:
    lines = [l for l in f if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0])
tot = 0.0
for row in csv.DictReader(lines):
    name = row.get("Kernel Name")
    if not name:
        continue
    v = float(row["Metric Value"].replace(",", ""))
    v = v / 1e3 if row["Metric Unit"] == "ns" else (v * 1e3 if row["Metric Unit"] == "ms" else v)
    short = re.sub(r"^void ", "", re.sub(r"\(.*", "", name)).replace("mb::", "")
    agg[short][0] += 1
    agg[short][1] += v
    tot += v
n_launch = sum(a[0] for a in agg.values())
out = [f"# ncu launch list — {tag}", "",
       f"`MB_NCU_RANGE=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none python bench.py "
       f"--steps 2 --warmup 3` (tools/profile.sh): the {n_launch} launches of the timed region (2 steps; "
       "per-launch times are serialised / cold-cache, so compare SHARES, not absolutes).", "",
       f"Total device time of the {n_launch} launches: {tot / 1e3:.2f} ms.", "",
       "| kernel | launches | total µs | share | avg µs |", "|---|---:|---:|---:|---:|"]
gemm_share = 0.0
for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
    out.append(f"| `{k}` | {n} | {t:.1f} | {100 * t / tot:.1f} % | {t / n:.2f} |")
    if k.startswith("gemm_bf16_kernel"):
        gemm_share += t / tot
out += ["", f"`gemm_bf16_kernel` (all template instances, <BN, EPI, CG>: EPI 0 bias, 1 GELU, 2 SwiGLU, 3 residual; CG 2 = "
        f"CTA pair): **{100 * gemm_share:.1f} %** of the device time.", ""]
open(f"profiles/{tag}_launches.md", "w").write("\n".join(out))

# ---- full capture of the four pixel-decoder GEMMs
raw = subprocess.run(["ncu", "-i", f"gpurun_out/prof_gemm_{tag}.ncu-rep", "--page", "raw", "--csv"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
keys = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum.per_second",
        "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__cluster_size", "sm__warps_active.avg.pct_of_peak_sustained_active"]
labels = ["pix.qkv  M=16384 N=3072 K=1024 bias", "pix.proj M=16384 N=1024 K=1024 bias+residual",
          "pix.fc1  M=16384 N=4096 K=1024 bias+GELU", "pix.fc2  M=16384 N=1024 K=4096 bias+residual"]
out = [f"# ncu --set full — the four pixel-decoder GEMMs of one step ({tag})", "",
       "`ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 147 -c 4 python bench.py --steps 1 "
       "--warmup 3` (tools/profile.sh).  Kernel: `mb::gemm_bf16_kernel<256, EPI, 2>` — 256x256 tile per CTA pair "
       "(tcgen05 cta_group::2), TMA-fed, staged TMA-store epilogue.", ""]
out.append("| metric | " + " | ".join(labels) + " |")
out.append("|---|" + "---:|" * len(labels))
for k in keys:
    if k not in idx:
        continue
    vals = [r[idx[k]] for r in rows[2:2 + len(labels)]]
    out.append(f"| `{k}` [{units[idx[k]]}] | " + " | ".join(vals) + " |")
out.append("")
open(f"profiles/{tag}_gemm_ncu.md", "w").write("\n".join(out))
print(open(f"profiles/{tag}_launches.md").read())
print(open(f"profiles/{tag}_gemm_ncu.md").read())
