"""Top warp-stall sites of a kernel from an `ncu --set full --import-source on` report (PC sampling per SASS instruction):
    python tools/ncu_hot_sites.py gpurun_out/prof_rf_fused_r02.ncu-rep [N] > profiles/<name>.md
Runs on CPU (reads the report with `ncu -i ... --page source --csv --print-source sass`)."""
import csv
import io
import subprocess
import sys


def main() -> int:
    rep = sys.argv[1]
    top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True,
                         text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    kernel = rows[0][1]
    hdr, data = rows[1], [r for r in rows[2:] if len(r) == len(rows[1])]
    ix = {h: i for i, h in enumerate(hdr)}
    num = lambda r, h: int(r[ix[h]]) if r[ix[h]].isdigit() else 0  # noqa: E731
    total = sum(num(r, "# Samples") for r in data)
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    agg = sorted(((h, sum(num(r, h) for r in data)) for h in stalls), key=lambda x: -x[1])
    print(f"# Warp-stall sites of `{kernel}`\n")
    print(f"From `{rep}` (PC sampling, {total} samples over {len(data)} SASS instructions; `tools/ncu_hot_sites.py`).\n")
    print("Stall reasons over the whole kernel: " + ", ".join(f"{h[6:]} {100 * v / total:.1f} %" for h, v in agg[:8]) + ".\n")
    print("| share of samples | SASS instruction | the instruction in front of it | main stall reasons |")
    print("|---:|---|---|---|")
    order = sorted(range(len(data)), key=lambda i: -num(data[i], "# Samples"))[:top_n]
    for i in order:
        r = data[i]
        s = num(r, "# Samples")
        why = sorted(((h[6:], num(r, h)) for h in stalls if num(r, h) > 0), key=lambda x: -x[1])[:2]
        prev = data[i - 1][ix["Source"]].strip() if i else ""
        print(f"| {100 * s / total:.1f} % | `{r[ix['Source']].strip()}` | `{prev}` | "
              + ", ".join(f"{k} {100 * v / max(s, 1):.0f} %" for k, v in why) + " |")
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
