"""Turn gpurun_out/final_* (ncu launch list + full captures) into the tracked summaries under profiles/."""
import collections
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
SRC = os.path.join(ROOT, "gpurun_out")
TAG = sys.argv[1] if len(sys.argv) > 1 else "r01"

METRICS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__grid_size", "launch__block_size",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum.per_second",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_issued.avg.per_cycle_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__ops_path_tensor_op_utchmma_src_tf32_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
]


def raw_page(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return rows[0], rows[1], rows[2:]


def main():
    os.makedirs(OUT, exist_ok=True)
    traffic = {}
    lines = [f"# ncu summaries ({TAG})", "",
             "One `ncu --set full --clock-control none` capture per hot kernel (2 launches each) of",
             "`scripts/profile_target.py <batch> 3`; per-launch values.  Times under ncu are cold-cache and serialised.", ""]
    for f in sorted(os.listdir(SRC)):
        m = re.match(r"final_(\w+)_b(\d+)\.ncu-rep", f)
        if not m:
            continue
        hdr, units, rows = raw_page(os.path.join(SRC, f))
        if not rows:
            continue
        kname, batch = m.group(1), int(m.group(2))
        lines += [f"## {kname} @ batch {batch}", "", "| metric | unit | " + " | ".join(f"launch {i}" for i in range(len(rows))) + " |",
                  "|---|---|" + "---|" * len(rows)]
        ki = hdr.index("Kernel Name")
        lines.append("| kernel | | " + " | ".join(r[ki][:60] for r in rows) + " |")
        for met in METRICS:
            if met in hdr:
                i = hdr.index(met)
                lines.append(f"| {met} | {units[i]} | " + " | ".join(r[i][:14] for r in rows) + " |")
        stalls = []
        for i, h in enumerate(hdr):
            if "issue_stalled" in h and h.endswith("per_issue_active.ratio"):
                try:
                    stalls.append((float(rows[0][i]), h.split("issue_stalled_")[1].split("_per_issue")[0]))
                except ValueError:
                    pass
        lines.append("| top stalls (warps per issue) | | " + ", ".join(f"{n} {v:.2f}" for v, n in sorted(stalls, reverse=True)[:5]) + " |")
        lines.append("")
        # dram traffic per launch for bench.py's roofline.traffic
        def val(r, name):
            i = hdr.index(name)
            v = float(r[i].replace(",", ""))
            u = units[i].lower()
            return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
        for r in rows:
            key = re.sub(r"^void ", "", r[ki]).split("(")[0]
            key = key.replace("roko::", "")
            traffic[f"{key}@B{batch}"] = val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum")
    with open(os.path.join(OUT, f"{TAG}_ncu_summary.md"), "w") as f:
        f.write("\n".join(lines) + "\n")
    with open(os.path.join(OUT, "traffic.json"), "w") as f:
        json.dump(traffic, f, indent=1, sort_keys=True)

    lp = os.path.join(SRC, "final_launches.csv")
    if os.path.exists(lp):
        tot, cnt = collections.defaultdict(float), collections.Counter()
        with open(lp) as f:
            rd = csv.DictReader(l for l in f if not l.startswith("=="))
            for row in rd:
                if row.get("Metric Name") != "gpu__time_duration.sum":
                    continue
                name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "")
                v = float(row["Metric Value"].replace(",", ""))
                v *= {"ns": 1e-3, "us": 1, "ms": 1e3}.get(row["Metric Unit"], 1e-3)
                tot[name] += v
                cnt[name] += 1
        ours = {k: v for k, v in tot.items() if k.startswith("roko::") and "ffma_peak" not in k}
        T = sum(ours.values())
        with open(os.path.join(OUT, f"{TAG}_launches.md"), "w") as f:
            f.write(f"# ncu launch list ({TAG}): `bench.py --steps 6 --warmup 3` under `ncu --metrics gpu__time_duration.sum`\n\n")
            f.write("Shares are over this repo's kernels inside the bench run (parity gate, warm-up, timed steps, e2e and\n"
                    "coalesced passes); absolute times are cold-cache and serialised by ncu.\n\n| kernel | launches | total us | mean us | share |\n|---|---|---|---|---|\n")
            for k, v in sorted(ours.items(), key=lambda kv: -kv[1]):
                f.write(f"| {k} | {cnt[k]} | {v:.1f} | {v / cnt[k]:.1f} | {100 * v / T:.1f} % |\n")
            f.write("\nOther kernels in the process (torch RNG fill, cat, memcpy, FP32 peak probe): "
                    + ", ".join(f"{k.split('<')[0][:40]} x{cnt[k]}" for k in tot if k not in ours)[:600] + "\n")
    print("wrote", os.listdir(OUT))


if __name__ == "__main__":
    main()
