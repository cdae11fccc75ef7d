"""Turn gpurun_out/<tag>_* (ncu launch list, full captures per batch size, sanitizer logs) into the tracked summaries
under profiles/:  <tag>_ncu_summary.md, <tag>_launches.md, traffic.json, <tag>_sanitizers.md, <tag>_sass_opcodes.md.

    python scripts/summarize_profiles.py r02
"""
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
TAG = sys.argv[1] if len(sys.argv) > 1 else "r02"

METRICS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__grid_size", "launch__block_size",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum.per_second",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_issued.avg.per_cycle_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
]


def raw_page(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return rows[0], rows[1], rows[2:]


def short(name):
    return re.sub(r"^void ", "", name).split("(")[0].replace("roko::", "")


def ncu_summaries():
    traffic = {}
    lines = [f"# ncu summaries ({TAG})", "",
             "`ncu --set full --clock-control none --import-source on` over `scripts/profile_target.py <batch> 3` (the first",
             "forward is skipped); one column per captured launch.  Times under ncu are cold-cache and serialised, clocks are",
             "whatever the box ran (`sm__cycles_elapsed / gpu__time_duration`).", ""]
    for f in sorted(os.listdir(SRC)):
        m = re.match(TAG + r"_b(\d+)\.ncu-rep", f)
        if not m:
            continue
        batch = int(m.group(1))
        hdr, units, rows = raw_page(os.path.join(SRC, f))
        if not rows:
            continue
        ki = hdr.index("Kernel Name")
        seen, keep = set(), []
        for r in rows:                                   # one launch per distinct kernel (the first captured)
            k = short(r[ki])
            if k not in seen:
                seen.add(k)
                keep.append(r)
        lines += [f"## batch {batch}", "", "| metric | unit | " + " | ".join(short(r[ki]) for r in keep) + " |", "|---|---|" + "---|" * len(keep)]
        for met in METRICS:
            if met in hdr:
                i = hdr.index(met)
                lines.append(f"| {met} | {units[i]} | " + " | ".join(r[i][:14] for r in keep) + " |")
        for r in keep:
            stalls = []
            for i, h in enumerate(hdr):
                if "issue_stalled" in h and h.endswith("per_issue_active.ratio"):
                    try:
                        stalls.append((float(r[i]), h.split("issue_stalled_")[1].split("_per_issue")[0]))
                    except ValueError:
                        pass
            lines.append(f"| top stalls: {short(r[ki])} | warps / issue | " + ", ".join(f"{n} {v:.2f}" for v, n in sorted(stalls, reverse=True)[:4]) + " |" + " |" * (len(keep) - 1))
        lines.append("")

        def val(r, name):
            i = hdr.index(name)
            v = float(r[i].replace(",", ""))
            return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(units[i].lower(), 1)
        per = collections.defaultdict(list)
        for r in rows:
            per[short(r[ki])].append(val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum"))
        for k, v in per.items():
            traffic[f"{k}@B{batch}"] = sum(v) / len(v)
        if batch == 2368:                                # DRAM bytes of one whole pass: front + 3 x (proj, rec) + head
            need = {"front_tc_kernel": 1, "proj_h_kernel<512>": 1, "proj_h_kernel<256>": 2, "rec_h_kernel": 3, "head_kernel": 1}
            if all(k in per for k in need):
                traffic["pass@B2368"] = sum(n * sum(per[k]) / len(per[k]) for k, n in need.items())
    with open(os.path.join(OUT, f"{TAG}_ncu_summary.md"), "w") as f:
        f.write("\n".join(lines) + "\n")
    with open(os.path.join(OUT, "traffic.json"), "w") as f:
        json.dump(traffic, f, indent=1, sort_keys=True)


def launch_list():
    lp = os.path.join(SRC, f"{TAG}_launches.csv")
    if not os.path.exists(lp):
        return
    tot, cnt = collections.defaultdict(float), collections.Counter()
    with open(lp) as f:
        rd = csv.DictReader(l for l in f if not l.startswith("=="))
        for row in rd:
            if row.get("Metric Name") != "gpu__time_duration.sum":
                continue
            name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "")
            v = float(row["Metric Value"].replace(",", "")) * {"ns": 1e-3, "us": 1, "ms": 1e3}.get(row["Metric Unit"], 1e-3)
            tot[name] += v
            cnt[name] += 1
    ours = {k: v for k, v in tot.items() if k.startswith("roko::") and "ffma_peak" not in k}
    T = sum(ours.values())
    with open(os.path.join(OUT, f"{TAG}_launches.md"), "w") as f:
        f.write(f"# ncu launch list ({TAG}): `bench.py --steps 6 --warmup 3 --min-region 0.02` under `ncu --metrics gpu__time_duration.sum`\n\n")
        f.write("Shares are over this repo's kernels inside the bench run (parity gate, warm-up, timed 128-window steps, e2e and\n"
                "coalesced passes); absolute times are cold-cache and serialised by ncu.  No cuBLAS / cuDNN kernel appears.\n\n"
                "| kernel | launches | total us | mean us | share |\n|---|---|---|---|---|\n")
        for k, v in sorted(ours.items(), key=lambda kv: -kv[1]):
            f.write(f"| {k} | {cnt[k]} | {v:.1f} | {v / cnt[k]:.1f} | {100 * v / T:.1f} % |\n")
        f.write("\nOther kernels in the process (torch RNG fill, cat, memcpy, FP32 peak probe): "
                + ", ".join(f"{k.split('<')[0][:40]} x{cnt[k]}" for k in tot if k not in ours)[:600] + "\n")


def sanitizers():
    logs = sorted(f for f in os.listdir(SRC) if re.match(TAG + r"_(memcheck|racecheck)_", f))
    if not logs:
        return
    with open(os.path.join(OUT, f"{TAG}_sanitizers.md"), "w") as f:
        f.write(f"# compute-sanitizer ({TAG})\n\n`scripts/capture_profiles.sh`: `scripts/profile_target.py <batch> 1` under memcheck / racecheck, CUDA graphs off,\n"
                "`ROKO_B200_REC_TC_MIN=32` for the 33 / 40-window runs (a ragged last 32-window group in `rec_h_kernel`; its clamped\n"
                "`gi` reads stay inside the batch), default threshold for the 5-window run (register-resident FFMA recurrence).\n"
                "`*_train_*`: three training steps (`scripts/train_profile.py <batch> 1`: forward with dropout, loss, backward, Adam).\n\n")
        for name in logs:
            body = open(os.path.join(SRC, name)).read().strip().splitlines()
            f.write(f"## {name}\n\n```\n" + "\n".join(body[-6:]) + "\n```\n\n")


def train_launches():
    """One training step's kernels (last step of scripts/train_profile.py under the ncu launch-list pass)."""
    src = os.path.join(SRC, f"{TAG}_train_launches.csv")
    if not os.path.exists(src):
        return
    rows = [r for r in csv.reader(open(src)) if len(r) > 5]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    launches = [(r[ki], float(r[vi].replace(",", "")) * (1e-3 if r[ui] == "ns" else 1.0)) for r in rows[1:]]
    start = max(i for i, (k, _) in enumerate(launches) if "embed_drop" in k)
    agg = collections.OrderedDict()
    for k, v in launches[start:]:
        k = re.sub(r"\(.*", "", k.replace("roko::", "").replace("void ", ""))
        k = "torch:" + k.split("<")[0][:40] if k.startswith("at::") else k
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v
    total = sum(a[1] for a in agg.values())
    events = ""
    log = os.path.join(SRC, f"{TAG}_train_profile.log")
    if os.path.exists(log):
        events = open(log).read().strip().splitlines()[-1]
    with open(os.path.join(OUT, f"{TAG}_train_launches.md"), "w") as f:
        f.write(f"# ncu launch list of one training step ({TAG}): `scripts/train_profile.py 128 2` under `ncu --metrics gpu__time_duration.sum`\n\n"
                "Last step of the run (from its `embed_drop_kernel` on): train-mode forward with dropout, cross-entropy, hand-written backward,\n"
                "fused Adam, batch 128, default chain (`ROKO_B200_TRAIN_TC=6`).\n"
                f"CUDA-event times of the same script without ncu (20 steps): {events}\n\n```\n")
        for k, a in sorted(agg.items(), key=lambda t: -t[1][1]):
            f.write(f"{a[1]:9.1f} us {100 * a[1] / total:5.1f}% {a[0]:3d}x  {k}\n")
        f.write(f"total {total:.1f} us over {sum(a[0] for a in agg.values())} launches\n```\n")


def sass():
    so = os.path.join(ROOT, "roko_b200", "libroko_b200.so")
    if not os.path.exists(so):
        return
    text = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
    per, cur = collections.defaultdict(collections.Counter), None
    for line in text.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0].replace("roko::", "")
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            op = m.group(1)
            for key in ("UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UBLKCP", "UTMALDG", "UTCBAR", "SYNCS", "HMMA", "MUFU", "FFMA", "F2FP", "LDGSTS"):
                if op.startswith(key):
                    per[cur][key] += 1
    keys = ["UTCHMMA", "LDTM", "STTM", "UBLKCP", "UTCBAR", "SYNCS", "HMMA", "MUFU", "F2FP", "FFMA", "LDGSTS"]
    with open(os.path.join(OUT, f"{TAG}_sass_opcodes.md"), "w") as f:
        f.write(f"# SASS opcode summary ({TAG}): `cuobjdump -sass roko_b200/libroko_b200.so`, static instruction counts per kernel\n\n"
                "`UTCHMMA` = tcgen05.mma, `LDTM` / `STTM` = tcgen05.ld / st, `UBLKCP` = cp.async.bulk (TMA bulk copy), `UTCBAR` = tcgen05.commit,\n"
                "`SYNCS` = mbarrier ops, `HMMA` = legacy mma.sync (round-1 kernels kept for A/B and the training GEMM), `F2FP` = packed fp32->fp16 converts.\n\n"
                "| kernel | " + " | ".join(keys) + " |\n|---|" + "---|" * len(keys) + "\n")
        for k in sorted(per):
            if any(per[k][x] for x in ("UTCHMMA", "LDTM", "UBLKCP", "HMMA")) or "kernel" in k:
                f.write(f"| {k[:70]} | " + " | ".join(str(per[k][x]) for x in keys) + " |\n")


def main():
    os.makedirs(OUT, exist_ok=True)
    ncu_summaries()
    launch_list()
    sanitizers()
    train_launches()
    sass()
    print("wrote", sorted(os.listdir(OUT)))


if __name__ == "__main__":
    main()
