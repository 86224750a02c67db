"""profiles/ncu_summary_r02.{md,json} + profiles/launches_r02.md from the CSV exports of one GPU visit (tools/gpu_r2_q.sh:
`ncu -i X.ncu-rep --page raw --csv > X.raw.csv` on the box, because gpurun copies back at most 64 MiB).
usage: summarize_profiles_r02.py [prefix=r2q_] [tag=r02]"""
import collections
import csv
import glob
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
prefix = sys.argv[1] if len(sys.argv) > 1 else "r2q_"
tag = sys.argv[2] if len(sys.argv) > 2 else "r02"
dst = os.path.join(ROOT, "profiles")

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
        "smsp__average_warp_latency_per_inst_issued.ratio",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum"]
UNIT = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1, "Tbyte": 1e12}

lines = [f"# ncu summaries, {tag}", "",
         "Captured on a B200 under `gpurun` (`tools/gpu_r2_q.sh`): `ncu --set full --clock-control none --import-source on`, step kernels = launch 31 of",
         "`tools/profile_step.py ENVS SUBSTEPS 34 [--norm-obs | --physics gnd_drag | --reward-id R]` (steady-state reset mix, saturating actions); the PPO",
         "kernels = the second minibatch of `tools/profile_ppo_fused.py 65536 32768 bf16x3`.  Times under ncu are cold-cache and serialised; bench.py's",
         "CUDA-event numbers are the ones to quote.  `traffic` = dram__bytes_read.sum + dram__bytes_write.sum.", ""]
summary = {}
for path in sorted(glob.glob(os.path.join(OUT, prefix + "*.raw.csv"))):
    name = os.path.basename(path)[len(prefix):-len(".raw.csv")]
    rows = list(csv.reader(open(path)))
    if len(rows) < 3:
        continue
    hdr, units = rows[0], rows[1]
    kn = hdr.index("Kernel Name")
    for j, vals in enumerate(rows[2:]):
        if len(vals) != len(hdr):
            continue
        key = name if len(rows) == 3 else f"{name}#{j}"
        kname = vals[kn].replace("CUtensorMap_st, CUtensorMap_st, CUtensorMap_st, CUtensorMap_st, ", "4 x CUtensorMap, ")
        rec = {"kernel": kname}
        for k in KEYS:
            if k in hdr:
                rec[k] = vals[hdr.index(k)]
                rec[k + "|unit"] = units[hdr.index(k)]
        try:
            rd = float(rec["dram__bytes_read.sum"]) * UNIT[rec["dram__bytes_read.sum|unit"]]
            wr = float(rec["dram__bytes_write.sum"]) * UNIT[rec["dram__bytes_write.sum|unit"]]
            rec["traffic_bytes"] = rd + wr
        except Exception:  # noqa: BLE001
            pass
        summary[key] = rec
# one table per family
def table(title, keys, cols):
    out = [f"## {title}", "", "| capture | kernel | " + " | ".join(c[1] for c in cols) + " |", "|---|---|" + "---|" * len(cols)]
    for k in keys:
        r = summary[k]
        cells = []
        for metric, _ in cols:
            v = r.get(metric, "")
            if metric == "traffic_bytes" and v != "":
                v = f"{v / 1e6:.2f} MB"
            elif v != "" and (r.get(metric + "|unit") or "") not in ("", "%", "cycle", "inst", "warp", "block", "register/thread", "thread"):
                v = f"{v} {r.get(metric + '|unit')}"
            cells.append(str(v))
        out.append(f"| {k} | `{r['kernel'][:60]}` | " + " | ".join(cells) + " |")
    return out + [""]
step_keys = [k for k in summary if "ppo" not in k]
ppo_keys = [k for k in summary if "ppo" in k]
lines += table("Environment step kernel variants", step_keys, [
    ("gpu__time_duration.sum", "time"), ("traffic_bytes", "DRAM traffic"), ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"), ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe %"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %"), ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "FP64 pipe %"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "threads / inst"), ("launch__registers_per_thread", "regs"),
    ("smsp__sass_inst_executed_op_local_ld.sum", "local ld"), ("smsp__sass_inst_executed_op_local_st.sum", "local st"),
    ("smsp__inst_executed.sum", "warp inst"), ("launch__grid_size", "grid")])
lines += table("PPO update kernels (second minibatch, bf16x3)", ppo_keys, [
    ("gpu__time_duration.sum", "time"), ("sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor pipe active %"),
    ("l1tex__m_xbar2l1tex_read_bytes.sum", "L2 -> SM bytes"), ("traffic_bytes", "DRAM traffic"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
    ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid")])
# warp-stall totals of the source page export, if present
for path in sorted(glob.glob(os.path.join(OUT, prefix + "*.source.csv"))):
    srows = list(csv.reader(open(path)))
    if len(srows) > 2:
        h = srows[1]
        tot = collections.Counter()
        for r in srows[2:]:
            if len(r) != len(h):
                continue
            for i, c in enumerate(h):
                if c.startswith("stall_") and "Not Issued" not in c:
                    try:
                        tot[c] += float(r[i])
                    except ValueError:
                        pass
        allv = sum(tot.values()) or 1
        lines += [f"warp-stall samples, {os.path.basename(path)[len(prefix):-11]}: " + ", ".join(f"{c[6:]} {100 * v / allv:.1f}%" for c, v in tot.most_common(8)), ""]
open(os.path.join(dst, f"ncu_summary_{tag}.md"), "w").write("\n".join(lines) + "\n")
json.dump({k: {m: v for m, v in r.items() if not m.endswith("|unit")} for k, r in summary.items()}, open(os.path.join(dst, f"ncu_summary_{tag}.json"), "w"), indent=1)

for src_name, title, cmd in ((prefix + "launches.csv", "bench.py headline", "python bench.py --steps 64 --warmup 3 --no-cpu --no-vecenv --no-ppo --no-configs --rotating-handles 8 --sweep 4194304"),
                             (prefix + "launches_ppo.csv", "fused PPO update", "python tools/profile_ppo_fused.py 65536 32768 bf16x3 (kernels of the library only)")):
    lp = os.path.join(OUT, src_name)
    if not os.path.exists(lp):
        continue
    rows = [r for r in csv.reader(open(lp)) if len(r) > 5]
    hi = next((i for i, r in enumerate(rows) if "Kernel Name" in r), None)
    if hi is None:
        continue
    h = rows[hi]
    kn, mv = h.index("Kernel Name"), h.index("Metric Value")
    agg, cnt = collections.Counter(), collections.Counter()
    for r in rows[hi + 1:]:
        try:
            agg[r[kn]] += float(r[mv].replace(",", "")); cnt[r[kn]] += 1
        except ValueError:
            pass
    tot = sum(agg.values()) or 1
    mode = "a" if src_name.endswith("_ppo.csv") else "w"
    with open(os.path.join(dst, f"launches_{tag}.md"), mode) as f:
        f.write(f"# ncu launch list, {tag}: {title}\n\n`ncu --metrics gpu__time_duration.sum --clock-control none -c N --csv {cmd}` (cold-cache, serialised: compare "
                "shares, not absolute times).\n\n| kernel | launches | total ns | share |\n|---|---|---|---|\n")
        for k, v in agg.most_common():
            f.write(f"| `{k.replace('CUtensorMap_st, CUtensorMap_st, CUtensorMap_st, CUtensorMap_st, ', '')[:110]}` | {cnt[k]} | {v:.0f} | {100 * v / tot:.1f}% |\n")
        f.write("\n")
print("wrote", f"ncu_summary_{tag}.md", len(summary), "kernels")
