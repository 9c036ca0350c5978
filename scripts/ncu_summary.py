"""Turn gpurun_out/*.ncu-rep + launch list into the committed summaries under profiles/ (run here, no GPU)."""
import collections, csv, io, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
kname = sys.argv[2] if len(sys.argv) > 2 else "agent_rollout"     # agent_rollout | agent_obs | generic_rollout
title = sys.argv[3] if len(sys.argv) > 3 else "kernel k_agent_rollout (bench.py workload)"
rep = os.path.join(ROOT, "gpurun_out", "%s_%s.ncu-rep" % (tag, kname))
out_dir = os.path.join(ROOT, "profiles")
os.makedirs(out_dir, exist_ok=True)

raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum", "lts__t_bytes.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__shared_mem_per_block_dynamic",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct"]
lines = ["# ncu --set full, %s, %s" % (title, tag), ""]
traffic = {}
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")]
    lines.append("## %s" % name)
    vals = {}
    for w in WANT:
        if w in hdr:
            vals[w] = (r[hdr.index(w)], units[hdr.index(w)])
            lines.append("- %s: %s %s" % (w, *vals[w]))
    def num(k):
        v, u = vals[k]
        v = float(v.replace(",", ""))
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(u, 1.0)
        return v * scale
    rd, wr = num("dram__bytes_read.sum"), num("dram__bytes_write.sum")
    traffic = {"dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes": rd + wr}
    lines.append("- **dram traffic per launch: %.1f MB (read %.1f + write %.1f)**" % ((rd + wr) / 1e6, rd / 1e6, wr / 1e6))
    lines.append("")

sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
srows = list(csv.reader(io.StringIO(sass)))
start = [i for i, r in enumerate(srows) if r and r[0] == "Address"][0]
sh = srows[start]
body = [r for r in srows[start + 1:] if len(r) == len(sh)]
ix = {h: i for i, h in enumerate(sh)}
tot = sum(float(r[ix["Instructions Executed"]]) for r in body)
ops = collections.Counter()
for r in body:
    toks = [t for t in r[ix["Source"]].split() if not t.startswith("@")]
    ops[toks[0].split(".")[0]] += float(r[ix["Instructions Executed"]])
lines.append("## SASS opcode mix (warp instructions executed: %d)" % tot)
lines.append(", ".join("%s %.1f%%" % (o, 100 * n / tot) for o, n in ops.most_common(16)))
stalls = [h for h in sh if h.startswith("stall_") and "Not Issued" not in h]
agg = {s: sum(float(r[ix[s]]) for r in body) for s in stalls}
ts = sum(agg.values()) or 1
lines.append("")
lines.append("## warp stall sampling (all samples)")
lines.append(", ".join("%s %.1f%%" % (k[6:], 100 * v / ts) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
proof = [o for o in ops if o in ("STG", "LDS", "LDG", "UTMALDG", "UTMASTG", "UBLKCP", "HMMA")]
lines.append("")
lines.append("memory/tensor opcodes present: %s (no tensor-core opcodes: nothing on this path is a contraction)" % ", ".join(sorted(proof)))
open(os.path.join(out_dir, "%s_ncu_%s.md" % (tag, kname)), "w").write("\n".join(lines) + "\n")
if kname != "agent_rollout":
    print("\n".join(lines))
    sys.exit(0)

# launch list (gpu__time_duration per launch of the bench command)
ll = os.path.join(ROOT, "gpurun_out", "%s_launches.csv" % tag)
if os.path.exists(ll):
    rows = [r for r in csv.reader(open(ll)) if len(r) > 10 and r[0].isdigit()]
    seq = [(r[4].split("(")[0][-60:], float(r[-1].replace(",", ""))) for r in rows]
    runs = []
    for k, t in seq:                      # run-length encoded launch sequence of the whole command
        if runs and runs[-1][0] == k:
            runs[-1][1] += 1
            runs[-1][2] += t
        else:
            runs.append([k, 1, t])
    total = sum(t for _, t in seq)
    with open(os.path.join(out_dir, "%s_launch_list.md" % tag), "w") as f:
        f.write("# ncu launch list of `%s` (%s)\n\n" % (os.environ.get("NCU_BENCH_CMD", "python bench.py --steps 512 --warmup 64 --e2e-steps 8"), tag))
        f.write("`ncu --metrics gpu__time_duration.sum --clock-control none -c 400` over the whole command, in launch order "
                "(run-length encoded) -- cold-cache, serialised: compare shares, not absolutes.\n\n")
        f.write("| # | kernel | consecutive launches | total us | share of command | mean us |\n|---|---|---|---|---|---|\n")
        for i, (k, n, t) in enumerate(runs):
            f.write("| %d | %s | %d | %.1f | %.1f%% | %.1f |\n" % (i, k, n, t / 1e3, 100 * t / total, t / n / 1e3))
        f.write("\n" + os.environ.get("NCU_BENCH_NOTE", "Sections of bench.py: set-up (reset / render / `k_fill_actions`), then the warm-up + TIMED region = the one "
                "long run of `k_agent_rollout<1, 1, 0>` (64 + 512 steps at 32 steps per launch: nothing else launches there, so "
                "the kernel's share of a timed step is 100 %), then the e2e leg (`cx_step` = `k_agent_rollout` with T = 1, one "
                "launch per play()), then the secondary board + layered-board measurement (`k_agent_rollout_obs`).") + "\n")
    import shutil
    shutil.copy(ll, os.path.join(out_dir, "%s_launches.csv" % tag))
# (profiles/roofline_traffic.json is maintained by hand since round 2: one key per launch geometry bench.py can time)
print(open(os.path.join(out_dir, "%s_ncu_agent_rollout.md" % tag)).read())
