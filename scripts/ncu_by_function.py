"""Development aid: aggregate an ncu report's executed instructions / stall samples by source function and line.
usage: python scripts/ncu_by_function.py gpurun_out/x.ncu-rep campx_b200/csrc/file.cu ENV_STEPS [top_lines]"""
import collections, csv, io, re, subprocess, sys
rep, cu, units = sys.argv[1], sys.argv[2], float(sys.argv[3])
top = int(sys.argv[4]) if len(sys.argv) > 4 else 0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
src = open(cu).read().split("\n")
starts = []
for i, l in enumerate(src, 1):
    m = re.match(r"^(?:template.*>\s*)?(__device__|__global__).*?(\w+)\(", l)
    if m:
        starts.append((i, m.group(2)))
hdr = None
fn, ln_agg, tot, tots = collections.defaultdict(lambda: [0, 0]), collections.defaultdict(lambda: [0, 0, ""]), 0, 0
cur_file_ok = False
for r in rows:
    if r and r[0] == "File Path":
        cur_file_ok = r[1].endswith(cu.split("/")[-1])
    if r and r[0] == "Line No":
        hdr = r
        iI, iS = hdr.index("Instructions Executed"), hdr.index("# Samples")
        continue
    if hdr is None or len(r) < len(hdr):
        continue
    try:
        ln, n, sm = int(r[0]), int(r[iI]), int(r[iS])
    except ValueError:
        continue
    name = "(other file)"
    if cur_file_ok:
        name = "other"
        for a, nm in starts:
            if ln >= a:
                name = nm
        ln_agg[ln][0] += n; ln_agg[ln][1] += sm; ln_agg[ln][2] = r[1]
    fn[name][0] += n; fn[name][1] += sm; tot += n; tots += sm
print("total warp-instructions per unit: %.1f" % (tot / units))
for k, (n, sm) in sorted(fn.items(), key=lambda kv: -kv[1][0])[:16]:
    print("%-26s %6.2f%% inst %6.2f%% samples (%.1f inst/unit)" % (k, 100 * n / tot, 100 * sm / max(1, tots), n / units))
for ln, (n, sm, s) in sorted(ln_agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%5d %6.2f%% inst %6.2f%% samp  %s" % (ln, 100 * n / tot, 100 * sm / max(1, tots), s.strip()[:100]))
