#!/usr/bin/env python
"""Turn an .ncu-rep (ncu --set full) into a short text summary for profiles/: key metrics per kernel
and the hottest source lines (needs nvdisasm line info from the built library).
usage: tools_ncu_summary.py <report.ncu-rep> <out.md> [symbols_per_launch]"""
import collections, csv, io, os, re, subprocess, sys, tempfile

rep, out = sys.argv[1:3]
nsym = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0
ROOT = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(ROOT, "image_compression_2_b200", "liblatentcodec.so")
KEYS = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg", "smsp__inst_executed.sum", "smsp__inst_issued.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "sm__inst_executed_pipe_fp64.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]


def run(cmd):
    return subprocess.run(cmd, capture_output=True, text=True).stdout


raw = list(csv.reader(io.StringIO(run(["ncu", "-i", rep, "--page", "raw", "--csv"]))))
hdr, units = raw[0], raw[1]
kname_i = hdr.index("Kernel Name")
lines = ["# ncu summary of `%s`" % os.path.basename(rep), "",
         "Captured with `ncu --set full --clock-control none --import-source on` on a B200 (bench.py workload, config 2).",
         "Per-launch values; cold-cache and serialised, so compare shares, not absolutes.", ""]
for row in raw[2:]:
    if len(row) != len(hdr):
        continue
    lines.append("## %s" % row[kname_i].split("(")[0])
    lines.append("")
    lines.append("| metric | value | unit |")
    lines.append("|---|---|---|")
    vals = {}
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            vals[k] = row[i]
            lines.append("| %s | %s | %s |" % (k, row[i], units[i]))
    try:
        inst = float(vals["smsp__inst_executed.sum"].replace(",", ""))
        cyc = float(vals["sm__cycles_elapsed.avg"].replace(",", ""))
        if nsym:
            lines.append("| warp instructions per symbol | %.1f | inst |" % (inst / nsym))
        lines.append("| issue-slot utilisation (inst / (cycles x 592 sub-partitions)) | %.3f | fraction |" % (inst / (cyc * 592)))
    except Exception:
        pass
    lines.append("")

# hottest source lines per kernel
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", LIB], cwd=tmp, capture_output=True)
cubins = [f for f in os.listdir(tmp) if f.endswith(".cubin")]
dis = run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubins[0])]).splitlines() if cubins else []
func = None; cur = None; table = collections.defaultdict(list)
for ln in dis:
    m = re.match(r'\s*\.section\s+\.text\.(\S+?),', ln)
    if m: func = m.group(1); cur = None; continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);', ln)
    if m and func: table[func].append(cur)
src = run(["ncu", "-i", rep, "--page", "source", "--csv"]).splitlines()
sections = [i for i, l in enumerate(src) if l.startswith('"Kernel Name"')] + [len(src)]
seen = set()
for a, b in zip(sections[:-1], sections[1:]):
    rows = list(csv.reader(src[a:b]))
    kname = rows[0][1].split("(")[0]
    if kname in seen:
        continue
    seen.add(kname)
    h = rows[1]
    if "# Samples" not in h:
        continue
    si, ii = h.index("# Samples"), h.index("Instructions Executed")
    data = [r for r in rows[2:] if len(r) == len(h)]
    fn = [f for f in table if kname in f]
    if not fn or len(table[fn[0]]) != len(data):
        continue
    agg = collections.defaultdict(lambda: [0.0, 0.0]); ts = ti = 0.0
    for k, r in enumerate(data):
        s = float(r[si] or 0); i = float(r[ii] or 0); ts += s; ti += i
        key = table[fn[0]][k] or ("?", 0)
        agg[key][0] += s; agg[key][1] += i
    lines.append("### hottest source lines of %s (stall samples / executed warp instructions)" % kname)
    lines.append("")
    lines.append("| samples | instructions | file:line | source |")
    lines.append("|---|---|---|---|")
    cache = {}
    for (f, l), (s, i) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:14]:
        p = os.path.join(ROOT, "image_compression_2_b200", "csrc", f)
        if f not in cache:
            cache[f] = open(p).read().splitlines() if os.path.exists(p) else []
        text = cache[f][l - 1].strip()[:90].replace("|", "\\|") if 0 < l <= len(cache[f]) else ""
        lines.append("| %.1f%% | %.1f%% | %s:%d | `%s` |" % (100 * s / ts, 100 * i / ti, f, l, text))
    lines.append("")
open(out, "w").write("\n".join(lines) + "\n")
print("wrote", out)
