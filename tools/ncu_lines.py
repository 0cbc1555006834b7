#!/usr/bin/env python
"""Per-CUDA-source-line hot spots: joins the SASS page of an .ncu-rep (stall samples, executed instructions
per SASS instruction) with nvdisasm --print-line-info of the SAME build of libstc_b200.so.

usage: tools/ncu_lines.py report.ncu-rep kernel_regex [launch_index] [top_n]
"""
import collections, csv, glob, io, os, re, subprocess, sys, tempfile
rep, rx = sys.argv[1], sys.argv[2]
idx = int(sys.argv[3]) if len(sys.argv) > 3 else 0
top = int(sys.argv[4]) if len(sys.argv) > 4 else 25
ROOT = os.path.dirname(os.path.abspath(__file__))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--kernel-name",
                      f"regex:{rx}"], capture_output=True, text=True).stdout
launches, cur, hdr = [], None, None
for row in csv.reader(io.StringIO(out)):
    if not row: continue
    if row[0] == "Kernel Name":
        cur = {"name": row[1], "rows": []}; launches.append(cur); hdr = None; continue
    if row[0] == "Address":
        hdr = row; continue
    if cur is not None and hdr is not None: cur["rows"].append(row)
L = launches[min(idx, len(launches) - 1)]
i_s, i_n, i_src = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Source")
mangled = None
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "stc_gnn_b200", "libstc_b200.so")], cwd=tmp, capture_output=True)
kname = re.sub(r"\(.*", "", L["name"]).split("::")[-1].split("<")[0]
lines = None
for cub in glob.glob(os.path.join(tmp, "*.cubin")):
    if "-" in os.path.basename(cub).split(".")[0]: continue
    dis = subprocess.run(["nvdisasm", "--print-line-info", "-c", cub], capture_output=True, text=True).stdout
    secs = re.split(r"\n//-+ \.text\.", dis)
    for sec in secs[1:]:
        head = sec.split("\n", 1)[0]
        if kname in head:
            cand = []
            cur_line = ("?", 0)
            for ln in sec.split("\n"):
                m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
                if m: cur_line = (os.path.basename(m.group(1)), int(m.group(2))); continue
                m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
                if m: cand.append((int(m.group(1), 16), cur_line, m.group(2).strip()))
            if lines is None or abs(len(cand) - len(L["rows"])) < abs(len(lines) - len(L["rows"])):
                lines = cand
if lines is None or len(lines) != len(L["rows"]):
    print(f"warning: SASS length mismatch ({0 if lines is None else len(lines)} vs {len(L['rows'])}): is the .so the profiled build?")
agg = collections.defaultdict(lambda: [0, 0])
for k, r in enumerate(L["rows"]):
    loc = lines[k][1] if lines and k < len(lines) else ("?", 0)
    agg[loc][0] += int(r[i_s] or 0); agg[loc][1] += int(r[i_n] or 0)
ts = sum(v[0] for v in agg.values()) or 1; tn = sum(v[1] for v in agg.values()) or 1
print(f"kernel {L['name'][:90]}\n launch {idx}: stall samples {ts}, warp instructions {tn}")
srcs = {}
def src(fn, ln):
    if fn not in srcs:
        p = glob.glob(os.path.join(ROOT, "stc_gnn_b200", "csrc", fn))
        srcs[fn] = open(p[0]).read().split("\n") if p else []
    return srcs[fn][ln - 1].strip()[:100] if 0 < ln <= len(srcs[fn]) else ""
for loc, (s, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100*s/ts:5.1f}% smp {100*n/tn:5.1f}% ins  {loc[0]}:{loc[1]:<4d} {src(*loc)}")
