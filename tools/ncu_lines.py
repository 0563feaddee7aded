"""Per-source-line executed warp instructions and stall samples from an ncu report captured with --import-source on.
usage: python tools/ncu_lines.py report.ncu-rep [top N]"""
import csv, subprocess, sys, collections
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 60
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[h]
ci = hdr.index("Instructions Executed"); si = hdr.index("# Samples"); so = hdr.index("Source")
print("sass rows", len(rows) - h - 1)
ops = collections.Counter(); tot = 0
for r in rows[h + 1:]:
    try: n = int(r[ci])
    except Exception: continue
    op = r[so].split()[0] if not r[so].startswith("@") else r[so].split()[1]
    ops[op.split(".")[0]] += n; tot += n
print("total warp instructions", tot)
for k, v in ops.most_common(40): print("  %-12s %10d %5.1f%%" % (k, v, 100.0 * v / tot))

# per source line, with the opcodes under it (combined view: a line row followed by its SASS rows)
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
agg = {}
fp = "?"; hdr = None; cur = None
for r in rows:
    if not r: continue
    if r[0] == "File Path": fp = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; ci = hdr.index("Instructions Executed"); si = hdr.index("# Samples"); continue
    if hdr is None or len(r) <= ci: continue
    if r[0] != "":
        cur = (fp, r[0]); a = agg.setdefault(cur, [0, 0, r[1].strip()[:110], collections.Counter()])
        try: a[0] += int(r[ci]); a[1] += int(r[si])
        except Exception: pass
    elif cur is not None:
        try: n = int(r[ci])
        except Exception: continue
        src = r[3].split()
        if not src: continue
        op = src[1] if src[0].startswith("@") and len(src) > 1 else src[0]
        agg[cur][3][op.split(".")[0]] += n
items = sorted(agg.items(), key=lambda kv: -kv[1][0])
T = sum(v[0] for v in agg.values()); TS = max(1, sum(v[1] for v in agg.values()))
print("per line: total", T, "samples", TS)
for (f, ln), v in items[:top]:
    print("%9d %5.1f%% smp %5.1f%% %s:%s %s\n%40s%s" % (v[0], 100.0 * v[0] / T, 100.0 * v[1] / TS, f, ln, v[2], "", " ".join("%s:%d" % (k, c // 1000) for k, c in v[3].most_common(6))))
