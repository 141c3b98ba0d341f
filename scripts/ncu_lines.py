"""Attribute executed warp instructions of one kernel to source lines.
   python scripts/ncu_lines.py <ncu source-page csv> <nvdisasm -g -c listing> <mangled-prefix> <units (bars/ticks)>"""
import re, csv, collections, sys
srccsv, sass, prefix, units = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
lines = open(sass).read().split('\n')
start = [i for i, l in enumerate(lines) if l.startswith(prefix)][0]
cur = None; amap = {}
for l in lines[start:]:
    if l.startswith('//-----') and amap: break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    m = re.match(r'\s+/\*([0-9a-f]+)\*/\s+(.*?);', l)
    if m: amap[int(m.group(1), 16)] = (cur, m.group(2))
rows = list(csv.reader(open(srccsv)))
hdr = rows[1]; ci = {h: i for i, h in enumerate(hdr)}
base = int(rows[2][0], 16)
agg = collections.Counter(); stall = collections.Counter(); tot = 0
for r in rows[2:]:
    if len(r) < 6: continue
    off = int(r[0], 16) - base
    n = int(r[ci['Instructions Executed']]); tot += n
    loc = amap.get(off, (None, ''))[0]
    agg[loc] += n
    stall[loc] += int(r[ci['# Samples']] or 0)
files = {}
def text(loc):
    if not loc: return ''
    import glob
    if loc[0] not in files:
        g = glob.glob('/root/repo/finmlkit_b200/csrc/' + loc[0])
        files[loc[0]] = open(g[0]).read().split('\n') if g else None
    f = files[loc[0]]
    return f[loc[1] - 1].strip()[:80] if f and loc[1] <= len(f) else ''
ts = sum(stall.values())
print(f'total warp instructions {tot}  per unit {tot / units:.1f}   samples {ts}')
for loc, n in agg.most_common(int(sys.argv[5]) if len(sys.argv) > 5 else 40):
    print(f"{n / units:8.1f} {100 * n / tot:5.1f}%  smp {100 * stall[loc] / max(ts, 1):5.1f}%  {loc}  {text(loc)}")
