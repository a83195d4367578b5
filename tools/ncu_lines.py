"""Join an ncu SASS page (csv) with nvdisasm -g line info: instruction and stall-sample shares per source line.

    cuobjdump -xelf all build/x.o; nvdisasm -g -c x.sm_100a.cubin > lines.txt
    ncu -i rep.ncu-rep --page source --csv --print-source sass --kernel-name K > sass.csv
    python tools/ncu_lines.py sass.csv lines.txt <mangled kernel name substring> [top] [samples]
"""
import csv, re, sys, collections

sass_csv, lines_txt, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
by_samples = len(sys.argv) > 5 and sys.argv[5] == 'samples'
# line info: list of (offset, file:line) for the kernel and the device functions that follow it
sections = {}
cur = None; line = None
for ln in open(lines_txt):
    m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
    if m:
        cur = m.group(1); sections[cur] = []; line = None; continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
    if m:
        line = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m and cur is not None:
        sections[cur].append((int(m.group(1), 16), line, m.group(2).strip()))
sec = [k for k in sections if kname in k]
assert sec, list(sections)[:5]
offs = sections[sec[0]]
rows = list(csv.reader(open(sass_csv)))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
base = int(data[0][ix["Address"]], 16) if data[0][ix["Address"]].startswith("0x") else int(data[0][ix["Address"]])
agg = collections.defaultdict(lambda: [0.0, 0.0])
tot_i = tot_s = 0.0
n = min(len(data), len(offs))
for k in range(n):
    r = data[k]
    e = float(r[ix["Instructions Executed"]] or 0); s = float(r[ix["# Samples"]] or 0)
    agg[offs[k][1]][0] += e; agg[offs[k][1]][1] += s
    tot_i += e; tot_s += s
print(f"instructions {tot_i:.0f}  samples {tot_s:.0f}  ({n} of {len(data)} SASS rows mapped)")
for ln, (e, s) in sorted(agg.items(), key=lambda kv: -kv[1][1 if by_samples else 0])[:top]:
    print(f"{str(ln):40s} {e / tot_i * 100:6.2f}% inst  {s / max(tot_s, 1) * 100:6.2f}% samples")
