"""Attributes the per-instruction counters of an ncu source page (SASS view, csv) to CUDA source lines using
nvdisasm --print-line-info of the same cubin.  usage: ncu_lines.py source.csv all.sass kernel_substring [top]"""
import re, csv, collections, sys
src_csv, sass, pat = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 50
lines = open(sass).read().split('\n')
start = next(i for i, l in enumerate(lines) if l.startswith('.text.') and pat in l)
cur = None; seq = []
for l in lines[start + 1:]:
    if l.startswith('//--------------------- .text'): break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m: seq.append((int(m.group(1), 16), m.group(2), cur))
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]; data = rows[2:]
assert len(data) == len(seq), (len(data), len(seq))
ie = hdr.index('Instructions Executed'); iss = hdr.index('# Samples')
agg = collections.Counter(); samp = collections.Counter(); tot = 0
for r, (off, txt, cur) in zip(data, seq):
    n = int(r[ie]); s = int(r[iss]); tot += n
    agg[cur] += n; samp[cur] += s
ts = sum(samp.values())
print('total inst', tot, 'samples', ts)
for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:top]:
    print(k, v, f"{100*v/tot:.1f}%  samples {100*samp[k]/ts:.1f}%")
print('--- top stalled instructions')
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
items = []
for idx, (r, (off, txt, cur)) in enumerate(zip(data, seq)):
    items.append((int(r[iss]), idx, txt, cur, int(r[ie]), {h: int(r[i]) for i, h in stall_cols if int(r[i]) > 0}))
for s, idx, txt, cur, n, st in sorted(items, reverse=True)[:25]:
    t3 = sorted(st.items(), key=lambda kv: -kv[1])[:2]
    print(f"{100*s/ts:5.1f}% L{cur[1] if cur else None} n={n:9d} {txt[:56]:56s} {t3}")
