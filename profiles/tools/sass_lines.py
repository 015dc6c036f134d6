"""Per-source-line share of stall samples and executed instructions: joins the SASS rows of
`ncu -i X.ncu-rep --page source --csv` with the line table of `nvdisasm -g <cubin>` (same instruction order).
usage: sass_lines.py <source_page.csv> <nvdisasm_-g_output> <kernel-name-substring> <source file>"""
import collections, csv, re, sys
src_csv, sass, kname, srcfile = sys.argv[1:5]
lines = open(sass).read().splitlines()
start = [i for i, l in enumerate(lines) if l.startswith('.text.') and kname in l][0]
cur, seq = None, []
for l in lines[start + 1:]:
    if l.startswith('.text.') or (l.startswith('//--------------------- .') and seq):
        break
    m = re.search(r'//## File "(?:.*/)?([^/"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1), int(m.group(2)))
        continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m:
        seq.append((int(m.group(1), 16), cur, m.group(2)))
rows = list(csv.reader(open(src_csv)))[2:]
base = int(rows[0][0], 16)
S, I = collections.Counter(), collections.Counter()
for r, (addr, cur, txt) in zip(rows, seq):
    assert int(r[0], 16) - base == addr, (r[0], addr)
    S[cur] += int(r[2]); I[cur] += int(r[5])
ts, ti = sum(S.values()), sum(I.values())
print('SASS rows', len(rows), 'matched', len(seq), 'samples', ts, 'warp instructions', ti)
text = open(srcfile).read().splitlines()
name = srcfile.split('/')[-1]
for k in sorted(S, key=lambda k: (k[0], k[1]) if k else ('', 0)):
    if S[k] > ts * 0.01 or I[k] > ti * 0.012:
        f, ln = k if k else ('?', 0)
        t = text[ln - 1].strip()[:95] if f == name else f
        print(f"{f[:14]:14s}{ln:4d} {100*S[k]/ts:5.1f}% smp {100*I[k]/ti:5.1f}% inst | {t}")
