"""Per-source-line executed-instruction profile: joins `nvdisasm -g -c` line info of one kernel with the per-SASS-instruction
counters of an `ncu --page source --csv` export (same instruction order).

    python scripts/line_profile.py <cubin> <mangled-kernel-substring> <ncu_source.csv> [top] [stall]
"""
import collections, csv, re, subprocess, sys

cubin, kern, ncsv = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
by_stall = len(sys.argv) > 5 and sys.argv[5] == "stall"      # order by stall samples instead of executed instructions
txt = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.split("\n")
# locate the function's text section
start = next(i for i, l in enumerate(txt) if l.startswith("_Z") and kern in l and l.rstrip().endswith(":"))
lines = []
cur = None
inl = []
for l in txt[start + 1:]:
    if l.strip().startswith(".section") or l.startswith("//-----"):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        # "inlined at" chain: attribute to the outermost texture.cu line as well
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        lines.append((cur, l.strip()))
rows = list(csv.reader(open(ncsv)))
hi = [i for i, r in enumerate(rows) if len(r) > 3 and r[0] == 'Address'][0]
hdr = rows[hi]
iex = hdr.index('Instructions Executed'); isamp = hdr.index('Warp Stall Sampling (All Samples)')
data = [r for r in rows[hi + 1:] if len(r) > iex and r[0].startswith('0x')]
assert len(data) == len(lines), (len(data), len(lines))
ex = collections.Counter(); sm = collections.Counter()
for (cur, sass), r in zip(lines, data):
    ex[cur] += int(r[iex] or 0); sm[cur] += int(r[isamp] or 0)
tex, tsm = sum(ex.values()), sum(sm.values())
src = {}
def text(f, n):
    if f not in src:
        try: src[f] = open("/root/repo/nvdiffrast_b200/csrc/" + f).read().split("\n")
        except Exception: src[f] = []
    return src[f][n - 1].strip()[:110] if 0 < n <= len(src[f]) else ""
print("total warp-inst %d, samples %d" % (tex, tsm))
for cur, _ in (sm if by_stall else ex).most_common(top):
    v = ex[cur]
    print("%5.1f%% ex %5.1f%% stall  %s:%d  %s" % (100 * v / tex, 100 * sm[cur] / max(tsm, 1), cur[0] if cur else "?", cur[1] if cur else 0, text(*cur) if cur else ""))
