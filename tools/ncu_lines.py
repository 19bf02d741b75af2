"""Join an `ncu --page source --csv` export (SASS rows) with nvdisasm -g line info and print
instructions executed / stall samples per source line.
usage: ncu_lines.py src.csv disasm.sass kernel_substring [top]"""
import csv, re, sys, collections
src, sass, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
rows = list(csv.reader(open(src)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
h = rows[hi]
ci, cs, csmp = h.index("Instructions Executed"), h.index("Source"), h.index("# Samples")
body = []
for r in rows[hi + 1:]:
    if r and r[0] in ("Address", "Kernel Name"):     # next kernel of a multi-kernel export
        break
    if len(r) > ci:
        body.append(r)
ins = [(r[cs], int(r[ci] or 0), int(r[csmp] or 0)) for r in body]
# nvdisasm: find function, walk lines, track current "//## File ..., line N"
lines = open(sass).read().splitlines()
start = [i for i, l in enumerate(lines) if l.startswith(".text.") and kern in l or (l.strip().startswith(".text.") and kern in l)]
start = start[0]
cur = None; seq = []
for l in lines[start + 1:]:
    if l.startswith(".text.") or l.strip().startswith(".section"):
        if seq: break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = int(m.group(2)); continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m: seq.append((cur, m.group(2)))
print(len(ins), len(seq), file=sys.stderr)
agg = collections.Counter(); smp = collections.Counter()
n = min(len(ins), len(seq))
for (ln, _), (s, c, sm) in zip(seq[:n], ins[:n]):
    agg[ln] += c; smp[ln] += sm
tot = sum(agg.values()); tots = sum(smp.values())
srcl = open("/root/repo/pychips_b200/csrc/median.cu").read().splitlines() if len(sys.argv) < 6 else open(sys.argv[5]).read().splitlines()
print(f"total inst {tot} samples {tots}")
for ln, c in agg.most_common(top):
    t = srcl[ln - 1].strip()[:90] if ln and ln <= len(srcl) else ""
    print(f"{ln:5d} {100*c/tot:5.1f}% inst {100*smp[ln]/max(tots,1):5.1f}% smp  {t}")
