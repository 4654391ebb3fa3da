"""Instruction mix / stall samples per opcode from `ncu --page source --csv` output."""
import csv, sys
from collections import Counter
rows = list(csv.reader(open(sys.argv[1])))
hdr = next(r for r in rows if 'Instructions Executed' in r)
data = [r for r in rows if len(r) == len(hdr) and r is not hdr and r[hdr.index('Instructions Executed')].isdigit()]
ia = hdr.index('Instructions Executed'); isamp = hdr.index('# Samples'); isrc = hdr.index('Source')
tot = sum(int(r[ia]) for r in data); tots = sum(int(r[isamp]) for r in data)
print('total warp-inst', tot, 'samples', tots, 'sass lines', len(data))
print('exec-count histogram', Counter(int(r[ia]) for r in data).most_common(6))
mix = Counter(); smp = Counter()
for r in data:
    toks = r[isrc].split()
    op = toks[1] if toks[0].startswith('@') else toks[0]
    op = op.split('.')[0]
    mix[op] += int(r[ia]); smp[op] += int(r[isamp])
for op, n in mix.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 25):
    print(f"{op:10s} {n:10d} {100*n/tot:5.1f}%  samples {100*smp[op]/max(tots,1):5.1f}%")
if len(sys.argv) > 3:
    top = sorted(data, key=lambda r: -int(r[isamp]))[:int(sys.argv[3])]
    for r in top: print(r[isamp], r[ia], r[isrc][:100])
