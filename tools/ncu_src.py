#!/usr/bin/env python
"""Per-SASS-line view of an ncu report: python tools/ncu_src.py rep.ncu-rep <kernel regex> [top]"""
import csv, subprocess, sys
rep, kre = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', 'regex:' + kre], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
# sections: "Kernel Name" row, header row, data rows
sec = []
i = 0
while i < len(rows):
    if rows[i] and rows[i][0] == 'Kernel Name':
        name = rows[i][1]; hdr = rows[i + 1]; i += 2; data = []
        while i < len(rows) and not (rows[i] and rows[i][0] == 'Kernel Name'):
            if len(rows[i]) == len(hdr): data.append(rows[i])
            i += 1
        sec.append((name, hdr, data))
    else:
        i += 1
for name, hdr, data in sec:
    ia = hdr.index('Instructions Executed'); isamp = hdr.index('# Samples'); isrc = hdr.index('Source')
    tot = sum(int(r[ia]) for r in data); ts = sum(int(r[isamp]) for r in data)
    print('==', name[:70], 'instr %.1fM samples %d sass %d' % (tot / 1e6, ts, len(data)))
    idx = sorted(range(len(data)), key=lambda k: -int(data[k][isamp]))[:top]
    for k in sorted(idx):
        r = data[k]
        print('%5d %-72s inst=%8.1fM samp=%6d (%4.1f%%)' % (k, r[isrc].strip()[:72], int(r[ia]) / 1e6, int(r[isamp]), 100 * int(r[isamp]) / ts))
