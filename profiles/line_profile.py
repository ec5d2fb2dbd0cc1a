#!/usr/bin/env python
"""Per-source-line warp-instruction / stall-sample shares of one kernel.

    python profiles/line_profile.py gpurun_out/prof.ncu-rep k1_trim_collapseILi2 [top_n]

Joins the SASS page of an ncu report (instructions executed, samples per SASS
address) with `nvdisasm -g` line info of the in-tree .so (built with -lineinfo).
"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, 'mirge_b200', 'csrc', 'libmirge_b200.so')


def main():
    rep, sym = sys.argv[1], sys.argv[2]
    top_n = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    tmp = tempfile.mkdtemp()
    subprocess.run(['cuobjdump', '-xelf', 'all', SO], cwd=tmp, capture_output=True)
    cubin = [f for f in os.listdir(tmp) if f.endswith('.cubin')][0]
    dis = subprocess.run(['nvdisasm', '-g', '-c', os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.split('\n')
    st = [i for i, l in enumerate(dis) if l.startswith('//-----') and sym in l][0]
    en = [i for i, l in enumerate(dis) if l.startswith('//-----') and i > st][0]
    cur, off2line = None, {}
    for l in dis[st:en]:
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r'\s+/\*([0-9a-f]{4,})\*/', l)
        if m:
            off2line[int(m.group(1), 16)] = cur
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    # a report may hold several launches: keep the N-th one (env LAUNCH, default 0)
    starts = [i for i, r in enumerate(rows) if r and r[0] == 'Kernel Name'] + [len(rows)]
    which = int(os.environ.get('LAUNCH', '0'))
    # ncu prints every launch twice (two views of the same SASS): step over the duplicates
    heads = starts[:-1]
    if len(heads) % 2 == 0 and all(rows[heads[k]][1] == rows[heads[k + 1]][1] for k in range(0, len(heads), 2)):
        which *= 2
    rows = rows[starts[which]:starts[which + 1]]
    hdr, data = rows[1], rows[2:]
    iA, iE, iS, iT = hdr.index('Address'), hdr.index('Instructions Executed'), hdr.index('# Samples'), hdr.index('Thread Instructions Executed')
    base = int(data[0][iA], 16)
    ins, smp, thr = collections.Counter(), collections.Counter(), collections.Counter()
    for r in data:
        k = off2line.get(int(r[iA], 16) - base)
        ins[k] += int(r[iE]); smp[k] += int(r[iS]); thr[k] += int(r[iT])
    ti, ts = sum(ins.values()), sum(smp.values())
    src = {}
    print('# %s  kernel %s: %d warp-instructions, %d samples' % (os.path.basename(rep), sym, ti, ts))
    print('%6s %6s %5s  %s' % ('instr', 'smpl', 'thr', 'source line'))
    for k, v in ins.most_common(top_n):
        txt = ''
        if k:
            p = os.path.join(ROOT, 'mirge_b200', 'csrc', k[0])
            if os.path.exists(p):
                src.setdefault(k[0], open(p).read().split('\n'))
                txt = src[k[0]][k[1] - 1].strip()[:100]
        print('%5.1f%% %5.1f%% %5.1f  %s:%s  %s' % (100 * v / ti, 100 * smp[k] / max(ts, 1), thr[k] / max(v, 1),
                                                  k[0] if k else '?', k[1] if k else '?', txt))


if __name__ == '__main__':
    main()
