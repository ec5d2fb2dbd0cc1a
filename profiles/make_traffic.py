"""profiles/r2_traffic.json from the committed-round ncu --set full captures: DRAM bytes (read + write) per bench step for
each kernel family, which bench.py reports as roofline.traffic when it runs the same 25 M-read workload.
    python profiles/make_traffic.py gpurun_out/r2_k0k1.ncu-rep gpurun_out/r2_k2.ncu-rep
"""
import csv
import io
import json
import os
import subprocess
import sys

UNIT = {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.0, 'Tbyte': 1e12}
fam = {'k0_parse_pack': 0.0, 'k1_trim_collapse': 0.0, 'k2_rounds': 0.0}
per_kernel = []
for rep in sys.argv[1:]:
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    iname = hdr.index('Kernel Name')
    ir, iw, it = hdr.index('dram__bytes_read.sum'), hdr.index('dram__bytes_write.sum'), hdr.index('gpu__time_duration.sum')
    for r in rows[2:]:
        name = r[iname]
        b = float(r[ir].replace(',', '')) * UNIT[units[ir]] + float(r[iw].replace(',', '')) * UNIT[units[iw]]
        f = 'k0_parse_pack' if 'k0_parse' in name else 'k2_rounds' if 'k2_round' in name else 'k1_trim_collapse' if 'k1_' in name else None
        if f:
            fam[f] += b
        per_kernel.append({'kernel': name.split('(')[0].replace('void ', ''), 'dram_bytes': b, 'time': r[it] + ' ' + units[it]})
json.dump({'workload': '25M reads, L=50, libs scale 1.0 (bench.py default)', 'captures': [os.path.basename(x) for x in sys.argv[1:]],
           'dram_bytes_per_step': fam, 'per_kernel': per_kernel}, open(os.path.join(os.path.dirname(__file__), 'r2_traffic.json'), 'w'), indent=1)
print(json.dumps(fam))
