"""Per-kernel timing of K1 / K2 variants selected through mirge_debug bits -- the cheap way to iterate on a kernel:
the 25 M-read workload is generated once (about 8 s), every variant then costs three feeds.

    gpurun -- python profiles/kernel_variants.py 25000000 0,0x100,0x200 0
        argv[1] reads, argv[2] K1 modes, argv[3] K2 modes (optional; loads the libraries), argv[4] exchange kernels (0/1),
        argv[5] pair-seed thresholds to run the K2 modes under (MIRGE_OPT_PAIR_SEEDS; default: the library default)
    mode bits: 1 all-column search for every read, 2 literal DP, 4 prefilter off, 0x100 no count updates (diagnostic),
    0x200 classification only (diagnostic), 0x800 count cache off,
    0x1000 (K2) no poly-T bypass
Diagnostic modes give wrong tables on purpose; they only attribute time."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from synth.world import World
from mirge_b200.engine import Engine
n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 25_000_000
modes = [int(x, 0) for x in sys.argv[2].split(',') if x] if len(sys.argv) > 2 else [0, 0x800, 0x100, 0x200, 4]
world = World(seed=20180626, scale=1.0)
rec = world.record_bytes(50)
host = torch.empty(n_reads * rec, dtype=torch.uint8, pin_memory=True)
world.fastq(n_reads, seed=1000, L=50, out=host.numpy())
dev_text = host.to('cuda:0')
eng = Engine(1, False, max(1 << 20, int(n_reads * 0.6)), 64, device=0)
eng.set_adapter('TGGAATTCTCGGGTGCCAAGGAACTCCAG')
for mode in modes:
    eng.set_trim_mode(mode)
    for it in range(3):
        eng.reset()
        if it == 1:
            eng.profile_enable(True)
        eng.feed_fastq_device(0, dev_text.data_ptr(), n_reads * rec, 33)
        eng.sync()
    prof = eng.profile_read()
    eng.profile_enable(False)
    p, k, U = eng.stats()
    print(hex(mode), 'kept', int(k[0]), 'U', U, {n: round(v[0] / max(1, v[1]), 3) for n, v in prof.items() if n.startswith('k1') and v[1] > 0}, flush=True)
K2MODES = [int(x, 0) for x in sys.argv[3].split(",") if x] if len(sys.argv) > 3 else []
# K2 variants (needs libraries)
from mirge_b200.engine import LIB_IDS
PAIRS = [int(x) for x in sys.argv[5].split(",") if x] if len(sys.argv) > 5 else [None]
for ps, mode in [(a, b) for a in (PAIRS if K2MODES else []) for b in K2MODES]:
    if mode == K2MODES[0]:
        if ps is not None:
            eng.pair_seeds(ps)
        for t, i in LIB_IDS.items():
            eng.load_library(i, *world.libs[t])
        print('pair_seeds', ps, flush=True)
    eng.set_trim_mode(mode)
    for it in range(3):
        eng.reset()
        eng.feed_fastq_device(0, dev_text.data_ptr(), n_reads * rec, 33)
        if it == 1:
            eng.profile_enable(True)
        eng.annotate(); eng.tally()
        eng.sync()
    prof = eng.profile_read()
    eng.profile_enable(False)
    d = {n: round(v[0] / max(1, v[1]), 3) for n, v in prof.items() if n.startswith('k2') and v[1] > 0}
    print(hex(mode), 'K2 total', round(sum(d.values()), 3), d, flush=True)

# exchange kernels on one GPU: export for 2 owners, reset, import everything back (argv[4] = 1)
if len(sys.argv) > 4 and int(sys.argv[4]):
    eng.set_trim_mode(0)
    for it in range(3):
        eng.reset()
        eng.feed_fastq_device(0, dev_text.data_ptr(), n_reads * rec, 33)
        eng.sync()
        if it == 1:
            eng.profile_enable(True)
        p, k, U = eng.stats()
        buf = torch.empty(U * eng.record_bytes(), dtype=torch.uint8, device='cuda')
        pc = eng.export_records_device(2, buf.data_ptr(), U)
        eng.reset_unique()
        eng.import_records_device(buf.data_ptr(), int(pc.sum()))
        eng.sync()
    prof = eng.profile_read()
    eng.profile_enable(False)
    print('exchange', {n: round(v[0] / max(1, v[1]), 3) for n, v in prof.items() if n.startswith(('exchange', 'reset')) and v[1] > 0}, 'U', U, flush=True)
