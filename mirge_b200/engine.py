"""numpy-facing wrapper of one device context (include/mirge_b200.h).

Everything here is plumbing: argument marshalling, error translation and
buffer allocation.  All arithmetic happens in the CUDA library.
"""
import ctypes as C

import numpy as np

from . import _capi
from ._capi import lib, MirgeError

LIB_IDS = {'mirna': 0, 'hairpin': 1, 'mature_trna': 2, 'pre_trna': 3, 'snorna': 4,
           'rrna': 5, 'ncrna_others': 6, 'mrna': 7, 'spike-in': 9}
# library slot searched by each round (utils/runAnnotationPipeline.py:574-599)
ROUND_LIB = [0, 1, 2, 3, 4, 5, 6, 7, 0, 9]
N_TOTALS = 12
TOTALS_KEYS = ['trimmedUniq', 'mirnaReads', 'hairpinReads', 'maturetrnaReads', 'pretrnaReads', 'snornaReads',
               'rrnaReads', 'ncrnaOthersReads', 'mrnaReads', 'spikeInReads', 'remReads']


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Engine:
    def __init__(self, n_samples=1, spike_in=False, max_unique=1 << 22, max_read_len=64, device=0):
        self.S = int(n_samples)
        self.spike_in = bool(spike_in)
        self.max_read_len = 64 if max_read_len <= 64 else 128
        self.max_unique = int(max_unique)
        self.h = _capi.ctx_p()
        rc = lib.mirge_create(device, self.S, int(self.spike_in), int(max_unique), int(max_read_len), C.byref(self.h))
        if rc:
            msg = lib.mirge_last_error(self.h if self.h else None)
            if self.h:
                lib.mirge_destroy(self.h)
                self.h = None
            raise MirgeError(rc, (msg or b'').decode())
        self.lib_names = {}
        self._keep = []     # host buffers that must outlive asynchronous copies

    # -- lifetime ---------------------------------------------------------------
    def close(self):
        if getattr(self, 'h', None):
            lib.mirge_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc:
            raise MirgeError(rc, lib.mirge_last_error(self.h).decode())

    def set_stream(self, cuda_stream):
        self._ck(lib.mirge_set_stream(self.h, C.c_void_p(int(cuda_stream))))

    def sync(self):
        self._ck(lib.mirge_sync(self.h))
        self._keep.clear()

    # -- stage 1 ----------------------------------------------------------------
    def set_adapter(self, spec, error_rate=0.12, min_overlap=3, min_length=16, qual_cutoff=10):
        self._ck(lib.mirge_set_adapter(self.h, spec.encode(), error_rate, min_overlap, min_length, qual_cutoff))

    @staticmethod
    def _as_array(buf):
        if isinstance(buf, np.ndarray):
            a = buf if buf.dtype == np.uint8 else buf.view(np.uint8)
            return np.ascontiguousarray(a)
        return np.frombuffer(buf, dtype=np.uint8)

    @classmethod
    def detect_phred(cls, buf):
        a = cls._as_array(buf)
        return int(lib.mirge_detect_phred(_ptr(a), a.nbytes))

    def feed_fastq(self, sample, buf, phred_base=None):
        """One record-aligned chunk of FASTQ text held in host memory."""
        a = self._as_array(buf)
        if phred_base is None:
            phred_base = self.detect_phred(a)
        self._keep.append(a)
        self._ck(lib.mirge_feed_fastq(self.h, sample, _ptr(a), a.nbytes, phred_base))
        return phred_base

    def feed_fastq_device(self, sample, dev_ptr, nbytes, phred_base=33):
        self._ck(lib.mirge_feed_fastq_device(self.h, sample, C.c_void_p(int(dev_ptr)), int(nbytes), phred_base))

    def stats(self):
        p = np.zeros(self.S, np.uint64)
        k = np.zeros(self.S, np.uint64)
        u = np.zeros(1, np.uint64)
        self._ck(lib.mirge_get_stats(self.h, _ptr(p), _ptr(k), _ptr(u)))
        self._keep.clear()
        return p, k, int(u[0])

    # -- stage 2 ----------------------------------------------------------------
    def load_library(self, lib_id, names, seqs):
        if isinstance(lib_id, str):
            lib_id = LIB_IDS[lib_id]
        off = np.zeros(len(seqs) + 1, np.uint64)
        if len(seqs):
            off[1:] = np.cumsum([len(s) for s in seqs], dtype=np.uint64)
        cat = np.frombuffer(''.join(seqs).encode() or b'\0', dtype=np.uint8)
        self._ck(lib.mirge_load_library(self.h, lib_id, len(seqs), _ptr(cat), _ptr(off)))
        self.lib_names[lib_id] = list(names)

    def annotate(self):
        self._ck(lib.mirge_annotate(self.h))

    def annotate_while_feeding(self, on=True):
        """Run the mapping rounds region by region inside feed_fastq (they hide behind the H2D copies)."""
        self._ck(lib.mirge_set_option(self.h, 1, int(on)))

    # -- genome passes ------------------------------------------------------------

    def pair_seeds(self, threshold=16):
        """MIRGE_OPT_PAIR_SEEDS: pair-of-pieces seed tables for short queries wherever the plain seeds would offer at
        least `threshold` candidates per query (1 = wherever possible, 0 = never).  Set before loading libraries."""
        self._ck(lib.mirge_set_option(self.h, 2, int(threshold)))
    def load_genome(self, names, seqs):
        """index.Libs/<sp>_genome (library slot 8)."""
        off = np.zeros(len(seqs) + 1, np.uint64)
        if len(seqs):
            off[1:] = np.cumsum([len(s) for s in seqs], dtype=np.uint64)
        cat = np.frombuffer(''.join(seqs).encode() or b'\0', dtype=np.uint8)
        self._ck(lib.mirge_load_genome(self.h, len(seqs), _ptr(cat), _ptr(off)))
        self.lib_names[8] = list(names)

    def genome_pass(self, reads, mode=0, param=1, seedlen=28, trim3=0, norc=False):
        """`bowtie genome -n/-v ... -a` for a list of sequences (str): counts[i][k] = alignments with k mismatches."""
        n = len(reads)
        counts = np.zeros((n, 4), np.uint32)
        if n == 0:
            return counts
        lens = np.array([len(r) for r in reads], np.uint8)
        stride = max(1, int(lens.max()))
        buf = np.zeros((n, stride), np.uint8)
        for i, r in enumerate(reads):
            buf[i, :len(r)] = np.frombuffer(r.encode() if isinstance(r, str) else r, dtype=np.uint8)
        self._ck(lib.mirge_genome_pass(self.h, n, _ptr(buf), stride, _ptr(lens), mode, param, seedlen, trim3, int(norc), _ptr(counts)))
        return counts

    def genome_pass_ids(self, ids, mode=0, param=1, seedlen=28, trim3=0, norc=False):
        """Same for unique sequences of the device table (ids = row numbers of export_unique)."""
        ids = np.ascontiguousarray(ids, dtype=np.uint32)
        counts = np.zeros((len(ids), 4), np.uint32)
        if len(ids):
            self._ck(lib.mirge_genome_pass_ids(self.h, len(ids), _ptr(ids), mode, param, seedlen, trim3, int(norc), _ptr(counts)))
        return counts

    # -- stage 3 ----------------------------------------------------------------
    def tally(self):
        self._ck(lib.mirge_tally(self.h))

    # -- export -----------------------------------------------------------------
    def export_unique(self, with_seq=True):
        _, _, U = self.stats()
        stride = self.max_read_len
        seq = np.zeros((U, stride), np.uint8) if with_seq else None
        ln = np.zeros(U, np.uint8)
        cnt = np.zeros((U, self.S), np.uint32)
        ro = np.zeros(U, np.int8)
        rf = np.zeros(U, np.int32)
        ps = np.zeros(U, np.int32)
        mm = np.zeros(U, np.int8)
        self._ck(lib.mirge_export_unique(self.h, U, stride, _ptr(seq), _ptr(ln), _ptr(cnt), _ptr(ro), _ptr(rf), _ptr(ps), _ptr(mm)))
        out = dict(len=ln, counts=cnt, round_of=ro, ref_of=rf, pos_of=ps, mm_of=mm)
        if with_seq:
            out['seq'] = seq.view('S%d' % stride).reshape(U)   # numpy strips the zero padding
        return out

    def export_tables(self):
        nm = len(self.lib_names.get(0, []))
        q = np.zeros((nm, self.S), np.int64)
        i = np.zeros((nm, self.S), np.int64)
        t = np.zeros((self.S, N_TOTALS), np.int64)
        lh = np.zeros((self.max_read_len + 1, self.S), np.uint64)
        rp = np.zeros(10, np.uint64)
        ra = np.zeros(10, np.uint64)
        self._ck(lib.mirge_export_tables(self.h, _ptr(q), _ptr(i), _ptr(t), _ptr(lh), _ptr(rp), _ptr(ra)))
        return dict(mir_quant=q, mir_iscan=i, lib_totals=t, len_hist=lh, round_processed=rp, round_aligned=ra)

    def export_tables_prelim(self):
        """Tables that exist before tally(): read-length histogram and per-round statistics."""
        lh = np.zeros((self.max_read_len + 1, self.S), np.uint64)
        rp = np.zeros(10, np.uint64)
        ra = np.zeros(10, np.uint64)
        self._ck(lib.mirge_export_tables(self.h, None, None, None, _ptr(lh), _ptr(rp), _ptr(ra)))
        return dict(len_hist=lh, round_processed=rp, round_aligned=ra)

    # -- multi-GPU exchange -------------------------------------------------------
    def record_bytes(self):
        return int(lib.mirge_record_bytes(self.h))

    def export_records_device(self, n_parts, dev_ptr, cap_records, counts_dev_ptr=None):
        """Records grouped by owner into dev_ptr.  With counts_dev_ptr (device int64[n_parts]) the call is fully
        asynchronous and returns None; without it the group sizes are fetched (synchronises) and returned."""
        if counts_dev_ptr is not None:
            self._ck(lib.mirge_export_records_device(self.h, n_parts, C.c_void_p(int(dev_ptr)), int(cap_records), C.c_void_p(int(counts_dev_ptr))))
            return None
        import torch
        pc = torch.empty(n_parts, dtype=torch.int64, device='cuda')
        torch.cuda.synchronize()         # the engine may run on a stream of its own
        self._ck(lib.mirge_export_records_device(self.h, n_parts, C.c_void_p(int(dev_ptr)), int(cap_records), C.c_void_p(pc.data_ptr())))
        self.sync()
        return pc.cpu().numpy().astype(np.uint64)

    def tables_words(self):
        return int(lib.mirge_tables_words(self.h))

    def pack_tables_device(self, dev_ptr):
        self._ck(lib.mirge_pack_tables_device(self.h, C.c_void_p(int(dev_ptr))))

    def unpack_tables_device(self, dev_ptr):
        self._ck(lib.mirge_unpack_tables_device(self.h, C.c_void_p(int(dev_ptr))))

    def reset(self):
        self._ck(lib.mirge_reset(self.h))

    def launch_count(self):
        return int(lib.mirge_launch_count(self.h))

    def force_slow_path(self, on=True, literal=False):
        """Test hook: send every read through the slow queue -- the all-column bit-vector search, or
        (literal=True) the cell-for-cell cutadapt DP."""
        self._ck(lib.mirge_debug(self.h, (3 if literal else 1) if on else 0, None))

    def set_trim_mode(self, mode):
        """Test hook: raw mirge_debug mode (0 default, 1 all-column search, 3 literal DP, 4 prefilter off)."""
        self._ck(lib.mirge_debug(self.h, int(mode), None))

    def trim_classes(self):
        """(reads through the all-column search, prefilter class-1 reads, prefilter class-2 reads, feed regions
        re-parsed by the exact K0 kernel)."""
        n = np.zeros(4, np.uint64)
        self._ck(lib.mirge_trim_classes(self.h, _ptr(n)))
        return tuple(int(x) for x in n)

    def slow_reads(self):
        n = np.zeros(1, np.uint64)
        self._ck(lib.mirge_debug(self.h, -1, _ptr(n)))
        return int(n[0])

    def reset_unique(self):
        self._ck(lib.mirge_reset_unique(self.h))

    def import_records_device(self, dev_ptr, n_records):
        self._ck(lib.mirge_import_records_device(self.h, C.c_void_p(int(dev_ptr)), int(n_records)))

    # -- measurement ---------------------------------------------------------------
    def profile_enable(self, on=True):
        self._ck(lib.mirge_profile_enable(self.h, int(on)))

    def profile_read(self):
        ms = np.zeros(32, np.float64)
        n = np.zeros(32, np.uint64)
        names = (C.c_char_p * 32)()
        self._ck(lib.mirge_profile_read(self.h, _ptr(ms), _ptr(n), C.cast(names, C.c_void_p)))
        return {names[k].decode(): (float(ms[k]), int(n[k])) for k in range(32) if names[k]}


def owner_of(seq, n_parts):
    """Owner rank of a sequence in the multi-GPU exchange."""
    b = seq if isinstance(seq, bytes) else seq.encode()
    return int(lib.mirge_hash_sequence(b, len(b)) % n_parts)
