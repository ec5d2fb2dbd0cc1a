"""ctypes binding of libmirge_b200.so (include/mirge_b200.h).

There is no CPU fallback: importing this module fails loudly when the CUDA
library has not been built (run `python -c "import __graft_entry__ as g; g.build()"`
or `make -C mirge_b200/csrc`).
"""
import ctypes as C
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'csrc', 'libmirge_b200.so')
HEADER_PATH = os.path.join(os.path.dirname(_HERE), 'include', 'mirge_b200.h')

if not os.path.exists(LIB_PATH):
    raise ImportError(
        'mirge_b200: %s is missing -- the CUDA extension is required (no CPU fallback). '
        'Build it with `make -C mirge_b200/csrc`.' % LIB_PATH)

lib = C.CDLL(LIB_PATH)

ctx_p = C.c_void_p
u64 = C.c_uint64
vp = C.c_void_p

_SIGS = {
    'mirge_create': (C.c_int, [C.c_int, C.c_int, C.c_int, u64, C.c_int, C.POINTER(ctx_p)]),
    'mirge_destroy': (None, [ctx_p]),
    'mirge_last_error': (C.c_char_p, [ctx_p]),
    'mirge_set_stream': (C.c_int, [ctx_p, vp]),
    'mirge_sync': (C.c_int, [ctx_p]),
    'mirge_reset': (C.c_int, [ctx_p]),
    'mirge_launch_count': (u64, [ctx_p]),
    'mirge_debug': (C.c_int, [ctx_p, C.c_int, vp]),
    'mirge_trim_classes': (C.c_int, [ctx_p, vp]),
    'mirge_prefilter_tables': (C.c_int, [C.c_char_p, C.c_double, C.c_int, vp, vp, vp, vp, vp]),
    'mirge_set_adapter': (C.c_int, [ctx_p, C.c_char_p, C.c_double, C.c_int, C.c_int, C.c_int]),
    'mirge_detect_phred': (C.c_int, [vp, u64]),
    'mirge_feed_fastq': (C.c_int, [ctx_p, C.c_int, vp, u64, C.c_int]),
    'mirge_feed_fastq_device': (C.c_int, [ctx_p, C.c_int, vp, u64, C.c_int]),
    'mirge_get_stats': (C.c_int, [ctx_p, vp, vp, vp]),
    'mirge_load_library': (C.c_int, [ctx_p, C.c_int, C.c_int, vp, vp]),
    'mirge_load_genome': (C.c_int, [ctx_p, C.c_int, vp, vp]),
    'mirge_genome_pass': (C.c_int, [ctx_p, u64, vp, C.c_int, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp]),
    'mirge_genome_pass_ids': (C.c_int, [ctx_p, u64, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp]),
    'mirge_annotate': (C.c_int, [ctx_p]),
    'mirge_set_option': (C.c_int, [ctx_p, C.c_int, C.c_int]),
    'mirge_pair_seed_plan': (C.c_int, [C.c_int, C.c_uint64, C.c_int, C.POINTER(C.c_int)]),
    'mirge_tally': (C.c_int, [ctx_p]),
    'mirge_export_unique': (C.c_int, [ctx_p, u64, C.c_int, vp, vp, vp, vp, vp, vp, vp]),
    'mirge_export_tables': (C.c_int, [ctx_p, vp, vp, vp, vp, vp, vp]),
    'mirge_record_bytes': (C.c_int, [ctx_p]),
    'mirge_export_records_device': (C.c_int, [ctx_p, C.c_int, vp, u64, vp]),
    'mirge_reset_unique': (C.c_int, [ctx_p]),
    'mirge_import_records_device': (C.c_int, [ctx_p, vp, u64]),
    'mirge_tables_words': (u64, [ctx_p]),
    'mirge_pack_tables_device': (C.c_int, [ctx_p, vp]),
    'mirge_unpack_tables_device': (C.c_int, [ctx_p, vp]),
    'mirge_hash_sequence': (u64, [C.c_char_p, C.c_int]),
    'mirge_profile_enable': (C.c_int, [ctx_p, C.c_int]),
    'mirge_profile_read': (C.c_int, [ctx_p, vp, vp, vp]),
}

for _name, (_res, _args) in _SIGS.items():
    _f = getattr(lib, _name)      # AttributeError here means the .so is stale
    _f.restype = _res
    _f.argtypes = _args


def declared_symbols():
    """Entry points declared in include/mirge_b200.h (used by the symbol test)."""
    with open(HEADER_PATH) as f:
        return sorted(set(re.findall(r'\b(mirge_[a-z_0-9]+)\s*\(', f.read())))


class MirgeError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__('mirge_b200 error %d: %s' % (code, msg))
        self.code = code
