"""mirge_b200 -- B200-native implementation of the miRge2.0 annotate-mode read pipeline.

Host modules mirror the reference's stage functions; all arithmetic of the hot
path runs in mirge_b200/csrc/libmirge_b200.so (CUDA, sm_100a) behind the C ABI
of include/mirge_b200.h.  There is no CPU fallback.
"""
__version__ = '2.0'
