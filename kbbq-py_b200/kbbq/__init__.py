"""kbbq -- B200-native drop-in for the FASTQ recalibration hot path of adamjorr/kbbq-py.

Same module and function names as the reference for that path (kbbq.recalibrate,
kbbq.compare_reads, kbbq.gatk.applybqsr, kbbq.covariate, kbbq.read, kbbq.main); the arithmetic
runs in hand-written CUDA kernels for sm_100a behind the C ABI in include/kbbq_b200.h.
Unlike the reference's __init__ (kbbq/__init__.py:8-11) nothing heavy is imported eagerly.
"""
__version__ = '0.0.0'
