"""snp_pipeline_b200 -- B200-native pileup -> consensus -> SNP-matrix -> distance path of the CFSAN SNP Pipeline.

Layout:
  csrc/          CUDA kernels (sm_100a) + the C ABI (include/snpgpu.h) -> libsnpgpu.so
  _lib.py        ctypes binding of that ABI
  pileup.py, call_consensus.py, merge_sites.py, snp_matrix.py, distance.py, utils.py
                 host-side mirror of the reference's modules of the same names (same entry points, argument
                 Namespaces, files and error conventions); the arithmetic goes through libsnpgpu
  cli.py         `cfsan_snp_pipeline`-compatible subcommands for the four hot-path steps
"""
__version__ = "0.1.0"
