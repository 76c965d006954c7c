"""lesseq_b200 -- B200-native (sm_100a) count + solve hot path of LESSeq behind a C ABI.

Only what the path needs lives here: ``csrc/`` (CUDA kernels, C ABI, host front end),
``api`` (ctypes mirror of include/lesseq_b200.h), ``synth`` (BASELINE-shaped synthetic inputs),
``dist`` (gene-range sharding over torch.distributed) and ``build`` (in-tree nvcc build).
Importing the package does not load the CUDA library; ``lesseq_b200.api`` does, and fails
loudly if it has not been built -- there is no CPU fallback.
"""
__version__ = "0.1"
