"""ldpcb200 -- B200-native belief-propagation decoder behind the LDPCDecoders.jl BP API.

Only the hot path of /root/reference/src/decoders/belief_propagation.jl is implemented:
csrc/ holds the CUDA kernels and the C ABI (include/ldpcb200.h); decoder.py mirrors the
reference's decoder interface on top of that ABI; codes.py builds the benchmark matrices.
"""
from . import _lib, codes, sharding                                              # noqa: F401
from .decoder import (BeliefPropagationDecoder, BeliefPropagationOSDDecoder, BPOTSDecoder,     # noqa: F401
                      BeliefPropagationScratchSpace,
                      decode_b, batchdecode_b, reset_b, ler_curve)

__all__ = ["BeliefPropagationDecoder", "BeliefPropagationOSDDecoder", "BPOTSDecoder", "BeliefPropagationScratchSpace", "decode_b", "batchdecode_b",
           "reset_b", "ler_curve", "codes"]
