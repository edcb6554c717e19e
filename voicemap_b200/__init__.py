"""voicemap_b200: B200-native implementation of voicemap's 1D-conv speaker-embedding hot path.

Host side (this package) mirrors the reference's Python interface (voicemap.models builders, voicemap.utils,
voicemap.librispeech); all network arithmetic runs in hand-written sm_100a CUDA behind the C ABI declared in
``include/voicemap_b200.h`` (``libvoicemap_b200.so``, built by ``python -m voicemap_b200.build``).
"""
__version__ = "0.1.0"
