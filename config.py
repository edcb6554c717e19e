"""config.py of the reference (config.py:1-5), re-exported for scripts that do ``from config import PATH``."""
from voicemap_b200.config import LIBRISPEECH_SAMPLING_RATE, PATH  # noqa: F401
