"""config.py of the reference (config.py:1-5): project PATH and the LibriSpeech sampling rate."""
import os

PATH = os.environ.get("VOICEMAP_PATH", os.path.dirname(os.path.dirname(os.path.realpath(__file__))))

LIBRISPEECH_SAMPLING_RATE = 16000
