from voicemap_b200.librispeech import *  # noqa: F401,F403
from voicemap_b200.librispeech import LibriSpeechDataset, label_to_sex, sex_to_label  # noqa: F401
