from voicemap_b200.models import *  # noqa: F401,F403
from voicemap_b200.models import build_siamese_net, get_baseline_convolutional_encoder, load_model  # noqa: F401
