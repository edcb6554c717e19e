from voicemap_b200.keras_compat import clone_model  # noqa: F401
from voicemap_b200.models import load_model  # noqa: F401
