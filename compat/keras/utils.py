from voicemap_b200.keras_compat import Sequence, plot_model, to_categorical  # noqa: F401
