from voicemap_b200.keras_compat import Adam  # noqa: F401
