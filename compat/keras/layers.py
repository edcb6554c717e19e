from voicemap_b200.keras_compat import Dense  # noqa: F401
