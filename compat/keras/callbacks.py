from voicemap_b200.keras_compat import Callback, CSVLogger, ModelCheckpoint, ReduceLROnPlateau  # noqa: F401
