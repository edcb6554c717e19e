"""`keras` names for the reference's scripts, served by voicemap_b200 (NOT Keras).

The reference's experiment scripts import a dozen Keras symbols next to `voicemap.*`
(`keras.optimizers.Adam`, `keras.callbacks.{CSVLogger, ModelCheckpoint, ReduceLROnPlateau, Callback}`,
`keras.utils.{plot_model, to_categorical, Sequence}`, `keras.layers.Dense`, `keras.models.{load_model, clone_model}`).
This directory is deliberately not on the import path by default: put `<repo>/compat` on PYTHONPATH (after the repo
root) to run those scripts against the B200 implementation without editing their import lines; leave it off and a
real Keras installation, if any, stays visible.  See INTEGRATION.md section 1.
"""
from . import backend, callbacks, layers, models, optimizers, utils  # noqa: F401

__version__ = "2.2.2-voicemap_b200-shim"
