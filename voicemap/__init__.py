"""Drop-in namespace: ``from voicemap.models import ...`` / ``voicemap.utils`` / ``voicemap.librispeech`` resolve to
the B200 implementation, so the reference's experiment scripts keep their import lines."""
