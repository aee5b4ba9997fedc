"""B200-native clip_whisper multimodal connector (drop-in for the reference's
src/clip_whisper/models connector path)."""
from . import _lib  # noqa: F401

__all__ = ["_lib"]
