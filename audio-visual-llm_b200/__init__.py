"""B200-native clip_whisper multimodal connector: a drop-in for the connector path of
rishabhjain16/audio-visual-llm's src/clip_whisper (see DESIGN.md / INTEGRATION.md).

Public surface mirrors the reference (src/clip_whisper/models/__init__.py:1-4):
    ClipWhisperModel, ModalityConnector (+ SimpleModalityConnector, create_modality_connector)
plus the fused entry point `fused_connector` and its `FusePlan`.
Importing the package does not need a GPU; calling any op without an sm_100 device raises.
"""
from . import _lib  # noqa: F401
from .connector_ops import FusePlan, fused_connector, linear_project  # noqa: F401
from .modality_connector import (  # noqa: F401
    BaseModalityConnector,
    MLPModalityConnector,
    ModalityConnector,
    SimpleModalityConnector,
    create_modality_connector,
)
from .clip_whisper_model import ClipWhisperModel  # noqa: F401
from .seq_adapt import adapt_mask, adaptive_projection  # noqa: F401

__all__ = [
    "ClipWhisperModel", "ModalityConnector", "SimpleModalityConnector", "MLPModalityConnector", "BaseModalityConnector",
    "create_modality_connector", "FusePlan", "fused_connector", "linear_project", "adaptive_projection",
    "adapt_mask",
]
