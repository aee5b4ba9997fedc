"""Importable alias of the `audio-visual-llm_b200/` package directory (a hyphen cannot be imported).

`import audio_visual_llm_b200` executes `audio-visual-llm_b200/__init__.py` under this module name, so
sub-modules resolve as `audio_visual_llm_b200.<name>` and all source lives in one place.
"""
from pathlib import Path as _Path

_real = _Path(__file__).resolve().parent.parent / "audio-visual-llm_b200"
__path__ = [str(_real)]
exec(compile((_real / "__init__.py").read_text(), str(_real / "__init__.py"), "exec"))
