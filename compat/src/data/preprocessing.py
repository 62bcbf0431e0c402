"""Shim for R/src/data/preprocessing.py."""
from audio_classification_icbhi_b200.preprocessing import AudioPreprocessor  # noqa: F401
