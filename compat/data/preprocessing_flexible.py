"""Shim for R/data/preprocessing_flexible.py."""
from audio_classification_icbhi_b200.preprocessing_flexible import FlexibleAudioPreprocessor  # noqa: F401
