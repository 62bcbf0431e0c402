"""Shim for the reference's `src.data` package (R/src/data/__init__.py)."""
from audio_classification_icbhi_b200.dataset import ICBHIDataset
from audio_classification_icbhi_b200.preprocessing import AudioPreprocessor

__all__ = ["ICBHIDataset", "AudioPreprocessor"]
