"""Shim for R/src/data/dataset.py."""
from audio_classification_icbhi_b200.dataset import GpuCollate, ICBHIDataset  # noqa: F401
