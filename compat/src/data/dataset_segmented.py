"""Shim for R/src/data/dataset_segmented.py."""
from audio_classification_icbhi_b200.dataset import GpuCollate, ICBHISegmentedDataset  # noqa: F401
