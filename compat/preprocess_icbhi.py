"""Shim for R/preprocess_icbhi.py (same CLI flags)."""
from audio_classification_icbhi_b200.segmenter import ICBHISegmenter, main  # noqa: F401

if __name__ == "__main__":
    main()
