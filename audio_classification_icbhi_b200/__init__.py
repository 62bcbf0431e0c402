"""B200-native (sm_100a) log-mel front end for the ICBHI lung-sound classifier.

Drop-in for the preprocessing path of AkZuza/audio-classification-icbhi
(src/data/preprocessing.py, data/preprocessing_flexible.py and their dataset / analyzer
callers).  All arithmetic runs in hand-written CUDA behind a C ABI (include/logmel_b200.h);
there is no CPU fallback.
"""
from .plan import LogMelPlan, make_aug_array, reference_filterbank, reference_window

__all__ = ["LogMelPlan", "make_aug_array", "reference_filterbank", "reference_window"]
__version__ = "0.1.0"
