"""compat/ serves the B200 classes under the reference's own import paths WITHOUT hiding the reference's other
packages: with compat/ first on sys.path, `src.models.cnn` still comes from the reference tree
(R/train_segmented.py:8-13, R/realtime_analyzer_parallel.py:18-20).  Needs /root/reference (build container only)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"

PROBE = r"""
import inspect, json
from src.models.cnn import LightweightCNN
from src.data.dataset_segmented import ICBHISegmentedDataset
from src.data.dataset import ICBHIDataset
from src.data.preprocessing import AudioPreprocessor
from src.data import ICBHIDataset as D2, AudioPreprocessor as P2
from data.preprocessing_flexible import FlexibleAudioPreprocessor
from preprocess_icbhi import ICBHISegmenter
import src.training.trainer_fixed as tf
print(json.dumps({k: inspect.getfile(v) for k, v in dict(cnn=LightweightCNN, seg=ICBHISegmentedDataset, ds=ICBHIDataset,
      pre=AudioPreprocessor, flex=FlexibleAudioPreprocessor, segm=ICBHISegmenter, trainer=tf.Trainer).items()}
      | {"same": D2 is ICBHIDataset and P2 is AudioPreprocessor}))
"""


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "src")), reason="the reference tree is not on this machine")
def test_reference_scripts_resolve_models_from_the_reference_and_data_from_the_b200_package():
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([os.path.join(ROOT, "compat"), ROOT, REF]), PYTHONDONTWRITEBYTECODE="1")
    res = subprocess.run([sys.executable, "-c", PROBE], capture_output=True, text=True, env=env, cwd="/tmp", timeout=300)
    assert res.returncode == 0, res.stderr[-2000:]
    import json
    files = json.loads(res.stdout.strip().splitlines()[-1])
    pkg = os.path.join(ROOT, "audio_classification_icbhi_b200")
    assert files["cnn"].startswith(REF) and files["trainer"].startswith(REF)
    for k in ("seg", "ds", "pre", "flex", "segm"):
        assert files[k].startswith(pkg), (k, files[k])
    assert files["same"] is True


def test_compat_packages_extend_their_path():
    """The shim packages must be path-extending packages, or they shadow the reference's `src` / `data`."""
    for rel in ("compat/src/__init__.py", "compat/data/__init__.py"):
        text = open(os.path.join(ROOT, rel)).read()
        assert "extend_path" in text, rel
