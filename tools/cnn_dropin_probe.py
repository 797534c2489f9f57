"""Worst errors of the real-CNN drop-in against the reference golden (tests/golden_cnn), eager and graph replay:
the numbers behind the tolerances in tests/test_gpu_dropin.py."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.test_gpu_dropin import cnn_dropin_errors
for mode in ("eager", "graph"):
    print(json.dumps(dict(mode=mode, **cnn_dropin_errors(mode))), flush=True)
