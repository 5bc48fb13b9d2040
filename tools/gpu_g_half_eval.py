"""Parity with fp32 pre-activations (BLSTMEncoder.g_half = False, the behaviour before the fp16-G default): the whole-path config-shape
tests, the recurrence tests and the model tests with the option forced off (twin.G_HALF off as well); the report of observed errors
goes to gpurun_out/r02_parity_config_shapes_g32.json (compare with profiles/r02_parity_config_shapes.json, the default)."""
import sys

import pytest

sys.path.insert(0, ".")
from lstm_ctc_b200 import blstm

_init = blstm.BLSTMEncoder.__init__


def init(self, *a, **k):
    _init(self, *a, **k)
    self.g_half = False


blstm.BLSTMEncoder.__init__ = init
from oracle import twin  # noqa: E402

twin.G_HALF = False
rc = pytest.main(["-q", "-m", "gpu", "tests/test_config_shapes_gpu.py", "tests/test_blstm_gpu.py", "tests/test_model_gpu.py",
                  "tests/test_full_size_gpu.py"] + sys.argv[1:])
import os  # noqa: E402

if os.path.exists("gpurun_out/r02_parity_config_shapes.json"):
    os.replace("gpurun_out/r02_parity_config_shapes.json", "gpurun_out/r02_parity_config_shapes_g32.json")
sys.exit(rc)
