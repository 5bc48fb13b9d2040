#!/usr/bin/env python
"""Drop-in for /root/reference/bin/nnet-init.py on the B200 path (see lstm_ctc_b200/cli.py)."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from lstm_ctc_b200.cli import nnet_init  # noqa: E402

if __name__ == "__main__":
    nnet_init()
