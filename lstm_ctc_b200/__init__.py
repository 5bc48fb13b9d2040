"""lstm_ctc_b200 -- B200-native (sm_100a) BiLSTM / mixture-output / CTC training hot path behind
the reference's `nnet` Python API (/root/reference/nnet/__init__.py:15-26)."""
__version__ = "0.1.0"
