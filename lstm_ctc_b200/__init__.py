"""lstm_ctc_b200 -- B200-native (sm_100a) BiLSTM / mixture-output / CTC training hot path behind the
reference's `nnet` Python API (/root/reference/nnet/__init__.py:15-26: the same eleven names, minus the
broken beam-search decoding graph).  nnet_type 'blstm' (nnet/bilstm.py) and 'lstm' (functional core of nnet/lstm.py)."""
__version__ = "0.1.0"

from .config import parse_config  # noqa: F401
from .class_prior import get_class_prior  # noqa: F401
from .funcs import train, validate  # noqa: F401
from .graph import (  # noqa: F401
    Saver, Session, create_graph_for_inference, create_graph_for_training_ctc, create_graph_for_validation_ctc,
    get_create_logits, get_optimizer, global_variables_initializer, local_variables_initializer, trainable_variables,
)
from .tfrecord import dataset_from_tfrecords, write_tfrecord  # noqa: F401
from .kaldi_io import BaseFloatMatrixWriter  # noqa: F401
from .pipeline import (  # noqa: F401
    OutOfRangeError, SyntheticDataset, create_pipeline_sequence_batch, create_pipeline_sequential,
)
