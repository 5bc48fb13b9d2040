"""oracle/ -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement (PyTorch-CPU fp64/fp32 + a plain-C CTC) of the reference's acoustic-model hot
path: stacked BiLSTM (/root/reference/nnet/bilstm.py:104-273), mixture output layer
(/root/reference/nnet/moe.py:29-72), CTC loss + gradient (tf.nn.ctc_loss, call site
/root/reference/nnet/graph.py:109-116), L2 / global-norm clip / optimizer
(/root/reference/nnet/graph.py:183-200).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  The product package (lstm_ctc_b200) never does.

PARITY PINNING.  The reference holds no tests, fixtures or golden vectors, and its arithmetic
lives in TensorFlow 1.8 (third-party, py2.7, not installable here).  The oracle is therefore
pinned by (1) the two known-answer vectors of upstream TF's ctc_loss_op_test.py
(tests/golden/ctc_tf_kat.json), (2) agreement with torch.nn.functional.ctc_loss and brute-force
path enumeration, (3) agreement of the BiLSTM restatement (stacked, projected, ragged lengths, final states,
input gradient) with torch.nn.LSTM(bidirectional, proj_size) on packed sequences and of the optimizers with
torch.optim / clip_grad_norm_ (tests/test_oracle_model_cpu.py), and (4) the closed form of one peephole step plus
torch.autograd.gradcheck.  Beyond that: "parity unpinned" against the TF binary itself.
"""
from .ctc import ctc_loss_grad, build_ctc_oracle  # noqa: F401
from .model import (  # noqa: F401
    OracleConfig, init_params, blstm_forward, create_moe, output_layer, ctc_loss_sum,
    training_loss, clip_by_global_norm, adam_step, sgd_step, momentum_step, l2_loss,
    greedy_decode, edit_distance, param_order,
    lstm_param_order, init_lstm_params, lstm_forward, lstm_training_loss,
)
