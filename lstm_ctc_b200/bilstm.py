"""`create_logits_blstm(nnet_input, sequence_length, nnet_config)` -- functional entry point with the
reference's signature and return triple (/root/reference/nnet/bilstm.py:25,273):
    logits [B,T,V], encoder [B, 2(H+P)], reg_loss list[(tensor, weight)]
Weights come from `model` when given, else from the default model of the current graph, else fresh
TF-default random initialisation (the reference creates variables at first call)."""
from .graph import _default_model
from .model import AcousticModel


def create_logits_blstm(nnet_input, sequence_length, nnet_config, model=None):
    m = model or _default_model[0]
    if m is None:
        m = AcousticModel(nnet_config)
        _default_model[0] = m
    training = nnet_config.get("is_training")
    training = True if training is None else bool(training)
    logits = m.forward_logits(nnet_input, sequence_length, training=training)
    encoder = m.enc.encoder_state()
    # reg_loss (bilstm.py:254-273): [(label-smoothing loss tensor, its weight)] when uniform_label_sm / prior_label_sm > 0 -- the
    # reference appends the UNWEIGHTED loss and graph.py:120-133 multiplies by the weight; empty otherwise
    reg_loss = []
    if m.sm_weight > 0:
        reg_loss.append((m.label_smoothing(logits) / m.sm_weight, m.sm_weight))
    return logits, encoder, reg_loss
