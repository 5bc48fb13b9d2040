"""Label-count file -> log prior, as /root/reference/nnet/class_prior.py:30-47: counts are read from the
first line (`[ c0 c1 ... ]`), normalised, logged; entries below 1e-10 get -1e10; and the blank, which
the count file keeps at index 0, is rotated to the LAST position (TF's CTC blank = num_classes-1)."""
import numpy as np

PRIOR_CUTOFF = 1e-10


def read_label_counts(path):
    with open(path) as fh:
        first = fh.readline()
    return [float(tok) for tok in first.strip().lstrip("[").rstrip("]").split()]


def get_class_prior(path):
    counts = np.asarray(read_label_counts(path), dtype=np.float32)
    dist = counts / counts.sum()
    with np.errstate(divide="ignore"):
        logp = np.log(dist)
    logp[dist < PRIOR_CUTOFF] = -1e10
    return np.roll(logp, -1)          # index 0 (blank) moves to the end, everything else shifts down
