"""Epoch loops -- same behaviour as /root/reference/nnet/funcs.py:23-152 (`train`, `validate`):
run the graph until the pipeline raises OutOfRangeError, keep a label-token-weighted running mean of the
pure CTC loss (`eval_loss`/`size`, funcs.py:42-54), log the lines the shell drivers grep
(`tr_loss = %f`, `cv_loss = %f`, `cv_eval = %f` -- scripts/train.sh:145,156-157) with TF's
`INFO:tensorflow:` prefix on stderr, and exit(1) on a NaN loss (funcs.py:64,76-81)."""
import math
import sys

from .pipeline import OutOfRangeError


def _info(msg):
    sys.stderr.write("INFO:tensorflow:%s\n" % msg)
    sys.stderr.flush()


def _fatal(msg):
    sys.stderr.write("FATAL:tensorflow:%s\n" % msg)
    sys.stderr.flush()


def _loop(sess, nodes, tag, evaluate, report_interval):
    step, processed, loss, acc = 0, 0, 0.0, 0.0
    try:
        while True:
            values = sess.run(nodes)
            n_tok = values["size"]
            if n_tok > 0:
                processed += n_tok
                loss += (values["eval_loss"] / n_tok - loss) * n_tok / processed
                if evaluate:
                    acc += (values["eval"] / n_tok - acc) * n_tok / processed
            step += 1
            if report_interval and step % report_interval == 0:
                line = "step = %d, batch_size = %d, loss = %f" % (step, n_tok, loss)
                if evaluate:
                    line += ", eval = %f" % acc
                _info(line)
            if math.isnan(loss):
                raise ValueError
    except OutOfRangeError:
        _info("done")
    except KeyboardInterrupt:
        _fatal("interrupted by user")
        sys.exit(1)
    except ValueError:
        _info("%s_loss = %f" % (tag, loss))
        _fatal("nan loss detected")
        sys.exit(1)
    _info("%s_loss = %f" % (tag, loss))
    if evaluate and tag == "cv":
        _info("cv_eval = %f" % acc)
    return True


def train(sess, graph, evaluate=False, report_interval=None):
    nodes = {k: graph[k] for k in ("size", "train", "summary", "loss", "eval_loss", "sequence_length")}
    if evaluate:
        nodes["eval"] = graph["eval"]
    return _loop(sess, nodes, "tr", evaluate, report_interval)


def validate(sess, graph, evaluate=False, report_interval=None):
    nodes = {k: graph[k] for k in ("size", "loss", "eval_loss")}
    if evaluate:
        nodes["eval"] = graph["eval"]
    return _loop(sess, nodes, "cv", evaluate, report_interval)
