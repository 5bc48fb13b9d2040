"""Command-line drivers of the hot path -- same positional arguments, switches, log lines and exit codes as
/root/reference/bin/nnet-init.py, nnet-train.py, nnet-validate.py and nnet-forward.py, on the B200 kernels.
The shell epoch drivers (scripts/train.sh:130-160) call these once per epoch and grep `tr_loss` / `cv_loss` /
`cv_eval` from stderr, which nnet.train / nnet.validate print in the reference's format.

Extra switch (ours): --device-splice true  applies frame splicing / subsampling on the GPU after the H2D copy instead
of on the host (same values; see lstm_ctc_b200/tfrecord.py)."""
import argparse
import os
import sys

import numpy

import lstm_ctc_b200 as nnet


def str2bool(v):
    if isinstance(v, bool):
        return v
    if v.lower() in ("yes", "true", "t", "y", "1"):
        return True
    if v.lower() in ("no", "false", "f", "n", "0"):
        return False
    raise argparse.ArgumentTypeError("Boolean value expected.")


def _info(msg):
    sys.stderr.write("INFO:tensorflow:%s\n" % msg)


def _fatal(msg):
    sys.stderr.write("FATAL:tensorflow:%s\n" % msg)


def _common(p, train=False, nnet_in=True, nnet_out=False):
    p.add_argument("tfrecords_scp", metavar="<tfrecords.scp>", type=str)
    p.add_argument("nnet_config", metavar="<nnet-config>", type=str)
    if nnet_in:
        p.add_argument("nnet_in", metavar="<nnet-in>", type=str)
    if nnet_out:
        p.add_argument("nnet_out", metavar="<nnet-out>", type=str)
    p.add_argument("--objective", type=str, default="xent")
    p.add_argument("--evaluate", type=str2bool, default="false")
    p.add_argument("--batch-size", type=int, default=256)
    p.add_argument("--batch-threads", type=int, default=8)          # accepted and ignored, like the reference (pipeline.py:27)
    p.add_argument("--num-parallel-calls", type=int, default=32)     # never forwarded by the reference either (nnet-train.py:143)
    p.add_argument("--report-interval", type=int, default=100)
    p.add_argument("--device-splice", type=str2bool, default="false")
    if train:
        p.add_argument("--optimizer", type=str, default="sgd")
        p.add_argument("--learn-rate", type=float, default=0.0001)
        p.add_argument("--seed", type=int, default=777)
        p.add_argument("--shuffle", type=str2bool, default="true")
        p.add_argument("--clip-norm", type=float, default=5.0)


def _build(args, is_training, training_graph):
    nnet_config = nnet.parse_config(args.nnet_config)
    nnet_config["is_training"] = is_training
    nnet_type = nnet_config.get("nnet_type")
    filename, tfrecord, input_dim = nnet.dataset_from_tfrecords(
        tfrecords_scp=args.tfrecords_scp, left_context=nnet_config.get("left_context"),
        right_context=nnet_config.get("right_context"), subsample=nnet_config.get("subsample"),
        shuffle=getattr(args, "shuffle", False) if training_graph else False, seed=getattr(args, "seed", None),
        device_splice=args.device_splice)
    if args.objective != "ctc":
        _fatal("unsupported objective: %s" % args.objective)
        sys.exit(1)
    if nnet.get_create_logits(nnet_type) is None:      # 'blstm', 'lstm'; the reference's 'cudnnlstm' builder does not run as shipped
        _fatal("unsupported nnet_type: %s" % nnet_type)
        sys.exit(1)
    init, pipeline = nnet.create_pipeline_sequence_batch(dataset=tfrecord, input_dim=input_dim, batch_size=args.batch_size,
                                                         batch_threads=args.batch_threads, num_epochs=1)
    if training_graph:
        graph = nnet.create_graph_for_training_ctc(pipeline=pipeline, nnet_config=nnet_config, learn_rate=args.learn_rate,
                                                   clip_norm=args.clip_norm, optimizer=args.optimizer, seed=args.seed)
    else:
        graph = nnet.create_graph_for_validation_ctc(pipeline=pipeline, nnet_config=nnet_config)
    return init, graph


def nnet_init(argv=None):
    """bin/nnet-init.py <tfrecords.scp> <nnet-config> <nnet-out>: initialise the variables, report the loss of the untrained
    net on the given data, save."""
    p = argparse.ArgumentParser(formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    _common(p, nnet_in=False, nnet_out=True)
    args = p.parse_args(argv)
    _info(" ".join(sys.argv))
    try:
        init, graph = _build(args, False, False)
        sess = nnet.Session()
        sess.run(init)
        nnet.validate(sess=sess, graph=graph, evaluate=args.evaluate, report_interval=args.report_interval)
        saver = nnet.Saver(nnet.trainable_variables())
        _info('saving nnet to "%s"' % args.nnet_out)
        saver.save(sess, args.nnet_out)
    except KeyboardInterrupt:
        _fatal("interrupted by user")
        sys.exit(1)


def nnet_train(argv=None):
    """bin/nnet-train.py <tfrecords.scp> <nnet-config> <nnet-in> <nnet-out>: one epoch."""
    p = argparse.ArgumentParser(formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    _common(p, train=True, nnet_out=True)
    args = p.parse_args(argv)
    _info(" ".join(sys.argv))
    try:
        init, graph = _build(args, True, True)
        sess = nnet.Session()
        sess.run(init)
        saver = nnet.Saver(nnet.trainable_variables())
        saver.restore(sess, args.nnet_in)
        nnet.train(sess=sess, graph=graph, evaluate=args.evaluate, report_interval=args.report_interval)
        _info('saving nnet to "%s"' % args.nnet_out)
        saver.save(sess, args.nnet_out)
    except KeyboardInterrupt:
        _fatal("interrupted by user")
        sys.exit(1)


def nnet_validate(argv=None):
    """bin/nnet-validate.py <tfrecords.scp> <nnet-config> <nnet-in>."""
    p = argparse.ArgumentParser(formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    _common(p)
    args = p.parse_args(argv)
    _info(" ".join(sys.argv))
    try:
        init, graph = _build(args, False, False)
        sess = nnet.Session()
        sess.run(init)
        nnet.Saver(nnet.trainable_variables()).restore(sess, args.nnet_in)
        nnet.validate(sess=sess, graph=graph, evaluate=args.evaluate, report_interval=args.report_interval)
    except KeyboardInterrupt:
        _fatal("interrupted by user")
        sys.exit(1)


def nnet_forward(argv=None):
    """bin/nnet-forward.py <tfrecords-scp> <nnet-config> <nnet-in> <nnet-output-wspecifier>: per utterance
    softmax(smooth * logits) [-> log] [- class prior] as a Kaldi float matrix (nnet-forward.py:77-96).

    --batch-size 1 (default) is the reference's mode, one utterance per forward pass; larger values run length-bucketed
    minibatches through the same kernels (forward.py) -- identical matrices, in the same order.  --blank-to-front applies the
    `select-feats $[ntargets-1],0-$[ntargets-2]` reorder of scripts/decode_ctc_lat.sh:161-163 on the device."""
    p = argparse.ArgumentParser(formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    p.add_argument("tfrecords_scp", metavar="<tfrecords-scp>", type=str)
    p.add_argument("nnet_config", metavar="<nnet-config>", type=str)
    p.add_argument("nnet_in", metavar="<nnet-in>", type=str)
    p.add_argument("nnet_output", metavar="<nnet-output-wspecifier>", type=str)
    p.add_argument("--apply-softmax", type=str2bool, default="true")
    p.add_argument("--apply-log", type=str2bool, default="true")
    p.add_argument("--report-interval", type=int, default=100)
    p.add_argument("--class-prior", type=str, default=None)
    p.add_argument("--smooth-factor", type=float, default=1.0)
    p.add_argument("--device-splice", type=str2bool, default="false")
    p.add_argument("--batch-size", type=int, default=1)
    p.add_argument("--blank-to-front", type=str2bool, default="false")
    p.add_argument("--io-threads", type=int, default=4)
    args = p.parse_args(argv)
    _info(" ".join(sys.argv))
    from .forward import BatchedForward, read_scp
    from .model import AcousticModel
    writer = nnet.BaseFloatMatrixWriter(args.nnet_output)
    nnet_config = nnet.parse_config(args.nnet_config)
    nnet_config["is_training"] = False
    if nnet.get_create_logits(nnet_config.get("nnet_type")) is None:
        _fatal("unsupported nnet_type: %s" % nnet_config.get("nnet_type"))
        sys.exit(1)
    if args.apply_log:
        args.apply_softmax = True
    class_prior = None if args.class_prior is None else nnet.get_class_prior(args.class_prior)
    model = AcousticModel(nnet_config)
    from . import graph as _graph
    _graph._default_model[0] = model
    nnet.Saver(model).restore(None, args.nnet_in)
    engine = BatchedForward(model, read_scp(args.tfrecords_scp), left_context=nnet_config.get("left_context"),
                            right_context=nnet_config.get("right_context"), subsample=nnet_config.get("subsample"),
                            device_splice=args.device_splice, batch_size=args.batch_size, smooth_factor=args.smooth_factor,
                            apply_softmax=args.apply_softmax, apply_log=args.apply_log, class_prior=class_prior,
                            blank_to_front=args.blank_to_front, io_threads=args.io_threads)

    def report(n):
        if args.report_interval and n % args.report_interval == 0:
            _info("processed = %d" % n)

    try:
        engine.run(writer.Write, report)
        _info("done")
    except KeyboardInterrupt:
        _fatal("interrupted by user")
        sys.exit(1)
    writer.Close()
    return engine
