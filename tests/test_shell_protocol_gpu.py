"""The epoch driver's side of the contract (/root/reference/scripts/train.sh:86-236, decode_ctc_lat.sh:150-165), exercised
the way the shell does it: every bin/nnet-*.py is a fresh PROCESS, stderr goes to `nnet.$iter.{tr,cv}.log`, the losses are
pulled out with `grep "^INFO:tensorflow:tr_loss" | awk '{print $NF}'`, checkpoints travel as PREFIXES (`$dir/nnet.$iter`),
the best one is accepted / rejected on cv_loss, its basename lands in `$dir/final.nnet`, and the decoder side runs
`nnet-forward.py ... $dir/$(cat $dir/final.nnet) ark:-`-style on it.  The bash below is our own restatement of that protocol
("the shell-driver protocol never exercised")."""
import os
import subprocess
import sys

import numpy as np
import pytest

from lstm_ctc_b200 import kaldi_io, tf_bundle, tfrecord as tfr

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

DRIVER = r"""
set -e
dir=$1; scp=$2; cfg=$3; py=$4; bin=$5
lr=0.004
# --- initialise + cross-validate the initial network (train.sh:86-103) ---
$py $bin/nnet-init.py --objective=ctc --evaluate=true --batch-size 4 $scp $cfg $dir/nnet.0 2> $dir/nnet.0.cv.log
cv_loss=$(grep "^INFO:tensorflow:cv_loss" $dir/nnet.0.cv.log | awk '{print $NF}')
cv_eval=$(grep "^INFO:tensorflow:cv_eval" $dir/nnet.0.cv.log | awk '{print $NF}')
(echo "cv_loss $cv_loss"; echo "cv_eval $cv_eval") > $dir/nnet.0.done
best=$dir/nnet.0
cv_best=$(grep "^cv_loss" $dir/nnet.0.done | awk '{print $NF}')
# --- epochs (train.sh:121-160, 178-190) ---
for iter in 1 2 3; do
  out=$dir/nnet.$iter
  $py $bin/nnet-train.py --objective=ctc --learn-rate=$lr --optimizer=adam --seed=$iter --shuffle=true --batch-size 4 \
      --batch-threads 2 --report-interval=2 $scp $cfg $best $out 2> $dir/nnet.$iter.tr.log
  tr_loss=$(grep "^INFO:tensorflow:tr_loss" $dir/nnet.$iter.tr.log | awk '{print $NF}')
  [ "$tr_loss" == "nan" ] && exit 3
  $py $bin/nnet-validate.py --objective=ctc --evaluate=true --batch-size 4 --batch-threads 2 --report-interval=2 \
      $scp $cfg $out 2> $dir/nnet.$iter.cv.log
  cv_loss=$(grep "^INFO:tensorflow:cv_loss" $dir/nnet.$iter.cv.log | awk '{print $NF}')
  cv_eval=$(grep "^INFO:tensorflow:cv_eval" $dir/nnet.$iter.cv.log | awk '{print $NF}')
  (echo "tr_loss $tr_loss"; echo "cv_loss $cv_loss"; echo "cv_eval $cv_eval") > $dir/nnet.$iter.done
  echo "nnet.$iter" > $dir/final.nnet
  if [ 1 == $(awk "BEGIN{print($cv_loss < $cv_best ? 1:0);}") ]; then
    best=$out; cv_best=$cv_loss; echo "accepted nnet.$iter $cv_loss"
  else
    echo "rejected nnet.$iter $cv_loss"
    lr=$(awk "BEGIN{print($lr*0.5)}")
  fi
done
echo "$(basename $best)" > $dir/final.nnet
# --- decoder side (decode_ctc_lat.sh:150-165): posteriors of the final network, log domain, blank moved to the front ---
$py $bin/nnet-forward.py --apply-log=true --blank-to-front=true --batch-size 8 $scp $cfg $dir/$(cat $dir/final.nnet) \
    ark,scp:$dir/post.ark,$dir/post.scp 2> $dir/forward.log
"""


def test_epoch_driver_protocol(cuda_dev, tmp_path):
    tmp = str(tmp_path)
    D, V = 8, 11
    rng = np.random.RandomState(5)
    scp = os.path.join(tmp, "feats.scp")
    lens = sorted(rng.randint(30, 61, size=10))
    with open(scp, "w") as fh:
        for i, n in enumerate(lens):
            p = os.path.join(tmp, "utt%03d.tfrecords" % i)
            tfr.write_tfrecord(p, rng.randn(n, D).astype(np.float32), rng.randint(0, V - 1, size=max(1, n // 12)))
            fh.write("utt%03d %d %d 1 %s\n" % (i, n, D, p))
    cfg = os.path.join(tmp, "nnet.config")
    with open(cfg, "w") as fh:
        fh.write("nnet_type blstm\ninput_dim %d\nleft_context 1\nright_context 1\nsubsample 3\nnum_layers 2\nnum_neurons 64\n"
                 "num_projects 64\nnum_targets %d\nuse_peepholes true\nnum_experts 4\nmoe_temp 10.0\ndropout_rate 0.9\n" % (D, V))
    drv = os.path.join(tmp, "driver.sh")
    with open(drv, "w") as fh:
        fh.write(DRIVER)
    r = subprocess.run(["bash", drv, tmp, scp, cfg, sys.executable, os.path.join(ROOT, "bin")], capture_output=True, text=True,
                       timeout=900, env=dict(os.environ, PYTHONPATH=ROOT))
    logs = "".join(open(os.path.join(tmp, f)).read()[-2000:] for f in sorted(os.listdir(tmp)) if f.endswith(".log"))
    assert r.returncode == 0, (r.stdout, r.stderr, logs)
    # every .done file carries parseable numbers, as the driver's awk arithmetic needs them
    done = {}
    for it in range(4):
        kv = dict(l.split() for l in open(os.path.join(tmp, "nnet.%d.done" % it)))
        done[it] = {k: float(v) for k, v in kv.items()}
        assert np.isfinite(done[it]["cv_loss"]) and 0.0 <= done[it]["cv_eval"]
    assert all("tr_loss" in done[it] for it in (1, 2, 3))
    # accept / reject decisions follow the cv_loss sequence; the final network is the best accepted one
    best, cvb = 0, done[0]["cv_loss"]
    for it in (1, 2, 3):
        verdict = "accepted" if done[it]["cv_loss"] < cvb else "rejected"
        assert "%s nnet.%d" % (verdict, it) in r.stdout
        if verdict == "accepted":
            best, cvb = it, done[it]["cv_loss"]
    assert best >= 1 and cvb < done[0]["cv_loss"]                         # training on the cv data lowers its loss
    final = open(os.path.join(tmp, "final.nnet")).read().strip()
    assert final == "nnet.%d" % best
    # checkpoints are Saver-V2 bundles under the prefix; each epoch's process restored the previous best one
    for it in range(4):
        assert os.path.exists(os.path.join(tmp, "nnet.%d.index" % it)) and os.path.exists(os.path.join(tmp, "nnet.%d.data-00000-of-00001" % it))
    names = tf_bundle.read_bundle(os.path.join(tmp, final))
    assert "Variable" in names and "Variable_2" in names and any(k.endswith("/kernel") for k in names)
    # the archive the decoder reads: keys in scp order, log posteriors, blank (last class) moved to column 0
    post = kaldi_io.read_float_matrix_ark(os.path.join(tmp, "post.ark"))
    assert [k for k, _ in post] == ["utt%03d" % i for i in range(10)]
    for (k, a), n in zip(post, lens):
        assert a.shape == (n // 3, V) and np.allclose(np.exp(a).sum(1), 1.0, atol=1e-4)
    assert len(open(os.path.join(tmp, "post.scp")).read().split("\n")) >= 10
