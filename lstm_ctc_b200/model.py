"""BiLSTM + mixture/affine output + CTC on the sm_100a kernels: the executable behind the graph
dicts of graph.py.  Mirrors the maths of create_logits_blstm (/root/reference/nnet/bilstm.py:25-273),
create_moe (/root/reference/nnet/moe.py:29-72) and the loss/optimizer wiring of
/root/reference/nnet/graph.py:51-209.  Every FLOP runs in liblstm_ctc_b200.so."""
import ctypes
import math
from typing import Dict, Optional

import torch

from . import _lib
from .blstm import BF16, F16, F32, BLSTMEncoder, ModelConfig, ParamSpec, _cast16, _ceil, _to_bf16
from .lstm import embed_uni_variables, extract_uni_variables, random_uni_variables, uni_prefix
from .ctc import ctc_loss_grad
from .gemm import gemm, grid_cap

OPT_CODES = {"sgd": 0, "momentum": 1, "adam": 2}


def _glorot(shape, gen):
    fan_in, fan_out = (shape[0], shape[1]) if len(shape) == 2 else (shape[0], shape[0])
    lim = math.sqrt(6.0 / (fan_in + fan_out))
    return (torch.rand(shape, generator=gen, dtype=torch.float32) * 2 - 1) * lim


def _trunc_normal(shape, std, gen):
    x = torch.randn(shape, generator=gen, dtype=torch.float32)
    for _ in range(8):
        bad = x.abs() > 2
        if not bad.any():
            break
        x[bad] = torch.randn(int(bad.sum()), generator=gen, dtype=torch.float32)
    return x.clamp_(-2, 2) * std


def random_tf_variables(cfg: ModelConfig, seed=None) -> Dict[str, torch.Tensor]:
    """TF default initialisers in the reference's variable layout (what nnet-init.py:73 produces):
    glorot-uniform LSTM kernels / peepholes / projection, zero biases, truncated-normal output layer
    (moe.py:33-58, bilstm.py:239-248)."""
    g = torch.Generator()
    if seed is None:
        g.seed()
    else:
        g.manual_seed(int(seed))
    tf = {}
    for i in range(cfg.num_layers):
        din = cfg.din(i)
        for d, c in (("fd", "frnn"), ("bd", "brnn")):
            pre = "%s%d/%s%d" % (d, i, c, i)
            tf[pre + "/kernel"] = _glorot((din + cfg.P, 4 * cfg.H), g)
            tf[pre + "/bias"] = torch.zeros(4 * cfg.H)
            if cfg.use_peepholes:
                for w in ("w_f_diag", "w_i_diag", "w_o_diag"):
                    tf[pre + "/" + w] = _glorot((cfg.H,), g)
            tf[pre + "/projection/kernel"] = _glorot((cfg.H, cfg.P), g)
    od = 2 * cfg.P
    if cfg.K > 0:
        std = 1.0 / math.sqrt(od)
        tf["Variable"] = _trunc_normal((od, cfg.K), std, g)
        tf["Variable_1"] = torch.zeros(cfg.K)
        tf["Variable_2"] = _trunc_normal((od, cfg.K * cfg.V), std, g)
        tf["Variable_3"] = torch.zeros(cfg.K * cfg.V)
    else:
        tf["Variable"] = _trunc_normal((od, cfg.V), 1.0 / math.sqrt(cfg.H), g)
        tf["Variable_1"] = torch.zeros(cfg.V)
    return tf


class AcousticModel:
    """Weights, activations and the forward / loss / backward / update sequence for one GPU."""

    MOS_BWD_ROWS = 32768          # row chunk of the recompute backward (z + dz of a chunk, 115 MB at K*V = 576, stay L2 resident)

    def __init__(self, nnet_config: dict, device=None, seed=None, init=True):
        self.cfg = c = ModelConfig(nnet_config)
        self.device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        # label-smoothing regulariser (bilstm.py:254-269): uniform wins over prior, like the reference's if/elif
        self.sm_weight, self.sm_prior = 0.0, None
        if (c.uniform_label_sm or 0) > 0:
            self.sm_weight = float(c.uniform_label_sm)
        elif (c.prior_label_sm or 0) > 0 and c.prior_label_path:
            from .class_prior import get_class_prior
            self.sm_weight = float(c.prior_label_sm)
            self.sm_prior = torch.from_numpy(get_class_prior(c.prior_label_path)).to(self.device).contiguous()
        self.rows_out = (c.K * c.V + c.K) if c.K > 0 else c.V
        self.ldz = _ceil(self.rows_out, 8)
        # output-layer variables are the unnamed tf.Variable's: L2-decayed, biases included (graph.py:186)
        specs = [ParamSpec("out/Wall", (self.rows_out, 2 * c.P), True), ParamSpec("out/ball", (self.rows_out,), True)]
        self.enc = BLSTMEncoder(c, self.device, extra_specs=specs)
        self.params = self.enc.params
        # mixture-layer backward in three bands of frames, outermost first: the top layer's BPTT visits frames T-1-s and s at scan
        # step s, so it starts on the outer band while the inner ones are still being processed beside it (backward())
        self.top_overlap = True
        self.ostream = torch.cuda.Stream(device=self.device) if torch.cuda.is_available() else None
        self._out16 = None
        self._outbf = None
        self._out_stale = True
        self.opt_state = None
        self.reg_loss = None
        self.global_step = 0
        self._nodecay = self.params.nodecay_ranges()
        if len(self._nodecay) > 24:
            raise _lib.LcbError(-3, "num_layers = %d: lcb_optimizer_step takes at most 24 no-decay (LSTM bias) ranges" % c.num_layers)
        self._sumsq = torch.zeros(1, dtype=torch.float64, device=self.device)
        self._gnorm = torch.zeros(1, dtype=F32, device=self.device)
        if init:
            self.from_tf_dict(random_uni_variables(c, seed) if c.uni else random_tf_variables(c, seed))

    # ------------------------------------------------------------------ variables
    def from_tf_dict(self, tf: Dict[str, torch.Tensor]):
        c = self.cfg
        if c.uni and (uni_prefix(0) + "/kernel") in tf:      # nnet_type 'lstm': uni-directional variables, embedded (lstm.py)
            tf = embed_uni_variables(c, {k: v.detach().cpu() for k, v in tf.items()})
        self.enc.from_tf_dict(tf)
        Wall, ball = self.params.w("out/Wall"), self.params.w("out/ball")
        dev = self.device
        if c.K > 0:
            K, V = c.K, c.V
            Wp, bp = tf["Variable"].to(dev, F32), tf["Variable_1"].to(dev, F32)
            W, b = tf["Variable_2"].to(dev, F32), tf["Variable_3"].to(dev, F32)
            assert W.shape == (2 * c.P, K * V) and Wp.shape == (2 * c.P, K)
            # row v*K+k  <-  reference column k*V+v (moe.py:60 reshapes [.., K, V])
            Wall[:K * V] = W.t().reshape(K, V, 2 * c.P).permute(1, 0, 2).reshape(K * V, 2 * c.P)
            Wall[K * V:] = Wp.t()
            ball[:K * V] = b.view(K, V).t().reshape(K * V)
            ball[K * V:] = bp
        else:
            Wall.copy_(tf["Variable"].to(dev, F32).t())
            ball.copy_(tf["Variable_1"].to(dev, F32))
        self.mark_stale()

    def to_tf_dict(self, grads=False, embedded=False) -> Dict[str, torch.Tensor]:
        """Reference-layout variables (or their gradients).  nnet_type 'lstm': the uni-directional set unless embedded=True."""
        c = self.cfg
        if c.uni and not embedded:
            return extract_uni_variables(c, self.to_tf_dict(grads, embedded=True))
        out = self.enc.to_tf_dict(grads)
        get = self.params.g if grads else self.params.w
        Wall, ball = get("out/Wall"), get("out/ball")
        if c.K > 0:
            K, V = c.K, c.V
            out["Variable"] = Wall[K * V:].t().clone()
            out["Variable_1"] = ball[K * V:].clone()
            out["Variable_2"] = Wall[:K * V].reshape(V, K, 2 * c.P).permute(1, 0, 2).reshape(K * V, 2 * c.P).t().clone()
            out["Variable_3"] = ball[:K * V].view(V, K).t().reshape(K * V).clone()
        else:
            out["Variable"] = Wall.t().clone()
            out["Variable_1"] = ball.clone()
        return out

    def mark_stale(self):
        self.enc.mark_stale()
        self._out_stale = True

    def _refresh(self):
        if self._out_stale:
            self._out16 = _cast16(self.params.w("out/Wall"), F16, self._out16)
            self._outbf = _cast16(self.params.w("out/Wall"), BF16, self._outbf)
            self._out_stale = False

    def num_params(self):
        return sum(int(v.numel()) for v in self.to_tf_dict().values())

    # ------------------------------------------------------------------ forward
    def _out_ws(self, T, B):
        """Output-layer buffers of a (T, B) minibatch: views of the encoder's grow-only arena (one allocation for the largest
        batch seen, whatever sequence of T the data brings)."""
        c, a = self.cfg, self.enc._arena
        N = T * B
        R = min(self.MOS_BWD_ROWS, N)
        ws = {"logits": a.flat("logits", N * c.V, F32).view(B, T, c.V),
              "Xbf": a.rows("outXbf", N, 2 * c.P, BF16),
              "dXtop": a.rows("dXtop", N, 2 * c.P, BF16),
              "dZ": a.rows("dZ", R if c.K > 0 else N, self.ldz, BF16)}
        if c.K > 0:
            ws["Z"] = a.rows("Z", R, self.ldz, F32)
        return ws

    def forward_logits(self, nnet_input, seq_len, training=True, seq_len_host=None):
        """create_logits_blstm: returns logits [B,T,V] f32 (batch-major).  Rows past seq_len hold the
        output layer applied to a zero encoder row (bias-only / MoE(0)), like the reference."""
        L = _lib.lib()
        c = self.cfg
        self._refresh()
        X = self.enc.forward(nnet_input, seq_len, training, seq_len_host=seq_len_host)
        B, T = nnet_input.shape[0], nnet_input.shape[1]
        ws = self._out_ws(T, B)
        keep = c.keep_prob if training else 1.0
        self._out_seed = self.enc.dropout_seed(255)
        self._xbf_done = None
        self._top_bf = None
        if training:
            hbf = self.enc._workspace(T, B, True)["Hbf"][-1]
            if hbf is not None and self.enc.bf16_twins:                    # the top layer's output projection already wrote the bf16 twin (lcb_gemm16_twin)
                self._top_bf = hbf
        if training and self._top_bf is None and self.enc.wstream is not None:
            # bf16 copy of the encoder output for the output layer's weight gradient: made on the side stream beside the output
            # layer and the CTC sweep (both leave most of the chip idle) instead of at the head of backward()
            main = torch.cuda.current_stream()
            self.enc.wstream.wait_stream(main)
            with torch.cuda.stream(self.enc.wstream):
                _to_bf16(X, ws["Xbf"])
                self._xbf_done = torch.cuda.Event()
                self._xbf_done.record(self.enc.wstream)
        _lib.check(L.lcb_output_fwd(_lib.ptr(X), X.stride(0), _lib.ptr(self._out16), _lib.ptr(self.params.w("out/ball")),
                                    _lib.ptr(ws["logits"]), T, B, 2 * c.P, c.V, c.K, c.tau, keep, self._out_seed,
                                    _lib.stream_ptr()), "lcb_output_fwd")
        self._top = (X, T, B)
        return ws["logits"]

    # ------------------------------------------------------------------ backward
    def backward(self, dlogits, bucket_ready=None):
        """dlogits [B,T,V] f32 -> all parameter gradients (accumulated into params.gflat)."""
        L = _lib.lib()
        c = self.cfg
        X16, T, B = self._top
        N = T * B
        ws = self._out_ws(T, B)
        st = _lib.stream_ptr()
        if getattr(self, "_top_bf", None) is not None:
            Xbf = self._top_bf
        elif getattr(self, "_xbf_done", None) is not None:
            torch.cuda.current_stream().wait_event(self._xbf_done)
            self._xbf_done = None
            Xbf = ws["Xbf"]
        else:
            Xbf = _to_bf16(X16, ws["Xbf"])
        gW, gb = self.params.g("out/Wall"), self.params.g("out/ball")
        dXtop = ws["dXtop"]
        ro = self.rows_out
        # the top BiLSTM layer's output dropout acts on d loss / d encoder output: fused into the GEMMs that produce it
        top_drop = (c.keep_prob, self.enc.dropout_seed(c.num_layers - 1)) if c.keep_prob < 1.0 else None
        if c.K == 0:
            dZ = ws["dZ"]
            _lib.check(L.lcb_pack_dlogits(_lib.ptr(dlogits), _lib.ptr(dZ), T, B, c.V, self.ldz, st), "lcb_pack_dlogits")
            gemm(dZ[:, :ro], self._outbf, 0, 1, out=dXtop, dropout=top_drop and top_drop + (0,))   # dX = dZ * W^T (+ top layer's mask)
            gemm(dZ[:, :ro], Xbf, 1, 1, out=gW)                                   # dW^T = dZ^T * X
            _lib.check(L.lcb_colsum(_lib.ptr(dZ), 1, N, ro, self.ldz, _lib.ptr(gb), st), "lcb_colsum")
        else:
            R = ws["Z"].shape[0]
            first = [True]

            def rows(n0, n1):
                """mixture backward of rows [n0, n1): recompute z, dz, dX (+ the top layer's mask), += dW, += db"""
                for m0 in range(n0, n1, R):
                    r = min(R, n1 - m0)
                    Z, dZ = ws["Z"][:r], ws["dZ"][:r]
                    gemm(X16[m0:m0 + r], self._out16, 0, 0, out=Z[:, :ro], bias=self.params.w("out/ball"))   # recompute z
                    _lib.check(L.lcb_mos_bwd_dz(_lib.ptr(Z), _lib.ptr(dlogits), _lib.ptr(dZ), m0, r, self.ldz, T, B, c.V, c.K,
                                                c.tau, c.keep_prob, self._out_seed, _lib.stream_ptr()), "lcb_mos_bwd_dz")
                    gemm(dZ[:, :ro], self._outbf, 0, 1, out=dXtop[m0:m0 + r], dropout=top_drop and top_drop + (m0 * 2 * c.P,))
                    gemm(dZ[:, :ro], Xbf[m0:m0 + r], 1, 1, out=gW, accumulate=not first[0])
                    first[0] = False
                    _lib.check(L.lcb_colsum(_lib.ptr(dZ), 1, r, ro, self.ldz, _lib.ptr(gb), _lib.stream_ptr()), "lcb_colsum")

            enc = self.enc
            s1, s2 = T // 6, T // 3
            if (self.top_overlap and self.ostream is not None and enc.overlap_wgrad and s1 >= 16
                    and L.lcb_lstm_rec_bwd_can_split(c.Hp) and not c.uni):
                # Bands of frames, outermost first: [0,s1) + [T-s1,T) on this stream (whole chip), then [s1,s2) + [T-s2,T-s1) and the
                # middle [s2,T-s2) on a side stream, capped to the SMs the BPTT clusters leave idle; the encoder launches the top
                # layer's BPTT over scan steps [0,s1) as soon as the first band is done, [s1,s2) after the second, [s2,T) after the
                # third (lcb_lstm_rec_bwd_range: bit-identical to one launch), so two thirds of this layer's backward hide behind it.
                main = torch.cuda.current_stream()
                rows(0, s1 * B)
                rows((T - s1) * B, N)
                ev1 = torch.cuda.Event(); ev1.record(main)
                with torch.cuda.stream(self.ostream), grid_cap(enc.idle_sms(B, 1)):
                    self.ostream.wait_event(ev1)
                    rows(s1 * B, s2 * B)
                    rows((T - s2) * B, (T - s1) * B)
                    ev2 = torch.cuda.Event(); ev2.record(self.ostream)
                    rows(s2 * B, (T - s2) * B)
                    ev3 = torch.cuda.Event(); ev3.record(self.ostream)
                    if bucket_ready is not None:
                        bucket_ready(["out/Wall", "out/ball"])
                enc.backward(dXtop, bucket_ready, top_dropped=top_drop is not None, top_ready=[(s1, ev1), (s2, ev2), (T, ev3)])
                return
            rows(0, N)
        if bucket_ready is not None:
            bucket_ready(["out/Wall", "out/ball"])
        self.enc.backward(dXtop, bucket_ready, top_dropped=top_drop is not None)

    # ------------------------------------------------------------------ loss
    def ctc(self, logits, labels, seq_len, check_labels=True):
        """tf.nn.ctc_loss(..., ignore_longer_outputs_than_inputs=True) + its gradient (graph.py:109-114)."""
        return ctc_loss_grad(logits, labels, seq_len, check_labels=check_labels)

    def loss_and_grad(self, nnet_input, seq_len, labels, bucket_ready=None, check_labels=True, seq_len_host=None, on_loss=None):
        """Forward + CTC + backward for one minibatch.  Returns (sum of per-utt CTC losses as a device
        scalar, per-utt losses).  Gradients are left in params.gflat (un-clipped, no L2 yet).
        on_loss(loss_sum, reg_loss): called once the loss kernels are enqueued and BEFORE the backward pass is, so a caller can
        start reading the step's loss back while the rest of the step still runs (graph.py: Session.run)."""
        self.params.gflat.zero_()
        logits = self.forward_logits(nnet_input, seq_len, training=True, seq_len_host=seq_len_host)
        loss, dlogits = self.ctc(logits, labels, seq_len, check_labels)
        self.reg_loss = self.label_smoothing(logits, dlogits)
        loss_sum = loss.sum()
        if on_loss is not None:
            on_loss(loss_sum, self.reg_loss)
        self.backward(dlogits, bucket_ready)
        return loss_sum, loss

    def label_smoothing(self, logits, dlogits=None):
        """reg_loss of create_logits_blstm (bilstm.py:254-269), added to the training loss by graph.py:120-133.
        Returns a device scalar (or None when the weight is 0); adds its gradient into dlogits."""
        if self.sm_weight <= 0:
            return None
        out = torch.zeros(1, dtype=F32, device=self.device)
        rows = logits.numel() // logits.shape[-1]
        _lib.check(_lib.lib().lcb_label_smooth(_lib.ptr(logits), _lib.ptr(dlogits), rows, logits.shape[-1], self.sm_weight,
                                               _lib.ptr(self.sm_prior), _lib.ptr(out), _lib.stream_ptr()), "lcb_label_smooth")
        return out

    # ------------------------------------------------------------------ update
    def optimizer_step(self, optimizer, learn_rate, clip_norm=5.0, l2_decay_weight=1e-5, momentum=0.9,
                       beta1=0.9, beta2=0.999, eps=1e-8):
        """L2 + clip_by_global_norm + apply_gradients (graph.py:183-200) on the flat buffers."""
        L = _lib.lib()
        opt = OPT_CODES[optimizer]
        ps = self.params
        if self.opt_state is None or self.opt_state[0] != opt:
            s1 = torch.zeros_like(ps.flat) if opt >= 1 else None
            s2 = torch.zeros_like(ps.flat) if opt == 2 else None
            self.opt_state = (opt, s1, s2)
            self.opt_t = 0
        self.opt_t += 1
        _, s1, s2 = self.opt_state
        nr = len(self._nodecay)
        arr = (ctypes.c_longlong * (2 * max(nr, 1)))()
        for k, (lo, hi) in enumerate(self._nodecay):
            arr[2 * k], arr[2 * k + 1] = lo, hi
        _lib.check(L.lcb_optimizer_step(_lib.ptr(ps.flat), _lib.ptr(ps.gflat), _lib.ptr(s1), _lib.ptr(s2), ps.total, opt,
                                        float(learn_rate), self.opt_t, beta1, beta2, eps, momentum, float(l2_decay_weight),
                                        float(clip_norm), arr, nr, _lib.ptr(self._sumsq), _lib.ptr(self._gnorm),
                                        _lib.stream_ptr()), "lcb_optimizer_step")
        self.global_step += 1
        self.mark_stale()

    def last_grad_norm(self):
        return float(self._gnorm.item())

    # ------------------------------------------------------------------ checkpoint (trainable variables only)
    def state_dict(self):
        """Reference checkpoints hold tf.trainable_variables() only (nnet-train.py:83-84,95): no optimizer
        slots, no global_step."""
        return {k: v.cpu() for k, v in self.to_tf_dict().items()}

    def load_state_dict(self, sd):
        self.from_tf_dict(sd)
