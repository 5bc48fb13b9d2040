"""Stacked BiLSTM encoder on the sm_100a kernels -- host-side mirror of
`create_logits_blstm` (/root/reference/nnet/bilstm.py:25-273) up to the encoder output.

What runs where (all compute is in liblstm_ctc_b200.so; torch only owns memory and streams):
  lcb_pack_input      [B,T,D] f32 -> time-major bf16                       (pipeline.py:35-61 layout)
  lcb_gemm_bf16       G = X*W_x + b for BOTH directions, all frames         (LSTMCell matmul, hoisted)
  lcb_lstm_rec_fwd    masked recurrence, both directions concurrently       (dynamic_rnn + LSTMCell)
  lcb_gemm_bf16       h = m*W_proj written straight into its half of [N,2P] (num_proj + tf.concat)
and the mirrored sequence for the backward pass (lcb_lstm_rec_bwd + dgrad/wgrad GEMMs).

Parameters live in DEVICE layout (gate columns packed (unit/8)*32 + gate*8 + unit%8, hidden size padded to 64) inside one
flat fp32 buffer; `to_tf_dict` / `from_tf_dict` convert to/from the reference's TF variable names and
shapes (fd{i}/frnn{i}/kernel, .../bias, .../w_{f,i,o}_diag, .../projection/kernel; bilstm.py:125-165).
"""
import math
from typing import Dict, List, Optional

import torch

from . import _lib
from .gemm import gemm, grid_cap

F32 = torch.float32
BF16 = torch.bfloat16      # gradient operands (range-safe without loss scaling)
F16 = torch.float16        # forward operands (|m| < 1, CMVN'd features, small weights: 8x finer than bf16)


def _ceil(x, m):
    return (x + m - 1) // m * m


def _kernels_serialised():
    """True when the process runs under a tool that executes one kernel at a time (Nsight Compute / Systems and
    compute-sanitizer inject through CUDA_INJECTION64_PATH; CUDA_LAUNCH_BLOCKING=1 serialises launches): a recurrence launch that
    waits in-kernel for GEMMs of another stream (BLSTMEncoder.fwd_flow_control) could never be fed there, so the encoder falls
    back to consecutive range launches -- same results, ~0.7 ms slower per C3 step."""
    import os
    return bool(os.environ.get("CUDA_INJECTION64_PATH")) or os.environ.get("CUDA_LAUNCH_BLOCKING", "0") not in ("", "0")


class ModelConfig:
    """The nnet_config keys consumed by create_logits_blstm (bilstm.py:39-99)."""

    def __init__(self, nnet_config: dict):
        c = nnet_config
        ctx = 1 + int(c.get("left_context") or 0) + int(c.get("right_context") or 0)
        self.input_dim = int(c["input_dim"]) * ctx                       # bilstm.py:48
        self.num_layers = int(c["num_layers"])
        self.H = int(c["num_neurons"])
        if c.get("num_projects") is None:
            raise TypeError("num_projects is required (2 * None at bilstm.py:199)")
        self.P = int(c["num_projects"])
        self.V = int(c["num_targets"])
        self.use_peepholes = bool(c.get("use_peepholes") or False)
        self.K = int(c.get("num_experts") or 0)
        self.tau = float(c["moe_temp"]) if c.get("moe_temp") is not None else 10.0
        if c.get("dropout_rate") is None:
            raise TypeError("dropout_rate (a keep-probability) is required ('%f' % None at bilstm.py:79)")
        self.keep_prob = float(c["dropout_rate"])
        is_training = c.get("is_training")
        self.is_training = True if is_training is None else bool(is_training)
        if not self.is_training:
            self.keep_prob = 1.0                                          # bilstm.py:98-99
        self.uniform_label_sm = c.get("uniform_label_sm")
        self.prior_label_sm = c.get("prior_label_sm")
        self.prior_label_path = c.get("prior_label_path")
        self.forget_bias = 5.0                                            # bilstm.py:133,154
        # nnet_type 'lstm': the uni-directional residual stack of create_logits_lstm (nnet/lstm.py:125-368) -- run on the same
        # kernels with the backward cells and every weight that reads their output pinned at zero (see lstm.py of this package)
        self.uni = (c.get("nnet_type") == "lstm")
        if self.uni:
            self.use_peepholes = True                                     # lstm.py:241,252 (hard-coded)
            self.forget_bias = 1.0                                        # LSTMCell default (lstm.py:238-244 passes none)
        # device layout
        self.Hp = _ceil(self.H, 64)
        if self.Hp > 512:
            raise _lib.LcbError(-3, "num_neurons > 512 does not fit the cluster-resident recurrence")
        if self.P % 8:
            raise _lib.LcbError(-3, "num_projects must be a multiple of 8")
        self.Dp0 = _ceil(self.input_dim, 8)
        self.residual0 = (not self.uni) and (self.input_dim == 2 * self.P)   # bilstm.py:199-200

    def uni_residual(self, i):
        """ResidualWrapper on every cell except layer 0 when input_dim != num_projects (lstm.py:236-260)."""
        return self.uni and (i > 0 or self.input_dim == self.P)

    def din(self, i):
        return self.input_dim if i == 0 else 2 * self.P

    def dinp(self, i):
        return self.Dp0 if i == 0 else 2 * self.P


class ParamSpec:
    def __init__(self, name, shape, decay):
        self.name, self.shape, self.decay = name, tuple(shape), decay
        self.numel = 1
        for s in shape:
            self.numel *= s
        self.offset = 0


class ParamStore:
    """One flat fp32 buffer for weights and one for gradients (bucket = contiguous slice).
    Order = gradient completion order of the backward pass (output layer first, layer 0 last) so
    data-parallel all-reduce can start on finished slices while BPTT of lower layers still runs."""

    def __init__(self, specs: List[ParamSpec], device):
        off = 0
        for s in specs:
            s.offset = off
            off += _ceil(s.numel, 64)          # 256-byte aligned slices
        self.specs = {s.name: s for s in specs}
        self.order = [s.name for s in specs]
        self.total = off
        self.flat = torch.zeros(off, dtype=F32, device=device)
        self.gflat = torch.zeros(off, dtype=F32, device=device)

    def w(self, name):
        s = self.specs[name]
        return self.flat[s.offset:s.offset + s.numel].view(s.shape)

    def g(self, name):
        s = self.specs[name]
        return self.gflat[s.offset:s.offset + s.numel].view(s.shape)

    def nodecay_ranges(self):
        return [(s.offset, s.offset + s.numel) for s in (self.specs[n] for n in self.order) if not s.decay]


def _cast16(src, dtype, dst=None):
    L = _lib.lib()
    if dst is None:
        dst = torch.empty(src.shape, dtype=dtype, device=src.device)
    assert dst.dtype == dtype
    _lib.check(L.lcb_cast_f32_16(_lib.ptr(src), _lib.ptr(dst), 1 if dtype == BF16 else 2, src.numel(), _lib.stream_ptr()),
               "lcb_cast_f32_16")
    return dst


def _to_bf16(src, dst):
    """fp16 activation -> bf16 scratch copy (same shape, both contiguous) for the wgrad GEMMs."""
    assert src.is_contiguous() and dst.is_contiguous() and src.numel() == dst.numel()
    _lib.check(_lib.lib().lcb_f16_to_bf16(_lib.ptr(src), _lib.ptr(dst), src.numel(), _lib.stream_ptr()), "lcb_f16_to_bf16")
    return dst


def _split_bf16(src, hi=None, lo=None):
    L = _lib.lib()
    if hi is None:
        hi = torch.empty(src.shape, dtype=BF16, device=src.device)
    if lo is None:
        lo = torch.empty(src.shape, dtype=BF16, device=src.device)
    _lib.check(L.lcb_split_f32_bf16(_lib.ptr(src), _lib.ptr(hi), _lib.ptr(lo), src.numel(), _lib.stream_ptr()),
               "lcb_split_f32_bf16")
    return hi, lo


class Arena:
    """Grow-only named device buffers.  A minibatch of (T, B) gets VIEWS of the first T*B rows of each buffer, so a stream of
    batches with different T (every real epoch: `_BatchSource` pads to the batch's own longest utterance) reuses one allocation
    sized for the largest batch seen, instead of one full activation set per distinct T.  Growing frees the old buffer first
    (after a device synchronisation, since side streams may still be reading it) and over-allocates by 1/8 so that a
    length-sorted epoch triggers O(log) regrowths."""

    def __init__(self, device, zero=False):
        self.device, self.zero = device, zero
        self.bufs = {}
        self.generation = 0           # bumped whenever a buffer moves (cached views become stale)

    def flat(self, name, numel, dtype, zero=None):
        buf = self.bufs.get(name)
        if buf is None or buf.numel() < numel or buf.dtype != dtype:
            if buf is not None:
                if buf.is_cuda:
                    torch.cuda.synchronize()
                self.bufs[name] = buf = None
            cap = int(numel + numel // 8) if name in self.bufs else int(numel)
            alloc = torch.zeros if (self.zero if zero is None else zero) else torch.empty
            buf = alloc(max(cap, 1), dtype=dtype, device=self.device)
            self.bufs[name] = buf
            self.generation += 1
        return buf[:numel]

    def rows(self, name, rows, cols, dtype, zero=None):
        return self.flat(name, rows * cols, dtype, zero).view(rows, cols)

    def bytes_allocated(self):
        return sum(b.numel() * b.element_size() for b in self.bufs.values() if b is not None)


class BLSTMEncoder:
    """Weights + forward/backward of the L-layer BiLSTM stack."""

    def __init__(self, cfg: ModelConfig, device, extra_specs: Optional[List[ParamSpec]] = None):
        self.cfg = cfg
        self.device = device
        c = cfg
        specs = list(extra_specs or [])                     # output layer first (its grads finish first)
        for i in reversed(range(c.num_layers)):
            specs += [ParamSpec("L%d/WpT" % i, (2, c.P, c.Hp), True),
                      ParamSpec("L%d/Wh" % i, (8 * c.Hp, c.P), True),
                      ParamSpec("L%d/peep" % i, (2, 3, c.Hp), True),
                      ParamSpec("L%d/bias" % i, (8 * c.Hp,), False),
                      ParamSpec("L%d/Wx" % i, (8 * c.Hp, c.dinp(i)), True)]
        self.params = ParamStore(specs, device)
        self._bf = {}            # bf16 operand copies, refreshed by refresh_operands()
        self._stale = True
        self.wstream = torch.cuda.Stream(device=device) if torch.cuda.is_available() else None   # wgrad side stream
        self.overlap_wgrad = True
        self.rstream = torch.cuda.Stream(device=device) if torch.cuda.is_available() else None   # operand-refresh side stream
        self.use_graphs = True
        self.pstream = torch.cuda.Stream(device=device) if torch.cuda.is_available() else None   # projection of the later frames
        # ---- schedule options (explicit attributes; defaults = what measured best on the C3 step) ----
        # fractions of the scan steps at which the forward recurrence is cut into launches: the pre-activations of the first chunk
        # are projected before the recurrence starts, the later chunks' beside it on a side stream
        self.head_fracs = [0.36]
        # ... when the forward recurrence occupies more clusters (B = 64: 96 SMs) and leaves the projection GEMMs beside it fewer
        # SMs than this: more, smaller chunks, so that each chunk's capped GEMM still fits under the recurrence of the one before
        self.head_fracs_tight = [0.34, 0.62, 0.85]
        self.tight_side_sms = 56
        # One recurrence launch per layer with in-kernel flow control (lcb_lstm_rec_fwd_range_hl, ready_steps): the first
        # flow_fracs[0] of the scan is projected before the launch, the chunks up to each following fraction beside it on the side
        # stream, each followed by lcb_store_i32 on the counter the kernel's prefetch warps wait on.  Clusters are then not
        # re-synchronised at chunk boundaries (which costs ~0.2 ms per layer once 16-utterance groups of different length run at
        # different speeds) and the chunks can be small.  False: consecutive range launches at head_fracs (e.g. under a profiler
        # that serialises kernels -- the waiting launch and the GEMM it waits for must be able to run concurrently).
        self.fwd_flow_control = not _kernels_serialised()
        self.flow_fracs = [0.3, 0.55, 0.8]
        self.overlap_hproj = False     # output projection of finished chunks on the side stream (c3: +-0, c2: 15 % slower)
        # Flow-controlled forward only: the launch publishes its progress (lcb_lstm_rec_fwd_range_pg) and the output projection
        # h = m W_proj of the scan steps before each of these fractions of T runs on the side stream beside the rest of the
        # recurrence (lcb_wait_progress); only the last steps' rows stay on the main stream between two layers.  Empty: the whole
        # projection follows the launch (the schedule before this option; same values bit for bit either way).
        # (Measured and dropped, profiles/r02_fwd_progress_ab.jsonl: also accumulating the next layer's head projection from two
        # K halves, the early one beside the recurrence -- the head GEMMs are bound by their fp32 output, K = 512 costs what
        # K = 1024 does.)
        self.fwd_hproj_fracs = [0.6, 0.85]
        # The hoisted pre-activations G = x W_x + b as fp16 instead of fp32: the projections write, and the recurrence reloads, half
        # the bytes (7.9 -> 3.9 GB per C3 step); the forget bias and the recurrent product are added in fp32 in the accumulator
        # either way.  Costs one fp16 rounding (2^-11 relative) of every pre-activation's input part: against the exact oracle the
        # config-shape errors do not move (logits 1.5e-2 -> 1.6e-2 of scale, gradients 4.3e-2 -> 4.2e-2 in the default-init regime,
        # 6.5e-4 -> 7.7e-4 / 7.3e-3 -> 7.4e-3 in the stable one; profiles/r02_parity_config_shapes_g32.json vs _shapes.json), the
        # C3 step gains 0.7 ms (profiles/r02_g_fp16_ab.jsonl).  False: fp32 G (round-1 / early round-2 behaviour).
        self.g_half = True
        # bf16 twins of the saved activations (wgrad operands) written by their producers -- the recurrence kernel (Mout_bf16) and the
        # output-projection GEMM (lcb_gemm16_twin) -- instead of conversion passes over the fp16 rows in backward()
        self.bf16_twins = True
        self.bwd_split_frac = 0.0      # > 0: BPTT as two launches at this fraction (lcb_lstm_rec_bwd_range; tests)
        # increasing fractions > 0.5 of the scan at which BPTT of layers 1.. is cut into consecutive launches; the rows of dX (and of
        # the next layer's dM) whose dG is final in BOTH directions after a launch are computed beside the next one (backward(),
        # "early rows").  Empty: one launch.
        self.bwd_early_fracs = [0.67, 0.85]
        self.l0_released = True        # layer 0's frame-sum gradients over released frames
        # BPTT as ONE launch per layer that publishes its progress (lcb_lstm_rec_bwd_range_pg); the early-rows GEMMs wait for the
        # release points on their own stream (lcb_wait_progress) instead of following a launch boundary.  False: one launch per
        # range, as in round 1 (bit-identical results either way)
        self.bwd_progress = True
        self.xstream = torch.cuda.Stream(device=device) if torch.cuda.is_available() else None   # early rows of dX / dM
        # share of the SMs a recurrence launch leaves idle that the GEMMs beside it may occupy (forward, BPTT): < 1 trades GEMM
        # time (hidden under the recurrence) for power and L2 headroom of the latency-bound clusters (tools/gpu_schedule_ab.py)
        self.side_sm_scale = [1.0, 1.0]
        self.ndir = 1 if cfg.uni else 2          # nnet_type 'lstm': only direction-0 clusters / column halves run (lstm.py)
        self.num_sms = _lib.lib().lcb_device_sm_count() if torch.cuda.is_available() else 0
        self._arena = Arena(device, zero=cfg.uni)   # uni: the never-written direction-1 halves must read as 0, not as garbage
        self._ws_key = None
        self._ws_views = None
        self.debug_dz = None           # tests: a dict here receives {layer: copy of dz [T*B, 8Hp] bf16} from backward()
        self._refresh_graphs = None
        self._refresh_done = None
        self.seed_base = 777           # reference default --seed (nnet-train.py:141-142); set_dropout_seed() overrides
        self.seed_rank = 0
        self.step_id = 0
        mt = _lib.ctypes.c_int()
        nc = _lib.ctypes.c_int()
        _lib.check(_lib.lib().lcb_lstm_rec_config(c.Hp, _lib.ctypes.byref(mt), _lib.ctypes.byref(nc)), "lcb_lstm_rec_config")
        self.rec_mt, self.rec_nc = mt.value, nc.value

    def set_dropout_seed(self, seed, rank=0):
        """`--seed` of nnet-train.py:141-142 (tf.set_random_seed) seeds the dropout streams as well; data-parallel ranks mix their
        rank in so that shards do not share masks."""
        self.seed_base = int(seed) & 0xffffff
        self.seed_rank = int(rank) & 0xff

    def dropout_seed(self, layer):
        """One mask stream per (seed, rank, training step, layer); layer 255 = mixture output layer."""
        return (((self.seed_base & 0xffffff) << 40) ^ ((self.seed_rank & 0xff) << 32) ^ ((self.step_id & 0xffffff) << 8)
                ^ (layer & 0xff))

    def idle_sms(self, B, which):
        """SMs a forward (which=0) / BPTT (which=1) recurrence launch over B utterances leaves free: the persistent-grid cap of
        the GEMMs that run beside it on a side stream."""
        used = _lib.lib().lcb_lstm_rec_grid(B, self.cfg.Hp, self.ndir, which)
        return max(8, int((self.num_sms - max(used, 0)) * self.side_sm_scale[which]))

    def _dropout(self, x, layer):
        dt = 2 if x.dtype == F16 else 1
        _lib.check(_lib.lib().lcb_dropout16(_lib.ptr(x), dt, x.numel(), self.cfg.keep_prob, self.dropout_seed(layer),
                                            _lib.stream_ptr()), "lcb_dropout16")

    # ------------------------------------------------------------------ TF <-> device layout
    def _tf_names(self, i, d):
        return "%s%d/%s%d" % ("fd" if d == 0 else "bd", i, "frnn" if d == 0 else "brnn", i)

    def from_tf_dict(self, tf: Dict[str, torch.Tensor]):
        """Load reference-layout variables (kernel [Din+P,4H] with gate blocks i,j,f,o; bias [4H];
        w_*_diag [H]; projection/kernel [H,P])."""
        c = self.cfg
        H, Hp, P = c.H, c.Hp, c.P
        ps = self.params
        for i in range(c.num_layers):
            din = c.din(i)
            Wx, Wh, bias, peep, WpT = ps.w("L%d/Wx" % i), ps.w("L%d/Wh" % i), ps.w("L%d/bias" % i), ps.w("L%d/peep" % i), ps.w("L%d/WpT" % i)
            Wx.zero_(); Wh.zero_(); bias.zero_(); peep.zero_(); WpT.zero_()
            for d in range(2):
                pre = self._tf_names(i, d)
                k = tf[pre + "/kernel"].to(device=self.device, dtype=F32)          # [din+P, 4H]
                assert k.shape == (din + P, 4 * H), (k.shape, din, P, H)
                # packed row of (unit u, gate) = d*4Hp + (u//8)*32 + gate*8 + u%8   <-  TF column gate*H + u
                kk = torch.zeros(Hp, 4, din + P, dtype=F32, device=self.device)
                kk[:H] = k.view(din + P, 4, H).permute(2, 1, 0)                      # [u, gate, din+P]
                kk = kk.view(Hp // 8, 8, 4, din + P).permute(0, 2, 1, 3).reshape(4 * Hp, din + P)
                Wx[d * 4 * Hp:(d + 1) * 4 * Hp, :din] = kk[:, :din]
                Wh[d * 4 * Hp:(d + 1) * 4 * Hp] = kk[:, din:]
                bb = torch.zeros(Hp, 4, dtype=F32, device=self.device)
                bb[:H] = tf[pre + "/bias"].to(device=self.device, dtype=F32).view(4, H).t()
                bias[d * 4 * Hp:(d + 1) * 4 * Hp] = bb.view(Hp // 8, 8, 4).permute(0, 2, 1).reshape(4 * Hp)
                if c.use_peepholes:
                    peep[d, 0, :H] = tf[pre + "/w_f_diag"].to(self.device, F32)
                    peep[d, 1, :H] = tf[pre + "/w_i_diag"].to(self.device, F32)
                    peep[d, 2, :H] = tf[pre + "/w_o_diag"].to(self.device, F32)
                WpT[d, :, :H] = tf[pre + "/projection/kernel"].to(self.device, F32).t()  # [P, H]
        self._stale = True

    def to_tf_dict(self, grads=False) -> Dict[str, torch.Tensor]:
        c = self.cfg
        H, Hp, P = c.H, c.Hp, c.P
        get = self.params.g if grads else self.params.w
        out = {}
        for i in range(c.num_layers):
            din = c.din(i)
            Wx, Wh, bias, peep, WpT = get("L%d/Wx" % i), get("L%d/Wh" % i), get("L%d/bias" % i), get("L%d/peep" % i), get("L%d/WpT" % i)
            for d in range(2):
                pre = self._tf_names(i, d)
                def unpack(m2):          # [4Hp, cols] packed rows -> [H, 4, cols]
                    return m2.reshape(Hp // 8, 4, 8, -1).permute(0, 2, 1, 3).reshape(Hp, 4, -1)[:H]
                rx = unpack(Wx[d * 4 * Hp:(d + 1) * 4 * Hp])[:, :, :din]                # [H,4,din]
                rh = unpack(Wh[d * 4 * Hp:(d + 1) * 4 * Hp])                            # [H,4,P]
                k = torch.cat([rx, rh], 2).permute(2, 1, 0).reshape(din + P, 4 * H)
                out[pre + "/kernel"] = k.clone()
                out[pre + "/bias"] = unpack(bias[d * 4 * Hp:(d + 1) * 4 * Hp].unsqueeze(1))[:, :, 0].t().reshape(4 * H).clone()
                if c.use_peepholes:
                    out[pre + "/w_f_diag"] = peep[d, 0, :H].clone()
                    out[pre + "/w_i_diag"] = peep[d, 1, :H].clone()
                    out[pre + "/w_o_diag"] = peep[d, 2, :H].clone()
                out[pre + "/projection/kernel"] = WpT[d, :, :H].t().clone()
        return out

    # ------------------------------------------------------------------ operand refresh (once per update)
    def mark_stale(self):
        self._stale = True

    def _refresh_layer(self, i):
        """16-bit GEMM operands of layer i + its folded recurrent weights W' = W_proj * W_h (fp32-accurate via a 3-term
        split-bf16 product on the tensor cores).  Every destination buffer is allocated once and then rewritten in place, so
        the launch sequence can be replayed from a CUDA graph."""
        c = self.cfg
        ps = self.params
        bf = self._bf
        bf[("Wx16", i)] = _cast16(ps.w("L%d/Wx" % i), F16, bf.get(("Wx16", i)))
        bf[("Wx", i)] = _cast16(ps.w("L%d/Wx" % i), BF16, bf.get(("Wx", i)))
        bf[("WpT16", i)] = _cast16(ps.w("L%d/WpT" % i), F16, bf.get(("WpT16", i)))
        Wh_hi, Wh_lo = _split_bf16(ps.w("L%d/Wh" % i), bf.get(("Wh", i)), bf.get(("Wh_lo", i)))
        Wp_hi, Wp_lo = _split_bf16(ps.w("L%d/WpT" % i), bf.get(("WpT", i)), bf.get(("WpT_lo", i)))
        bf[("Wh", i)], bf[("Wh_lo", i)], bf[("WpT", i)], bf[("WpT_lo", i)] = Wh_hi, Wh_lo, Wp_hi, Wp_lo
        fold = bf.get(("fold32", i))
        if fold is None:
            fold = torch.empty(8 * c.Hp, c.Hp, dtype=F32, device=self.device)
        for d in range(2):
            rows = slice(d * 4 * c.Hp, (d + 1) * 4 * c.Hp)
            # W'^T[g,h] = sum_p Wh[g,p] * WpT[p,h]
            gemm(Wh_hi[rows], Wp_hi[d], 0, 1, out=fold[rows])
            gemm(Wh_hi[rows], Wp_lo[d], 0, 1, out=fold[rows], accumulate=True)
            gemm(Wh_lo[rows], Wp_hi[d], 0, 1, out=fold[rows], accumulate=True)
        bf[("fold32", i)] = fold
        bf[("fold16", i)] = _cast16(fold, F16, bf.get(("fold16", i)))   # forward recurrence: (W')^T, fp16
        # BPTT keeps W' itself ([units, packed gate cols], bf16) in tensor memory: same product, other orientation
        foldb = bf.get(("foldb32", i))
        if foldb is None:
            foldb = torch.empty(2 * c.Hp, 4 * c.Hp, dtype=F32, device=self.device)
        for d in range(2):
            rows = slice(d * 4 * c.Hp, (d + 1) * 4 * c.Hp)
            out = foldb[d * c.Hp:(d + 1) * c.Hp]
            # W'[u,g] = sum_p WpT[p,u] * Wh[g,p]
            gemm(Wp_hi[d], Wh_hi[rows], 1, 0, out=out)
            gemm(Wp_lo[d], Wh_hi[rows], 1, 0, out=out, accumulate=True)
            gemm(Wp_hi[d], Wh_lo[rows], 1, 0, out=out, accumulate=True)
        bf[("foldb32", i)] = foldb
        bf[("fold", i)] = _cast16(foldb, BF16, bf.get(("fold", i)))

    def refresh_operands(self):
        """Rebuild the 16-bit operands after a weight update (once per step).  ~100 tiny launches: the first call runs them
        eagerly (allocating the buffers), the second captures them into two CUDA graphs -- layer 0 (replayed on the calling
        stream) and layers 1.. (replayed on a side stream, so they execute on the SMs layer 0's recurrence leaves idle;
        forward() waits on `_refresh_done` before it touches layer 1)."""
        if not self._stale:
            return
        c = self.cfg
        self._refresh_done = None
        have_buffers = ("fold", c.num_layers - 1) in self._bf
        if not (self.use_graphs and have_buffers and self.device.type == "cuda"):
            for i in range(c.num_layers):
                self._refresh_layer(i)
            self._stale = False
            return
        main = torch.cuda.current_stream()
        if self._refresh_graphs is None:
            torch.cuda.synchronize()
            L = _lib.lib()
            n0 = L.lcb_launch_count(0)
            g0 = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g0, capture_error_mode="thread_local"):   # the pipeline prefetch thread may pin memory meanwhile
                self._refresh_layer(0)
            n1 = L.lcb_launch_count(0)
            g1 = None
            if c.num_layers > 1:
                g1 = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g1, capture_error_mode="thread_local"):
                    for i in range(1, c.num_layers):
                        self._refresh_layer(i)
            n2 = L.lcb_launch_count(0)
            L.lcb_launch_count_add(-(n2 - n0))        # captured, not run
            self._refresh_graphs = (g0, g1)
            self._refresh_nodes = n2 - n0
        g0, g1 = self._refresh_graphs
        _lib.lib().lcb_launch_count_add(self._refresh_nodes)
        if g1 is not None:
            self.rstream.wait_stream(main)            # the weights the graph reads were written on the calling stream
            with torch.cuda.stream(self.rstream):
                g1.replay()
                ev = torch.cuda.Event()
                ev.record(self.rstream)
            self._refresh_done = ev
        g0.replay()
        self._stale = False

    def _await_refresh(self):
        if self._refresh_done is not None:
            torch.cuda.current_stream().wait_event(self._refresh_done)
            self._refresh_done = None

    # ------------------------------------------------------------------ workspaces
    def _workspace(self, T, B, training):
        """Views of the arena for a (T, B) minibatch.  Inference keeps ONE m buffer and two ping-pong layer outputs; training
        keeps every layer's m / gates / c for BPTT."""
        key = (T, B, training, self._arena.generation, self.g_half)
        if self._ws_key == key:
            return self._ws_views
        c, a = self.cfg, self._arena
        N, nl = T * B, c.num_layers
        ws = {"X0": a.rows("X0", N, c.Dp0, F16),
              "G": a.rows("G16" if self.g_half else "G", N, 8 * c.Hp, F16 if self.g_half else F32),
              "rec_ws": a.flat("rec_ws", max(16, _lib.lib().lcb_lstm_rec_workspace_bytes(B, c.Hp)), torch.uint8),
              "ready": a.flat("ready", max(c.num_layers, 1), torch.int32),
              "fwd_prog_words": _lib.lib().lcb_lstm_rec_fwd_progress_words(B, c.Hp, self.ndir),
              "cfin": a.flat("cfin", B * 2 * c.Hp, F32).view(B, 2, c.Hp),
              "mfin": a.flat("mfin", B * 2 * c.Hp, F32).view(B, 2, c.Hp)}
        ws["fwd_prog"] = a.flat("fwd_prog", max(1, nl * ws["fwd_prog_words"]), torch.int32)
        if training:
            ws["M"] = [a.rows("M%d" % i, N, 2 * c.Hp, F16) for i in range(nl)]
            ws["Hout"] = [a.rows("Hout%d" % i, N, 2 * c.P, F16) for i in range(nl)]
            # bf16 twins of the layer outputs, written by the output-projection GEMM's second store (lcb_gemm16_twin): the next
            # layer's / the output layer's weight-gradient operand.  None where the layer output is modified after the GEMM
            # (layer-0 residual, the uni-directional stack's residual add): backward() converts those.
            ws["Hbf"] = [None if (c.uni_residual(i) or (i == 0 and c.residual0)) else a.rows("Hbf%d" % i, N, 2 * c.P, BF16) for i in range(nl)]
            ws["gates"] = [a.rows("gates%d" % i, N, 2 * c.Hp, torch.int64) for i in range(nl)]   # 4 x fp16
            ws["cst"] = [a.rows("cst%d" % i, N, 2 * c.Hp, F32) for i in range(nl)]
            ws["dM"] = a.rows("dM", N, 2 * c.Hp, F32)
            # double-buffered by layer parity: the wgrad GEMMs of layer i run on a side stream while the main
            # stream already produces layer i-1's tensors
            ws["dG"] = [a.rows("dG%d" % k, N, 8 * c.Hp, BF16) for k in range(2)]
            # three, by layer % 3: layer i's dX is written (early rows) while layer i+1's wgrad still reads ITS dH = layer i+2's dX
            ws["dX"] = [a.rows("dX%d" % k, N, 2 * c.P, BF16) for k in range(3)]
            ws["dfold"] = [a.rows("dfold%d" % k, 4 * c.Hp, c.Hp, F32) for k in range(2)]
            ws["Xbf"] = [a.flat("Xbf%d" % k, N * max(c.Dp0, 2 * c.P), BF16) for k in range(2)]   # bf16 copies for wgrad
            # bf16 twins of every layer's m, written by the recurrence kernel beside the fp16 rows (wgrad operand; was: two rotating
            # scratch copies made by a conversion pass in backward())
            ws["Mbf"] = [a.rows("Mbf%d" % k, N, 2 * c.Hp, BF16) for k in range(nl)]
            ws["bwd_carry"] = a.flat("bwd_carry", B * 2 * c.Hp * 2, F32)
            ws["bwd_prog_words"] = _lib.lib().lcb_lstm_rec_bwd_progress_words(B, c.Hp, self.ndir)
            ws["bwd_prog"] = a.flat("bwd_prog", max(1, nl * ws["bwd_prog_words"]), torch.int32)
        else:
            m1 = a.rows("M0", N, 2 * c.Hp, F16)
            ho = [a.rows("Hout%d" % k, N, 2 * c.P, F16) for k in range(min(2, nl))]
            ws["M"] = [m1] * nl
            ws["Hout"] = [ho[i % len(ho)] for i in range(nl)]
        key = (T, B, training, a.generation, self.g_half)          # (allocation above may have bumped the generation)
        self._ws_key, self._ws_views = key, ws
        return ws

    def workspace_bytes(self):
        return self._arena.bytes_allocated()

    # ------------------------------------------------------------------ forward
    @property
    def head_frac(self):
        """First recurrence-launch boundary as a fraction of T (0: one launch); setting it selects a two-launch schedule."""
        return self.head_fracs[0] if self.head_fracs else 0.0

    @head_frac.setter
    def head_frac(self, v):
        self.head_fracs = [float(v)] if v and v > 0 else []

    def forward(self, nnet_input, seq_len, training=True, seq_len_host=None):
        """nnet_input [B,T,D] f32 cuda (zero padded), seq_len [B] int32 cuda.
        seq_len_host: the same lengths on the host (numpy / CPU tensor / sequence; every batch assembler has them) -- lets the
        recurrence skip, per 16-utterance group, the scan steps in which no utterance of the group is live
        (lcb_lstm_rec_fwd_range_hl); None = every group runs all T steps.  The result is the same either way.
        Returns the encoder output [T*B, 2P] fp16 (time-major rows n = t*B + b)."""
        L = _lib.lib()
        c = self.cfg
        assert nnet_input.is_cuda and nnet_input.dtype == F32 and nnet_input.dim() == 3
        B, T, D = nnet_input.shape
        assert D == c.input_dim, (D, c.input_dim)
        self.step_id += 1 if training else 0
        self.refresh_operands()
        ws = self._workspace(T, B, training)
        st = _lib.stream_ptr()
        nnet_input = nnet_input.contiguous()
        seq_len = seq_len.to(device=self.device, dtype=torch.int32).contiguous()
        _lib.check(L.lcb_pack_input(_lib.ptr(nnet_input), _lib.ptr(ws["X0"]), B, T, D, c.Dp0, st), "lcb_pack_input")
        ws["ready"].zero_()                         # flow-control counters (one per layer) of this pass
        ws["fwd_prog"].zero_()                      # progress words of the recurrence launches (per layer)
        ws["cfin"].zero_()                          # utterances of length 0 keep the zero initial state (bilstm.py:140-144)
        ws["mfin"].zero_()
        X = ws["X0"]
        nd = self.ndir
        lens_host = None
        if seq_len_host is not None:
            lens_host = (_lib.ctypes.c_int32 * B)(*[int(v) for v in (seq_len_host.tolist() if hasattr(seq_len_host, "tolist") else seq_len_host)])
        H4n = nd * 4 * c.Hp                         # gate columns that exist: both directions' or direction 0's
        side_cap = self.idle_sms(B, 0)              # SMs the forward recurrence clusters leave free
        pw = ws["fwd_prog_words"]
        for i in range(c.num_layers):
            if i == 1:
                self._await_refresh()               # layers 1.. were refreshed on the side stream during layer 0's recurrence
            peep = self.params.w("L%d/peep" % i) if c.use_peepholes else None
            gates = ws["gates"][i] if training else None
            cst = ws["cst"][i] if training else None
            last = (i == c.num_layers - 1)
            W16, bias, G = self._bf[("Wx16", i)], self.params.w("L%d/bias" % i), ws["G"]
            kin = X.shape[1] if (i == 0 or nd == 2) else c.P     # uni: layers 1.. read the forward half of the layer below only

            def rec(s0, s1, ready=None, progress=None):
                _lib.check(L.lcb_lstm_rec_fwd_range_pg(_lib.ptr(G), 2 if G.dtype == F16 else 0, _lib.ptr(self._bf[("fold16", i)]), _lib.ptr(peep),
                                                       _lib.ptr(seq_len),
                                                       lens_host, _lib.ptr(ready), _lib.ptr(ws["M"][i]),
                                                       _lib.ptr(ws["Mbf"][i]) if (training and self.bf16_twins) else None, _lib.ptr(gates), _lib.ptr(cst),
                                                       _lib.ptr(ws["cfin"]) if last else None, _lib.ptr(ws["mfin"]) if last else None,
                                                       T, B, c.Hp, nd, c.forget_bias, s0, s1, _lib.ptr(progress),
                                                       _lib.ptr(ws["rec_ws"]), ws["rec_ws"].numel(),
                                                       _lib.stream_ptr()), "lcb_lstm_rec_fwd_range_pg")

            Hout = ws["Hout"][i]
            Hbf = ws["Hbf"][i] if (training and self.bf16_twins) else None
            # h = m * W_proj with DropoutWrapper(output_keep_prob) (bilstm.py:128,137) applied in the GEMM epilogue: element
            # [n, d*P + p] of Hout uses element n*2P + d*P + p of the layer's mask stream
            drop = (c.keep_prob, self.dropout_seed(i)) if (training and c.keep_prob < 1.0) else None

            def hproj(s0, s1):
                """Output projection of the frames scan steps [s0, s1) visited: rows [s0,s1) of the forward, [T-s1,T-s0) of the
                backward direction's column half."""
                for d, (r0, r1) in enumerate(((s0 * B, s1 * B), ((T - s1) * B, (T - s0) * B))[:nd]):
                    gemm(ws["M"][i][r0:r1, d * c.Hp:(d + 1) * c.Hp], self._bf[("WpT16", i)][d], 0, 0,
                         out=Hout[r0:r1, d * c.P:(d + 1) * c.P],
                         dropout=(drop + (r0 * 2 * c.P + d * c.P,)) if drop else None,
                         out_bf16=Hbf[r0:r1, d * c.P:(d + 1) * c.P] if Hbf is not None else None)

            # scan-step boundaries of the recurrence launches / projection chunks (training, side stream available, long enough
            # sequences)
            bounds = []
            flow = training and self.pstream is not None and self.fwd_flow_control
            if training and self.pstream is not None:
                for fr in (self.flow_fracs if flow else (self.head_fracs_tight if side_cap <= self.tight_side_sms else self.head_fracs)):
                    b_ = int(math.ceil(fr * T))
                    if b_ >= 16 and b_ <= T - 16 and (not bounds or b_ >= bounds[-1] + 16):
                        bounds.append(b_)
            if not bounds:
                gemm(X[:, :kin], W16[:H4n, :kin], 0, 0, out=G[:, :H4n], bias=bias[:H4n])
                rec(0, T)
                hproj(0, T)
            else:
                # The recurrence needs G only for the frames it is about to visit: project the first scan steps of each direction
                # (frames [0,b0) for the forward, [T-b0,T) for the backward cells) on the whole chip, start the recurrence on them,
                # and project the following chunks, in the order the scan needs them, on a side stream on the SMs the 64
                # recurrence CTAs leave idle (grid capped); each further recurrence launch resumes bit-identically
                # (lcb_lstm_rec_fwd_range).  Chunks grow geometrically: the capped GEMM of chunk k+1 (1.17 ms per full sequence)
                # has to fit under the recurrence of chunks <= k (2.13 ms per full sequence).  The output projection of a chunk
                # follows on the side stream once its recurrence is done; only the last chunk's stays on the main stream.
                H4 = 4 * c.Hp
                main = torch.cuda.current_stream()
                bounds = bounds + [T]

                def proj(s0, s1):
                    r0, r1 = s0 * B, s1 * B
                    gemm(X[r0:r1, :kin], W16[:H4, :kin], 0, 0, out=G[r0:r1, :H4], bias=bias[:H4])
                    if nd == 2:
                        r0, r1 = (T - s1) * B, (T - s0) * B
                        gemm(X[r0:r1], W16[H4:], 0, 0, out=G[r0:r1, H4:], bias=bias[H4:])

                proj(0, bounds[0])
                ready = ws["ready"][i:i + 1] if flow else None
                if flow:
                    _lib.check(L.lcb_store_i32(_lib.ptr(ready), bounds[0], _lib.stream_ptr()), "lcb_store_i32")
                head_done = torch.cuda.Event()
                head_done.record(main)
                chunk_ready = []
                with torch.cuda.stream(self.pstream), grid_cap(side_cap):
                    self.pstream.wait_event(head_done)
                    for k in range(1, len(bounds)):
                        proj(bounds[k - 1], bounds[k])
                        if flow:
                            _lib.check(L.lcb_store_i32(_lib.ptr(ready), bounds[k], _lib.stream_ptr()), "lcb_store_i32")
                        ev = torch.cuda.Event()
                        ev.record(self.pstream)
                        chunk_ready.append(ev)
                prev = 0
                rel = []                            # scan steps at which finished rows are released to the output projection
                for fr in (self.fwd_hproj_fracs if (flow and pw > 0) else []):
                    s_ = (int(fr * T) // 16) * 16
                    if s_ >= 16 and s_ <= T - 16 and (not rel or s_ > rel[-1]):
                        rel.append(s_)
                if flow and rel:
                    # ... and it publishes its progress: the output projection of the scan steps before each release point follows
                    # on the side stream (behind the chunks, beside the rest of the scan); only the rows of the last steps stay
                    # between two layers.  (The launch is enqueued BEFORE the waits: a tool that serialises kernels in launch
                    # order still terminates.)
                    prog = ws["fwd_prog"][i * pw:(i + 1) * pw]
                    rec(0, T, ready, prog)
                    side_hp = torch.cuda.Event()
                    with torch.cuda.stream(self.pstream), grid_cap(side_cap):
                        for s0_, s1_ in zip([0] + rel[:-1], rel):
                            _lib.check(L.lcb_wait_progress(_lib.ptr(prog), pw, s1_, _lib.stream_ptr()), "lcb_wait_progress")
                            hproj(s0_, s1_)
                        side_hp.record(self.pstream)
                    hproj(rel[-1], T)
                    main.wait_event(side_hp)
                elif flow:
                    # every chunk is enqueued: ONE launch over the whole scan, its prefetch warps wait for the counter
                    rec(0, T, ready)
                    main.wait_event(chunk_ready[-1])
                    hproj(0, T)
                for k, b_ in enumerate(bounds if not flow else []):
                    if k > 0:
                        main.wait_event(chunk_ready[k - 1])
                    rec(prev, b_)
                    if k + 1 < len(bounds) and self.overlap_hproj:
                        rec_done = torch.cuda.Event()
                        rec_done.record(main)
                        with torch.cuda.stream(self.pstream), grid_cap(side_cap):
                            self.pstream.wait_event(rec_done)
                            hproj(prev, b_)
                    prev = b_
                if flow:
                    pass
                elif self.overlap_hproj:
                    hproj(bounds[-2], T)
                    hp_done = torch.cuda.Event()
                    hp_done.record(self.pstream)
                    main.wait_event(hp_done)
                else:
                    hproj(0, T)
            if c.uni_residual(i):
                # DropoutWrapper(ResidualWrapper(cell)): out = dropout(x + h) -- the GEMM epilogue wrote dropout(h); add dropout(x)
                # with the same mask (forward half of the columns only; the other half is identically zero in this mode)
                src = ws["X0"] if i == 0 else ws["Hout"][i - 1]
                keep = c.keep_prob if training else 1.0
                _lib.check(L.lcb_masked_add16(_lib.ptr(Hout), 2 * c.P, _lib.ptr(src), src.stride(0), T * B, c.P, 2, keep,
                                              self.dropout_seed(i), 0, 2 * c.P, st), "lcb_masked_add16")
            if i == 0 and c.residual0:              # finput = finput + concat(...)  iff input_dim == 2*num_projects (bilstm.py:199-200)
                _lib.check(L.lcb_add_f16(_lib.ptr(Hout), _lib.ptr(ws["X0"]), Hout.numel(), st), "lcb_add_f16")
            X = Hout
        self._last = (T, B, seq_len, training)
        return X

    def encoder_state(self):
        """`encoder` of bilstm.py:206-208: concat(c_fw, h_fw, c_bw, h_bw) of the LAST layer, [B, 2(H+P)]."""
        T, B, _, training = self._last
        ws = self._workspace(T, B, training)
        c = self.cfg
        i = c.num_layers - 1
        out = []
        for d in range(2):
            h = gemm(ws["mfin"][:, d].contiguous().to(F16), self._bf[("WpT16", i)][d], 0, 0)
            out += [ws["cfin"][:, d, :c.H], h]
        return torch.cat(out, 1)

    # ------------------------------------------------------------------ backward
    def backward(self, dXtop, bucket_ready=None, top_dropped=False, top_ready=None):
        """dXtop [T*B, 2P] bf16 = d loss / d encoder output.  Accumulates parameter gradients into
        params.gflat (which the caller zeroed).  bucket_ready(name_list) is called as soon as the
        gradients of a layer are final (data-parallel all-reduce hook).

        top_ready: [(scan step s_k, event)] with increasing s_k, the last one T -- the rows of dXtop for the frames BPTT visits
        in scan steps < s_k (frames [0, s_k) and [T - s_k, T)) are final once the event has fired.  The top layer's BPTT then runs
        as consecutive range launches [s_{k-1}, s_k), each started as soon as its rows exist (AcousticModel.backward).

        Stream structure: the serial chain (dM GEMM -> BPTT -> dX GEMM) stays on the current stream; the
        weight-gradient GEMMs of layer i are enqueued on a side stream and execute on the SMs the next layer's
        BPTT clusters leave free."""
        L = _lib.lib()
        c = self.cfg
        T, B, seq_len, training = self._last
        assert training
        ws = self._workspace(T, B, True)
        N = T * B
        ps = self.params
        nd = self.ndir
        H4n = nd * 4 * c.Hp
        bwd_cap = self.idle_sms(B, 1)        # SMs the BPTT clusters leave free: grid cap of everything that runs beside them
        main = torch.cuda.current_stream()
        side = self.wstream if (self.overlap_wgrad and self.wstream is not None) else main
        overlap = side is not main
        side_done = {}                       # layer -> event: its side-stream work (reads of set k, dH) has finished
        dH = dXtop
        if overlap:
            side.wait_stream(main)           # the forward activations the side stream converts are complete

        def dfold_rows(d, r0, r1, M, dG_, acc):
            """share of the rows [r0, r1) of dz in dW'^T[g,h] = sum_n dz_n[g] * m_prev(n)[h] of direction d (prev = the row B
            before / after): into ws["dfold"][d], accumulating if acc"""
            if d == 0:
                r0 = max(r0, B)
            else:
                r1 = min(r1, N - B)
            if r1 <= r0:
                return
            sh = -B if d == 0 else B
            gemm(dG_[r0:r1, d * 4 * c.Hp:(d + 1) * 4 * c.Hp], M[r0 + sh:r1 + sh, d * c.Hp:(d + 1) * c.Hp], 1, 1,
                 out=ws["dfold"][d], accumulate=acc)

        def wgrad(i, dH_this, after, released=()):
            """Weight gradients of layer i on the side stream.  `after` = event on the main stream behind the serial chain
            (dX, dropout, dM GEMMs) of the layer BELOW, so these GEMMs never take SMs from it: the BPTT launch that follows
            leaves them 84 idle SMs for ~3.8 ms (capped grid); layer 0's run alone on the whole chip.
            `released` (layer 0): [(frame blocks, event)] -- frames whose dG was final after a partial BPTT launch; their share of
            the frame-sums dW' and dW_x is accumulated beside the remaining launches, only the outer frames' share is left for the
            tail after BPTT."""
            k = i & 1
            dG = ws["dG"][k]
            gWpT, gWh, gWx = ps.g("L%d/WpT" % i), ps.g("L%d/Wh" % i), ps.g("L%d/Wx" % i)
            X16 = ws["X0"] if i == 0 else ws["Hout"][i - 1]
            split = overlap and i == 0        # nothing runs beside layer 0's weight gradients: direction 1 goes to the main stream
            with torch.cuda.stream(side):
                if overlap and not split:
                    side.wait_event(after)
                # bf16 copies of the fp16 forward activations (tcgen05 kind::f16 cannot mix f16 x bf16 operands); layer 0's are
                # made while its BPTT still runs
                if self.bf16_twins and i > 0 and ws["Hbf"][i - 1] is not None:
                    X = ws["Hbf"][i - 1]             # written beside the fp16 rows by the layer below's output projection
                else:
                    X = _to_bf16(X16, ws["Xbf"][k][:X16.numel()].view(X16.shape))
                # m: written by the forward recurrence beside the fp16 rows, or converted here
                M = ws["Mbf"][i] if self.bf16_twins else _to_bf16(ws["M"][i], ws["Mbf"][i])
                if split:
                    # dW_p^T = dH^T * M needs nothing from layer 0's BPTT: both directions run beside it (capped grid)
                    with grid_cap(max(8, bwd_cap - 4)):
                        for d in range(nd):
                            gemm(dH_this[:, d * c.P:(d + 1) * c.P], M[:, d * c.Hp:(d + 1) * c.Hp], 1, 1, out=gWpT[d])
                        done_blocks = []
                        for blocks, ev in released:
                            if isinstance(ev, tuple):          # (event before the launch, progress words, count, scan step)
                                side.wait_event(ev[0])
                                _lib.check(L.lcb_wait_progress(_lib.ptr(ev[1]), ev[2], ev[3], _lib.stream_ptr()), "lcb_wait_progress")
                            else:
                                side.wait_event(ev)
                            for (t0, t1) in blocks:
                                for d in range(nd):
                                    dfold_rows(d, t0 * B, t1 * B, M, dG, bool(done_blocks))
                                gemm(dG[t0 * B:t1 * B, :H4n], X[t0 * B:t1 * B], 1, 1, out=gWx[:H4n], accumulate=bool(done_blocks))
                                done_blocks.append((t0, t1))
                    converted = torch.cuda.Event()
                    converted.record(side)
                    side.wait_event(after)
                    # frames not covered yet: the two outer bands (or everything)
                    rest = [(0, T)] if not done_blocks else [(0, min(b[0] for b in done_blocks)), (max(b[1] for b in done_blocks), T)]
                # layers 1..: beside the BPTT of the layer below (capped grid); layer 0's tail runs alone on the whole chip
                cap_ctx = grid_cap(max(8, bwd_cap - 4) if (overlap and i > 0) else 0)
                cap_ctx.__enter__()
                xin = X if (i == 0 or nd == 2) else X[:, :c.P]      # uni: layers 1.. read the forward half of the layer below only

                def direction(d):
                    dHd = dH_this[:, d * c.P:(d + 1) * c.P]
                    Md = M[:, d * c.Hp:(d + 1) * c.Hp]
                    dGd = dG[:, d * 4 * c.Hp:(d + 1) * 4 * c.Hp]
                    # dW_p^T[p,h] = sum_n dH[n,p] * M[n,h]
                    if not split:
                        gemm(dHd, Md, 1, 1, out=gWpT[d])
                    dfold = ws["dfold"][d]
                    if T > 1:
                        # dW'^T[g,h] = sum_n dz_n[g] * m_prev(n)[h]; prev = t-1 (fwd) / t+1 (bwd): a row shift of B
                        if split and done_blocks:
                            for (t0, t1) in rest:
                                dfold_rows(d, t0 * B, t1 * B, M, dG, True)
                        elif d == 0:
                            gemm(dGd[B:], Md[:N - B], 1, 1, out=dfold)
                        else:
                            gemm(dGd[:N - B], Md[B:], 1, 1, out=dfold)
                        df_hi, df_lo = _split_bf16(dfold)
                        rows = slice(d * 4 * c.Hp, (d + 1) * 4 * c.Hp)
                        # dW_h[g,p] = sum_h dW'^T[g,h] * WpT[p,h]
                        gemm(df_hi, self._bf[("WpT", i)][d], 0, 0, out=gWh[rows])
                        gemm(df_lo, self._bf[("WpT", i)][d], 0, 0, out=gWh[rows], accumulate=True)
                        # dW_p^T[p,h] += sum_g Wh[g,p] * dW'^T[g,h]
                        gemm(self._bf[("Wh", i)][rows], df_hi, 1, 1, out=gWpT[d], accumulate=True)
                        gemm(self._bf[("Wh", i)][rows], df_lo, 1, 1, out=gWpT[d], accumulate=True)

                direction(0)
                if nd == 2:
                    if split:
                        with torch.cuda.stream(main), grid_cap(0):
                            main.wait_event(converted)
                            direction(1)
                            dir1_done = torch.cuda.Event()
                            dir1_done.record(main)
                    else:
                        direction(1)
                # dW_x[g,k] = sum_n dz_n[g] * x_n[k]   (both directions at once)
                gWxn = gWx[:H4n, :xin.shape[1]]
                if split and done_blocks:
                    for (t0, t1) in rest:
                        if t1 > t0:
                            gemm(dG[t0 * B:t1 * B, :H4n], xin[t0 * B:t1 * B], 1, 1, out=gWxn, accumulate=True)
                else:
                    gemm(dG[:, :H4n], xin, 1, 1, out=gWxn)
                if split and nd == 2:
                    side.wait_event(dir1_done)
                cap_ctx.__exit__(None, None, None)
                if bucket_ready is not None:
                    bucket_ready(["L%d/WpT" % i, "L%d/Wh" % i, "L%d/peep" % i, "L%d/bias" % i, "L%d/Wx" % i])
                if overlap:
                    ev = torch.cuda.Event()
                    ev.record(side)
                    side_done[i] = ev

        def mark():
            if not overlap:
                return None
            ev = torch.cuda.Event()
            ev.record(main)
            return ev

        def dm_rows(i, dH_, r0, r1):
            """dM[r0:r1] = dH[r0:r1] * W_p^T of layer i, both directions"""
            if r1 <= r0:
                return
            for d in range(nd):
                gemm(dH_[r0:r1, d * c.P:(d + 1) * c.P], self._bf[("WpT", i)][d], 0, 1, out=ws["dM"][r0:r1, d * c.Hp:(d + 1) * c.Hp])

        def dx_rows(i, dG_, dXn_, r0, r1, dHcur=None):
            """dX[r0:r1] = dG[r0:r1] * W_x of layer i, with the mask of layer i-1's output dropout (same seed and element indices as
            the forward pass) in the epilogue"""
            if r1 <= r0:
                return
            pw = 2 * c.P if nd == 2 else c.P         # uni: only the forward half of the layer below exists (the other half stays 0)
            gemm(dG_[r0:r1, :H4n], self._bf[("Wx", i)][:H4n, :pw], 0, 1, out=dXn_[r0:r1, :pw],
                 dropout=(c.keep_prob, self.dropout_seed(i - 1), r0 * 2 * c.P) if c.keep_prob < 1.0 else None)
            if c.uni_residual(i):
                # residual path of layer i: d loss / d x_i += d loss / d (x_i + h_i) = this layer's (already masked) output gradient,
                # then through layer i-1's output dropout like the GEMM result
                _lib.check(L.lcb_masked_add16(_lib.ptr(dXn_[r0:r1]), 2 * c.P, _lib.ptr(dHcur[r0:r1]), dHcur.stride(0), r1 - r0, c.P, 1,
                                              c.keep_prob, self.dropout_seed(i - 1), r0 * 2 * c.P, 2 * c.P, _lib.stream_ptr()),
                           "lcb_masked_add16")

        # "Early rows": scan step s of BPTT visits frame T-1-s in the forward and frame s in the backward direction, so after the
        # scan steps [0, Tb) with Tb > T/2 the frames [T-Tb, Tb) have their final dG in BOTH directions.  Their rows of dX -- and of
        # the dM of the layer below, which needs nothing else -- are computed on a side stream beside the launch over
        # [Tb, T); only the rows of the first and last T-Tb frames stay on the serial chain between two layers' BPTT.  With several
        # cuts each launch releases the two bands of frames between its cut and the previous one.
        cuts = []
        if overlap and self.xstream is not None and L.lcb_lstm_rec_bwd_can_split(c.Hp):
            for fr in self.bwd_early_fracs:
                Tb = int(math.ceil(fr * T))
                lo = cuts[-1] + 16 if cuts else (T - Tb) + 16       # first cut: at least 16 frames final in both directions
                if fr > 0.5 and Tb >= lo and T - Tb >= 16:
                    cuts.append(Tb)
        released0 = []
        pwords = ws["bwd_prog_words"]
        use_pg = bool(cuts) and self.bwd_progress and pwords > 0
        if use_pg:
            ws["bwd_prog"].zero_()
        early = None                         # (first row, end row, event): rows of dH and of this layer's dM made on xstream
        pending = None                       # (layer, its dH): weight gradients not yet enqueued
        for i in reversed(range(c.num_layers)):
            k = i & 1
            if overlap and (i + 2) in side_done:
                main.wait_event(side_done[i + 2])             # buffer set k is free again
            if c.keep_prob < 1.0 and i == c.num_layers - 1 and not top_dropped:
                self._dropout(dH, i)                # same (seed, index) mask as the forward pass, on the gradient
            dM, dG = ws["dM"], ws["dG"][k]
            # dM = dH * W_p^T
            top_s0 = 0                       # scan steps of the top layer already run by the launches below
            if top_ready and i == c.num_layers - 1 and early is None and L.lcb_lstm_rec_bwd_can_split(c.Hp):
                peep_t = ps.w("L%d/peep" % i) if c.use_peepholes else None
                gpeep_t = ps.g("L%d/peep" % i) if c.use_peepholes else None
                for kk, (sk, ev) in enumerate(top_ready):
                    main.wait_event(ev)
                    a0, a1 = top_s0, min(sk, T)
                    if a1 >= T - a1:                                 # the bands have met: one block of rows
                        dm_rows(i, dH, a0 * B, (T - a0) * B)
                    else:
                        dm_rows(i, dH, a0 * B, a1 * B)
                        dm_rows(i, dH, (T - a1) * B, (T - a0) * B)
                    if kk + 1 < len(top_ready):
                        _lib.check(L.lcb_lstm_rec_bwd_range(_lib.ptr(dM), _lib.ptr(ws["gates"][i]), _lib.ptr(ws["cst"][i]),
                                                            _lib.ptr(self._bf[("fold", i)]), _lib.ptr(peep_t), _lib.ptr(seq_len),
                                                            _lib.ptr(dG), _lib.ptr(ps.g("L%d/bias" % i)), _lib.ptr(gpeep_t),
                                                            T, B, c.Hp, nd, a0, a1, _lib.ptr(ws["bwd_carry"]),
                                                            _lib.ptr(ws["rec_ws"]), ws["rec_ws"].numel(), _lib.stream_ptr()),
                                   "lcb_lstm_rec_bwd_range")
                        top_s0 = a1
            elif early is None:
                dm_rows(i, dH, 0, N)
            else:
                dm_rows(i, dH, 0, early[0])
                dm_rows(i, dH, early[1], N)
                main.wait_event(early[2])
                early = None
            chain_issued = mark()                             # behind BPTT(i+1) and this layer's dropout / dM GEMMs
            peep = ps.w("L%d/peep" % i) if c.use_peepholes else None
            gpeep = ps.g("L%d/peep" % i) if c.use_peepholes else None
            # BPTT in one launch, or as two launches over consecutive scan ranges joined by the carry buffer -- bit-identical
            # (lcb_lstm_rec_bwd_range): at Tb for the early rows above (layers 1..), or at bwd_split_frac (experiments)
            if cuts and all(cc > top_s0 for cc in cuts):
                ranges = list(zip([top_s0] + cuts, cuts + [T]))
            elif top_s0 > 0:
                ranges = [(top_s0, T)]
            else:
                Ts = int(math.ceil(self.bwd_split_frac * T)) if self.bwd_split_frac > 0 and L.lcb_lstm_rec_bwd_can_split(c.Hp) else 0
                ranges = [(0, T)] if (Ts < 1 or Ts >= T) else [(0, Ts), (Ts, T)]
            dXn = ws["dX"][i % 3] if i > 0 else None
            prev_cut = None
            prog = ws["bwd_prog"][i * pwords:(i + 1) * pwords] if use_pg else None
            if use_pg:
                # one launch over the whole scan; the release points below are waited for on the consumers' stream
                pre = torch.cuda.Event()
                pre.record(main)
                _lib.check(L.lcb_lstm_rec_bwd_range_pg(_lib.ptr(dM), _lib.ptr(ws["gates"][i]), _lib.ptr(ws["cst"][i]),
                                                       _lib.ptr(self._bf[("fold", i)]), _lib.ptr(peep),
                                                       _lib.ptr(seq_len), _lib.ptr(dG), _lib.ptr(ps.g("L%d/bias" % i)), _lib.ptr(gpeep),
                                                       T, B, c.Hp, nd, top_s0, T, _lib.ptr(ws["bwd_carry"]), _lib.ptr(prog),
                                                       _lib.ptr(ws["rec_ws"]), ws["rec_ws"].numel(), _lib.stream_ptr()), "lcb_lstm_rec_bwd_range_pg")
            for (s0, s1) in ranges:
                if not use_pg:
                    _lib.check(L.lcb_lstm_rec_bwd_range(_lib.ptr(dM), _lib.ptr(ws["gates"][i]), _lib.ptr(ws["cst"][i]),
                                                        _lib.ptr(self._bf[("fold", i)]), _lib.ptr(peep),
                                                        _lib.ptr(seq_len), _lib.ptr(dG), _lib.ptr(ps.g("L%d/bias" % i)), _lib.ptr(gpeep),
                                                        T, B, c.Hp, nd, s0, s1, _lib.ptr(ws["bwd_carry"]),
                                                        _lib.ptr(ws["rec_ws"]), ws["rec_ws"].numel(), _lib.stream_ptr()), "lcb_lstm_rec_bwd_range")
                if cuts and s1 < T:
                    # frames final in both directions now: [T-s1, s1), minus those the previous cut already released
                    blocks = [(T - s1, s1)] if prev_cut is None else [(T - s1, T - prev_cut), (prev_cut, s1)]
                    prev_cut = s1
                    if use_pg:
                        launched = (pre, prog, pwords, s1)
                    else:
                        launched = torch.cuda.Event()
                        launched.record(main)
                    if i == 0:
                        if self.l0_released:
                            released0.append((blocks, launched))      # layer 0 has no dX: its weight gradients use the released frames
                        continue
                    with torch.cuda.stream(self.xstream), grid_cap(bwd_cap):
                        if use_pg:
                            self.xstream.wait_event(pre)
                            _lib.check(L.lcb_wait_progress(_lib.ptr(prog), pwords, s1, _lib.stream_ptr()), "lcb_wait_progress")
                        else:
                            self.xstream.wait_event(launched)
                        for (t0, t1) in blocks:
                            dx_rows(i, dG, dXn, t0 * B, t1 * B, dH)
                            dm_rows(i - 1, dXn, t0 * B, t1 * B)
                        ev = torch.cuda.Event()
                        ev.record(self.xstream)
                    early = ((T - s1) * B, s1 * B, ev)
            if self.debug_dz is not None:
                self.debug_dz[i] = dG.clone()
            if pending is not None:
                wgrad(pending[0], pending[1], chain_issued)   # layer i+1's weight gradients run beside this BPTT
            dH_this = dH
            if i > 0:
                # dX = dG * W_x (layer i+1's wgrad reads ITS dH from another of the three buffers)
                if early is None:
                    dx_rows(i, dG, dXn, 0, N, dH)
                else:
                    dx_rows(i, dG, dXn, 0, early[0], dH)
                    dx_rows(i, dG, dXn, early[1], N, dH)
                dH = dXn
            pending = (i, dH_this)
        wgrad(pending[0], pending[1], mark(), released0)
        if overlap:
            main.wait_stream(side)
        return None
