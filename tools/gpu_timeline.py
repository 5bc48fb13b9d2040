"""Per-call device timeline of one c3 training step (GPU box): every C-ABI call is bracketed by CUDA events on the
stream it is issued on; prints start offset / duration / stream, and the idle gaps of the main stream."""
import sys
import torch
sys.path.insert(0, ".")
import bench  # noqa: E402
from lstm_ctc_b200 import _lib  # noqa: E402
from lstm_ctc_b200.model import AcousticModel  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "c3"
w = bench.WORKLOADS[wl]
dev = torch.device("cuda:0")
model = AcousticModel(bench.nnet_config(w, 0.9), dev, seed=1234)
x, lens, y = [t.to(dev) for t in bench.synth_batch(w, 777)]


def step():
    model.loss_and_grad(x, lens, y, check_labels=False, seq_len_host=lens.cpu())
    model.optimizer_step("adam", 4e-4, clip_norm=5.0, l2_decay_weight=1e-5)


for _ in range(3):
    step()
torch.cuda.synchronize()
L = _lib.lib()
recs = []
SKIP = {"lcb_status_string", "lcb_device_error", "lcb_launch_count", "lcb_gemm_set_max_ctas", "lcb_lstm_rec_workspace_bytes",
        "lcb_lstm_rec_config", "lcb_ctc_workspace_bytes", "lcb_version", "lcb_lstm_rec_max_clusters"}


class Proxy:
    def __getattr__(self, k):
        f = getattr(L, k)
        if k in SKIP or not k.startswith("lcb_"):
            return f

        def wrap(*a):
            st = torch.cuda.current_stream()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(st)
            r = f(*a)
            e.record(st)
            tag = k
            if k == "lcb_gemm16":
                tag = "gemm M%d N%d K%d a%d b%d c%d acc%d" % (a[0], a[1], a[2], a[5], a[9], a[13], a[15])
            recs.append((tag, st.cuda_stream, s, e))
            return r
        return wrap


main = torch.cuda.current_stream().cuda_stream
_lib_lib = _lib.lib
_lib.lib = lambda: Proxy()
t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
t0.record()
step()
t1.record()
torch.cuda.synchronize()
_lib.lib = _lib_lib
print("step %.3f ms (instrumented), %d calls" % (t0.elapsed_time(t1), len(recs)))
rows = [(t0.elapsed_time(s), t0.elapsed_time(e), tag, st) for tag, st, s, e in recs]
rows.sort()
last_end = 0.0
gap_tot = 0.0
for a, b, tag, st in rows:
    m = "M" if st == main else "s"
    gap = ""
    if st == main:
        if a - last_end > 0.02:
            gap = "   <-- main idle %.3f ms before" % (a - last_end)
            gap_tot += a - last_end
        last_end = max(last_end, b)
    print("%9.3f %8.3f %s %s%s" % (a, b - a, m, tag, gap))
print("main-stream idle (gaps > 20 us) total %.3f ms" % gap_tot)
