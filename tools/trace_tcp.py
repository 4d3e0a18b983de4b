"""Dev tool: timeline of one minibatch step of the pipelined training kernel (CTA 0)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from multi_modal_normative_modeling_b200 import EnsembleTrainer, workloads, _lib
if os.environ.get('NMB_LIB'): _lib.LIB_PATH = os.path.abspath(os.environ['NMB_LIB'])   # A/B against another build

seeds = int(sys.argv[1]) if len(sys.argv) > 1 else 1
dev = torch.device("cuda", 0)
hw = workloads.build_host_workload()
wl = workloads.to_device(hw, dev, n_seeds=seeds)
specs = wl.specs
if 'big' in sys.argv:      # trace an early-fusion (D=348) member
    specs = [sp for sp, t in zip(wl.specs, wl.tags) if t[1].startswith('early')]
    sys.argv.remove('big')
tr = EnsembleTrainer(specs, device=dev)
tr.train_steps(8)
buf = torch.zeros(4096, dtype=torch.int64, device=dev)
lib = _lib.load()
_lib.check(lib.nmb_debug_tcp_trace(buf.data_ptr(), 2))
tr.train_steps(8)
torch.cuda.synchronize()
_lib.check(lib.nmb_debug_tcp_trace(None, 0))
b = buf.cpu().numpy().astype(np.int64)
nz = b[:2048][b[:2048] > 0]        # wall-clock stamps; the fine (cycle) stamps live above 2048
t0 = nz.min()
E = 37 if len(sys.argv) <= 2 else int(sys.argv[2]); S = 62 if len(sys.argv) <= 3 else int(sys.argv[3])
print("span us", (nz.max() - t0) / 1e3)
print("epilogue items: arrive start end (us)  | wait work")
for k in range(E):
    a, s, e = b[3 * k:3 * k + 3]
    if a: print("E%02d %8.2f %8.2f %8.2f | %6.2f %6.2f" % (k + 1, (a - t0) / 1e3, (s - t0) / 1e3, (e - t0) / 1e3, (s - a) / 1e3, (e - s) / 1e3))
for grp, tag in ((1, "G"), (2, "A")):
    print("group %d items" % grp)
    o1 = 3 * E * grp + 5 * S
    for k in range(E):
        a, s_, e = b[o1 + 3 * k:o1 + 3 * k + 3]
        if a: print("%s%02d %8.2f %8.2f %8.2f | %6.2f %6.2f" % (tag, k + 1, (a - t0) / 1e3, (s_ - t0) / 1e3, (e - t0) / 1e3, (s_ - a) / 1e3, (e - s_) / 1e3))
print("mma steps: deps tiles issued")
for k in range(S):
    a, s, e = b[3 * E + 3 * k:3 * E + 3 * k + 3]
    if a: print("S%02d %8.2f %8.2f %8.2f | %6.2f %6.2f" % (k, (a - t0) / 1e3, (s - t0) / 1e3, (e - t0) / 1e3, (s - a) / 1e3, (e - s) / 1e3))
print("producer steps: deps issued")
for k in range(S):
    a, e = b[3 * E + 3 * S + 2 * k:3 * E + 3 * S + 2 * k + 2]
    if a: print("P%02d %8.2f %8.2f" % (k, (a - t0) / 1e3, (e - t0) / 1e3))

# fine stamps (libraries built with -DNMB_TCP_FINE_TRACE only): SM clock cycles of group 0's publishing thread
def fine(name, off, n, per):
    v = b[off:off + n]
    v = v[v > 0]
    if len(v) < 2:
        return
    d = np.diff(v)
    print(name, "(deltas in SM cycles, %d per row)" % per)
    for i in range(0, len(d), per):
        print("   " + " ".join("%6d" % x for x in d[i:i + per]))
fine("head->latent item, half 0: [tmem ld | z math + scratch stores | per plane group: template load, split, 4 stores ...]", 2048, 64, 12)
fine("reconstruction item, half 0, per chunk: [ld issue | column sums of previous chunk | ld wait | math | planes | loop + prefetch]", 2048 + 64, 256, 5)

# loop-level stamps of group 0 (fine-trace builds): per item [decoded | before acc wait | after acc wait | item done | fenced | barrier | published]
lv = b[2048 + 512:2048 + 512 + 8 * E].reshape(E, 8)
if (lv > 0).any():
    print("group 0 item loop, SM cycles: item | prep | acc-wait | work | fence | bar | publish | -> next item decoded")
    prev_end = None
    rows = [(k, lv[k]) for k in range(E) if lv[k][0] > 0]
    for n, (k, r) in enumerate(rows):
        nxt = rows[n + 1][1][0] - r[6] if n + 1 < len(rows) else 0
        print("  E%02d %6d %6d %6d %6d %6d %6d | %6d" % (k + 1, r[1] - r[0], r[2] - r[1], r[3] - r[2], r[4] - r[3], r[5] - r[4], r[6] - r[5], nxt))
