"""(Needs a library built with `python tools/build.py --tuning`: the timeline stamps are compiled out otherwise.)
Phase timeline of the five-phase persistent decode kernel (decode_mega2.cu): per phase, over the CTAs, the time from
the phase's grid-barrier wait to the arrival at the next barrier, and the hand-over between them.
   python tools/trace_mega2.py [B] [layer]"""
import ctypes as C
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import clipcap_b200 as cc
from clipcap_b200 import synthetic

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
cfg = cc.EngineConfig(max_images=B, max_beam=1, max_ctx=80)
eng = cc.Engine(cfg)
synthetic.load_synthetic(eng)
info = (C.c_int * 4)()
eng.lib.ccb_debug_mega_info(eng._h, info)
ncta, L = info[1], cfg.lm_layers
nbar = 5 * L + 1
per = 2 * (nbar + 2)
trace = torch.zeros(ncta * per + ncta * 64, dtype=torch.int64, device="cuda")
images = synthetic.synthetic_images(B, cfg, device="cuda")
p = eng.gen_params("greedy", 8, stop_token=-1, max_stops=0)
eng.lib.ccb_debug_set_mega(eng._h, 3)
eng.lib.ccb_debug_mega_trace(eng._h, C.c_void_p(trace.data_ptr()))   # (before the first call: the step is captured in a CUDA graph)
eng.caption_images(images, p)
eng.caption_images(images, p)
torch.cuda.synchronize()
t = trace.cpu()[:ncta * per].view(ncta, per).double() / 1e3   # us
# stamp order per CTA: arrive#0, then per layer (wait, arrive) x 5, then wait, arrive (ln_f)
names = ["A qkv", "B attn", "C proj", "D fc", "E fc2"]
tot = (t[:, 1 + 10 * L + 1].max() - t[:, 0].min())
print("ncta %d; kernel span (embed arrive -> ln_f arrive) %.1f us = %.2f us per layer" % (ncta, tot, tot / L))
lay = [int(sys.argv[2])] if len(sys.argv) > 2 else list(range(2, L - 1))
acc = {n: [0.0] * 6 for n in names}
for l in lay:
    for k, n in enumerate(names):
        w = t[:, 1 + 10 * l + 2 * k]          # wait stamp (phase start)
        a = t[:, 1 + 10 * l + 2 * k + 1]      # arrive stamp (phase end)
        nxt = t[:, 1 + 10 * l + 2 * k + 2]    # next phase's wait
        body = a - w
        v = [float(body.median()), float(body.max()), float(a.max() - w.min()), float(nxt.median() - a.max()), float(a.max() - a.median()), float(w.max() - w.min())]
        for i in range(6):
            acc[n][i] += v[i] / len(lay)
print("%-8s %10s %10s %12s %12s %14s %12s" % ("phase", "body med", "body max", "first->last", "hand-over", "last - median", "wait spread"))
s = 0.0
for n in names:
    v = acc[n]
    s += v[2] + v[3]
    print("%-8s %10.2f %10.2f %12.2f %12.2f %14.2f %12.2f" % (n, *v))
print("sum (first->last + hand-over): %.2f us per layer" % s)

r = trace.cpu()[ncta * per:].view(ncta, 64).double() / 1e3
labels = {0: "start", 1: "rowstat / x polled", 2: "staged", 3: "acc ready", 4: "tile 0 sent", 5: "recv ready", 6: "done", 7: "first mma"}
for kind, kn in enumerate(["A qkv", "C proj", "D fc", "E fc2"]):
    base = r[:, kind * 8]
    for k in (1, 7, 2, 3, 4, 5, 6):
        col = r[:, kind * 8 + k]
        ok = (col > 0) & (base > 0)
        if int(ok.sum()) == 0:
            continue
        dlt = (col - base)[ok]
        print("  L1 %-7s %-20s n %3d  med %6.2f  max %6.2f  min %6.2f" % (kn, labels[k], int(ok.sum()), float(dlt.median()), float(dlt.max()), float(dlt.min())))
