"""(Needs a library built with `python tools/build.py --tuning`: the timeline stamps are compiled out otherwise.)
A/B of the decode step: persistent kernel vs operator-per-kernel chain (GPT2-XL, synthetic weights).
   python tools/ab_mega.py [B] [T] [mode] [trace]"""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import clipcap_b200 as cc
from clipcap_b200 import synthetic

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
T = int(sys.argv[2]) if len(sys.argv) > 2 else 8
mode = sys.argv[3] if len(sys.argv) > 3 else "greedy"
want_trace = len(sys.argv) > 4 and sys.argv[4] == "trace"
beam = 5 if mode == "beam" else 1
cfg = cc.EngineConfig(max_images=B, max_beam=beam, max_ctx=80)
eng = cc.Engine(cfg)
synthetic.load_synthetic(eng)
torch.cuda.empty_cache()
images = synthetic.synthetic_images(B, cfg, device="cuda")
p = eng.gen_params(mode, T, stop_token=-1, max_stops=0, top_p=0.9 if mode == "sample" else 0.0, beam_size=beam, seed=1)
L = cfg.lm_layers
res = {}
for flag in (0, 1):
    have = eng.lib.ccb_debug_set_mega(eng._h, flag)
    if flag == 1 and want_trace:
        ncta = 148
        per = 2 * (8 * L + 2)
        trace = torch.zeros(ncta * per + ncta * 64, dtype=torch.int64, device="cuda")
        eng.lib.ccb_debug_mega_trace(eng._h, C.c_void_p(trace.data_ptr()))
    for it in range(3):
        tokens, lengths, scores = eng.caption_images(images, p)
        torch.cuda.synchronize()
        pre, dec, steps = eng.last_timing()
    print("mega=%d (covered=%d): prefill %.2f ms, decode %.3f ms/step over %d steps" % (flag, have, pre, dec / max(steps, 1), steps))
    res[flag] = tokens.cpu()
a, b = res[0], res[1]
same_rows = int((a == b).all(dim=-1).sum())
print("rows identical: %d / %d; tokens identical: %.4f" % (same_rows, a.shape[0] if a.dim() == 2 else a.shape[0] * a.shape[1], float((a == b).float().mean())))
print("off:", a.reshape(-1, a.shape[-1])[0].tolist())
print("on :", b.reshape(-1, b.shape[-1])[0].tolist())
if want_trace:
    r = trace.cpu()[148 * per:].view(148, 64).double()
    t = trace.cpu()[:148 * per].view(148, -1)
    base1 = float(t[:, 11].double().median())  # layer 1 ln1.wait
    rn = ["x.polled", "x.issued", "mma.first_x", "mma.commit", "epi.ready", "epi.stored", "-", "-"]
    for kind, kn in enumerate(["qkv", "proj", "fc", "fc2"]):
        for k in range(6):
            col = r[:, kind * 8 + k]
            col = col[col > 0]
            if col.numel():
                print("  L1 %-4s %-12s med %7.2f max %7.2f min %7.2f" % (kn, rn[k], (float(col.median()) - base1) / 1e3, (float(col.max()) - base1) / 1e3, (float(col.min()) - base1) / 1e3))
    names_a = ["enter", "qkv folded", "new token", "b0 read", "b0 issued", "b1 read", "b1 issued", "b2 read", "b2 issued", "b3 read", "b3 issued", "batches done"]
    for c in (0, 70, 147):
        print("  CTA %3d attention unit of compute warp 1 (us after attn.wait): " % c + "  ".join("%s %.2f" % (names_a[k], (float(r[c, 32 + k]) - float(t[c, 11 + 2])) / 1e3) for k in range(12) if r[c, 32 + k] > 0))
    for c in (0, 70, 147):
        b0 = float(t[c, 11 + 3])
        f = lambda k: "%.2f" % ((float(r[c, k]) - b0) / 1e3) if r[c, k] > 0 else "-"
        print("  CTA %3d (us after its attn.wait): helper go %s, helpers done %s %s %s, compute warps done %s, thread 0 past helpers %s" % (
            c, f(48), f(45), f(46), f(47), " ".join(f(50 + w) for w in range(8)), f(44)))
    for c in ():
        print("  CTA %3d fc units: x issue  " % c + " ".join("%6.2f" % ((float(v) - base1) / 1e3) if v > 0 else "   -  " for v in r[c, 44:54]))
        print("                   w ready  " + " ".join("%6.2f" % ((float(v) - base1) / 1e3) if v > 0 else "   -  " for v in r[c, 54:64]))
        print("                   x landed " + " ".join("%6.2f" % ((float(v) - base1) / 1e3) if v > 0 else "   -  " for v in r[c, 32:42]))
    t0 = int(t[:, 0][t[:, 0] > 0].min())
    # per layer: [wait?] arrive ... program order per CTA: l=0: A A W A A W A A W A A ; l>0: W A A W A A W A A W A A
    names0 = ["ln1.arr", "qkv.arr", "attn.wait", "attn.arr", "proj.arr", "ln2.wait", "ln2.arr", "fc.arr", "gelu.wait", "gelu.arr", "fc2.arr"]
    names = ["ln1.wait"] + names0
    def show(layer):
        off = 0 if layer == 0 else 11 + 12 * (layer - 1)
        nm = names0 if layer == 0 else names
        base = float(t[:, off].double().median())
        print("layer %d (us, median / max over CTAs, relative to the layer's first stamp):" % layer)
        for k, n in enumerate(nm):
            col = t[:, off + k].double()
            print("  %-10s %7.2f %7.2f" % (n, (float(col.median()) - base) / 1e3, (float(col.max()) - base) / 1e3))
    show(0); show(1); show(24)
    last = 11 + 12 * (L - 1) + 1
    print("kernel span %.1f us" % ((int(t[:, last].max()) - t0) / 1e3))
