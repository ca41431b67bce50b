"""(Needs a library built with `python tools/build.py --tuning`: the timeline stamps are compiled out otherwise.)
Timeline of the GEMMs inside one replayed decode step (GPT2-XL, B=64): per launch, first CTA entry / last exit."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import clipcap_b200 as cc
from clipcap_b200 import synthetic
B = 64
cfg = cc.EngineConfig(max_images=B, max_beam=1, max_ctx=80)
eng = cc.Engine(cfg)
synthetic.load_synthetic(eng)
torch.cuda.empty_cache()
images = synthetic.synthetic_images(B, cfg, device="cuda")
LAUNCHES, CTAS = 1024, 512
trace = torch.zeros(LAUNCHES * CTAS * 8, dtype=torch.int64, device="cuda")
eng.lib.ccb_debug_gemm_trace(eng._h, C.c_void_p(trace.data_ptr()), CTAS * 8, LAUNCHES)
p = eng.gen_params("greedy", 8, stop_token=-1, max_stops=0)
eng.caption_images(images, p)
torch.cuda.synchronize()
t = trace.cpu().view(LAUNCHES, CTAS, 8)
rows = []
for n in range(LAUNCHES):
    m = t[n, :, 0] > 0
    if m.any():
        rows.append((int(t[n, m, 0].min()), int(t[n, m, 7].max()), int(m.sum()), int(t[n, m, 2].median()), int(t[n, m, 4].median()), n))
rows.sort()
# the decode step's launches are the last 48*4+1
dec = rows[-193:]
t0 = dec[0][0]
print("decode step: %d GEMM launches, span %.1f us" % (len(dec), (dec[-1][1] - t0) / 1e3))
prev_end = None
for i, (a, b, n, ft, acc, _) in enumerate(dec[:14]):
    gap = (a - prev_end) / 1e3 if prev_end else 0.0
    print("launch %3d: ctas %3d start %8.2f  first_tile +%5.2f  acc_ready +%5.2f  dur %6.2f us   gap before %6.2f us" % (
        i, n, (a - t0) / 1e3, (ft - a) / 1e3, (acc - a) / 1e3, (b - a) / 1e3, gap))
    prev_end = b
durs = [(b - a) / 1e3 for a, b, *_ in dec]
gaps = [(dec[i + 1][0] - dec[i][1]) / 1e3 for i in range(len(dec) - 1)]
print("sum of GEMM durations %.1f us, sum of gaps %.1f us" % (sum(durs), sum(gaps)))
for k, nm in enumerate(["qkv", "proj", "fc", "fc2"]):
    dd = [durs[i] for i in range(k, 192, 4)]
    gg = [gaps[i] for i in range(k, 191, 4)]
    print("  %-5s dur mean %.2f  gap after mean %.2f" % (nm, sum(dd) / len(dd), sum(gg) / len(gg)))

names = ["entry", "setup", "first_tile", "mma_issued", "acc_ready", "reduce_go", "epi_done", "exit"]
for i in range(4, 8):
    a, b, n, ft, acc, slot = dec[i]
    tt = t[slot]
    m = tt[:, 0] > 0
    print("launch %d (%d CTAs):" % (i, int(m.sum())), end=" ")
    for k in range(8):
        col = tt[m, k]
        col = col[col > 0]
        if col.numel():
            print("%s %.2f/%.2f" % (names[k], (float(col.median()) - a) / 1e3, (float(col.max()) - a) / 1e3), end="  ")
    print()
