import sys, ctypes as C
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ldso_b200 import capi, synth
win = synth.make_window(nF=8, pts_per_frame=250, seed=42)
ctx = capi.Context(win.w, win.h, win.levels); ctx.load_synth_window(win)
ctx.optimize_begin(); ctx.gn_iterations(0, 30); ctx.synchronize()
buf = (C.c_longlong*80)()
ctx.L.ldso_b200_debug_clocks(ctx.ctx, buf)
t = np.array(buf[:7]); print("K3 stamps delta cycles [stage-in wait | backup+solve | orthogonalise | xAd | step+refresh | stage-out]:", np.diff(t), "total", t[-1]-t[0])
print("K3 solve: sort+permute %d | factorisation %d | back-substitution %d cycles" % (buf[20] - buf[1], buf[21] - buf[20], buf[22] - buf[21]))
print("K3 solve front: [inputs staged -> solve entry | SVecI | rank sort | permuted copy] %s" % [int(buf[40] - buf[1]), int(buf[41] - buf[40]), int(buf[42] - buf[41]), int(buf[20] - buf[42])])
print("K3 block step 4 (row 16): panel %d | helper far update (thread 96) %d | barrier wait %d | near update %d | barrier wait %d" % tuple(buf[24:29]))
t = np.array(buf[64:72]); print("K1 stamps delta cycles:", np.diff(t), "total", t[-1]-t[0])
t = np.array(buf[72:76]); print("K2b diag-CTA stamps delta cycles:", np.diff(t))
t = np.array(buf[76:78]); print("K2b select-CTA cycles:", np.diff(t))
cap = 1024
sp = (C.c_longlong * (3 * cap))()
n = ctx.L.ldso_b200_debug_cta_spans(ctx.ctx, sp, cap)
a = np.array(sp[:3 * n]).reshape(n, 3)
t0 = a[:, 0].min()
st, en, sm = a[:, 0] - t0, a[:, 1] - t0, a[:, 2]
print(f"K1 CTA spans (ns, globaltimer): n={n} start p0/p50/p100 = {st.min()}/{int(np.median(st))}/{st.max()}  end p0/p50/p100 = {en.min()}/{int(np.median(en))}/{en.max()}")
dur = en - st
print(f"   duration p0/p50/p90/p100 = {dur.min()}/{int(np.median(dur))}/{int(np.percentile(dur, 90))}/{dur.max()}  distinct SMs {len(set(sm.tolist()))}  max CTAs/SM {np.bincount(sm.astype(int)).max()}")
order = np.argsort(en)[-6:]
print("   last finishers (cta, start, end, dur, sm):", [(int(i), int(st[i]), int(en[i]), int(dur[i]), int(sm[i])) for i in order])

# one-iteration wall-clock timeline (graph launch): K3 -> K1 -> K2a -> K2b, relative to K3's start
k3s, k3e, k2as, k2ae, k2bs, k2be = buf[14], buf[15], buf[16], buf[17], buf[18], buf[19]
k1s, k1e = int(a[:, 0].min()), int(a[:, 1].max())
print("timeline ns rel. K3 start: K3 [0, %d]  K1 [%d, %d]  K2a [%d, %d]  K2b [%d, %d]" %
      (k3e - k3s, k1s - k3s, k1e - k3s, k2as - k3s, k2ae - k3s, k2bs - k3s, k2be - k3s))

def timeline(tag):
    ctx.L.ldso_b200_debug_clocks(ctx.ctx, buf)
    n = ctx.L.ldso_b200_debug_cta_spans(ctx.ctx, sp, cap)
    a = np.array(sp[:3 * n]).reshape(n, 3)
    k3s = buf[14]
    print("%s timeline ns rel. K3 start: K3 [0, %d]  K1 [%d, %d]  K2a [%d, %d]  K2b [%d, %d]" %
          (tag, buf[15] - k3s, int(a[:, 0].min()) - k3s, int(a[:, 1].max()) - k3s, buf[16] - k3s, buf[17] - k3s, buf[18] - k3s, buf[19] - k3s))
    d = a[:, 1] - a[:, 0]
    print("   K1 CTA duration p50/p100 %d/%d ns; K3 stamps %s fact %d backsub %d" % (int(np.median(d)), d.max(), np.diff(np.array(buf[:7])), buf[21] - buf[20], buf[22] - buf[21]))

import torch
flush = torch.empty(192 * 1024 * 1024, dtype=torch.uint8, device="cuda")
for rep in range(3):
    flush.fill_(rep); torch.cuda.synchronize()
    ctx.gn_iterations(3, 1); ctx.synchronize()
    timeline("L2-cold")
ctx.gn_iterations(3, 3); ctx.synchronize()
timeline("L2-warm")
print("use_pdl/use_graph probe: launches per gn_iterations(3,1):", (lambda a: (ctx.gn_iterations(3, 1), ctx.launch_count() - a)[1])(ctx.launch_count()))
