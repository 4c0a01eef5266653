import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, ROOT + "/tests", ROOT + "/oracle", ROOT + "/tools"]
import numpy as np, synth, ilf_oracle as O
import vvcsoftware_vtm_b200 as v
K = ("y", "cb", "cr")
w, h, bd, cl = (int(a) for a in sys.argv[1:5])
rng = np.random.default_rng(int(sys.argv[5]))
ctu = 1 << cl
cw, ch = (w + ctu - 1) // ctu, (h + ctu - 1) // ctu
pic = synth.picture(rng, w, h, bd, "mix")
prm = synth.sao_params(rng, cw, ch, bd, p_off=0.2)
want = O.sao(pic, bd, bd, cl, prm)
with v.InLoopFilter(w, h, bd, bd, cl) as f:
    f.upload(0, *(pic[k] for k in K)); f.set_sao_params(0, prm); f.sao_process(0); got = f.download(0)
P = prm.view(synth.SAO_DT).reshape(-1)
for pi, k in enumerate(K):
    bad = np.argwhere(got[k] != want[k])
    print(k, "mismatches", len(bad))
    if not len(bad): continue
    sz = ctu >> (1 if pi else 0)
    from collections import Counter
    c = Counter()
    for y, x in bad:
        ci = (y // sz) * cw + (x // sz)
        c[(ci, int(P["type"][ci, pi]), int(P["avail"][ci]))] += 1
    print(" per (ctu,type,avail):", dict(c))
    print(" x%8:", Counter((bad[:, 1] % 8).tolist()), " y%8:", Counter((bad[:, 0] % 8).tolist()))
    print(" x in ctu:", sorted(Counter((bad[:, 1] % sz).tolist()).items())[:20])
    print(" y in ctu:", sorted(Counter((bad[:, 0] % sz).tolist()).items())[:20])
    for y, x in bad[:8]:
        print("  ", y, x, "in", pic[k][y, x], "got", got[k][y, x], "want", want[k][y, x])
