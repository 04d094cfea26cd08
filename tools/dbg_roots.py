import sys, numpy as np
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from gpu_util import ctx, synth, vb
import oracle
audio = synth.utterance(9, 16000, seconds=2.0)
c = ctx()
N, hop, p = 400, 160, 12
F = c.n_frames_of(audio.size, N, hop)
d = c.to_device(audio)
_, ac, _ = c.lpc(c.frames(d.ptr, F, N, hop, vb.WINDOW_HANN_SYMMETRIC), p)
a = c.lpc_to_resonances(ac, p, True, 16000.0, precision=0, want_roots=True)
b = c.lpc_to_resonances(ac, p, True, 16000.0, precision=1, want_roots=True)
na, nb = a["n_res"].to_host(), b["n_res"].to_host()
bad = np.nonzero(na != nb)[0]
print("mismatch frames", bad[:10], len(bad), "of", F)
ra, rb = a["roots"].to_host(), b["roots"].to_host()
resa, resb = a["resonances"].to_host(), b["resonances"].to_host()
for f in bad[:3]:
    print("frame", f, "n", na[f], nb[f])
    print(" roots f32+polish:", np.round(ra[f,:,0]+1j*ra[f,:,1], 6))
    print(" roots f64       :", np.round(rb[f,:,0]+1j*rb[f,:,1], 6))
    print(" res a", resa[f,:8,0]); print(" res b", resb[f,:8,0])
    st, ref, _ = oracle.find_roots_mut(ac.to_host()[f][::-1].astype(np.complex128))
    print(" oracle:", np.round(ref[:p], 6))
print("max res diff on agreeing frames", np.max(np.abs(resa[na==nb]-resb[na==nb])))
import os
os.environ["VBX_ROOTS_KERNEL"] = "u"
b2 = c.lpc_to_resonances(ac, p, True, 16000.0, precision=1, want_roots=True)
rb2 = b2["roots"].to_host()
dif = np.abs(rb - rb2).max(axis=(1, 2))
print("frames where new f64 != old f64 roots:", np.count_nonzero(dif > 1e-9), "of", F)
f = int(np.argmax(dif))
print("worst frame", f, "\n new", np.round(rb[f,:,0]+1j*rb[f,:,1], 5), "\n old", np.round(rb2[f,:,0]+1j*rb2[f,:,1], 5))
