"""Per-role clock64 timeline of one CTA pair of the default attention kernel (trace build, variant 0x804).
Prints, per KV tile, when each hand-shake of the S -> softmax -> P -> PV chain happened (cycles since the pair's first stamp)."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent)); sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))
import kernels as K
from qie_b200 import _lib as L
dev = "cuda:0"
H = 24
s = K.seq(1, 8192, 256)
qkv = torch.randn(K.rows(s), 3 * H * 128, device=dev).bfloat16()
NJ, NEV, ROLES = 32, 8, 11
buf = torch.zeros(2 * ROLES * NJ * NEV, dtype=torch.int64, device=dev)
L.check(L.lib().qie_attn_set_trace(L.ptr(buf)))
variant = int(sys.argv[1], 0) if len(sys.argv) > 1 else 0x804
for _ in range(3):
    K.attn(s, qkv, H, variant)
torch.cuda.synchronize()
t = buf.cpu().view(2, ROLES, NJ, NEV)
nz = t[t > 0]
t0 = [int(t[c][t[c] > 0].min()) for c in range(2)]
print("clock origin per CTA (the two SMs of a TPC have separate counters):", t0, "skew", t0[1] - t0[0])
sm_ev = ["s_full ok", "S in regs", "max done", "exp done", "pv_done ok", "P stored", "p_full arrived", "loop top"]
names = {0: ("S-issuer", ["s_free ok", "k_full ok", "issued"]), 1: ("PV-issuer", ["v_full ok", "p_full ok", "issued", "S(j) complete", "PV(j) complete"]),
         10: ("producer", ["k_empty ok", "v_empty ok"])}
for wg in range(2):
    for q in range(4):
        names[2 + wg * 4 + q] = (f"softmax wg{wg} quad{q}", sm_ev)
J0, J1 = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (12, 22)
for c in range(2):
    print(f"==== CTA rank {c} (cycles since its first stamp)")
    for role in sorted(names):
        nm, evs = names[role]
        print(f"-- {nm}: " + " | ".join(evs))
        for j in range(J0, J1):
            row = t[c, role, j, :len(evs)]
            if (row > 0).any():
                print(f"   j={j:2d} " + " ".join(f"{int(v) - t0[c]:8d}" if v > 0 else "       -" for v in row))
# steady-state summary on the leader: period of S issue and of PV issue
for role, ev in ((0, 2), (1, 2)):
    x = t[0, role, 8:NJ, ev]
    x = x[x > 0]
    if len(x) > 2:
        d = (x[1:] - x[:-1]).float()
        print(f"role {role} issue period over tiles 8..: mean {d.mean():.0f} cycles, min {d.min():.0f}, max {d.max():.0f}")
