"""Which of the three split-fp16 products of the attention contractions can be dropped inside the 1e-3 A budget?  (CPU; emulates
the formats inside the oracle exactly like tools/precision_probe.py.)   python tools/precision_probe_terms.py [Nt Na]
Result at 64/512 (RMSD of x_denoised vs fp64, worst of t_hat = 4608, 100): all three products 4.9e-5 / 5.1e-5 A; without
P_lo.V_hi 1.4e-3; without P_hi.V_lo 2.4e-3; without q_lo.k_hi 1.9e-3; without q_hi.k_lo 2.0e-3 -> every product is needed."""
import sys
src = open(__file__.replace("precision_probe_terms.py", "precision_probe.py")).read().split('SDPA["fp16x2 QK, fp16x2 PV"]')[0]
exec(src)
three = [(0, 0), (0, 1), (1, 0)]
V = {
    "QK3 PV3": (three, three),
    "QK3 PV hh+hl (drop Plo.Vhi)": (three, [(0, 0), (0, 1)]),
    "QK3 PV hh+lh (drop Phi.Vlo)": (three, [(0, 0), (1, 0)]),
    "QK3 PV hh": (three, [(0, 0)]),
    "QK hh+hl (drop qlo.khi) PV3": ([(0, 0), (0, 1)], three),
    "QK hh+lh (drop qhi.klo) PV3": ([(0, 0), (1, 0)], three),
    "QK hh PV3": ([(0, 0)], three),
}
print("attention variant (linears fp16x2(3)) | rmsd vs fp64 at t=4608,100,10,1,0.2")
for name, (qt, pt) in V.items():
    SDPA[name] = make_sdpa(fp16, 2, qt, fp16, 2, pt)
    run("fp16x2(3)", name)
