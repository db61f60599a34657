"""Runs and times the pair-energy physics kernels at the benchmark shape (B=16, Na=2048, 32 ligand atoms) and with
rows = all atoms (the O(Na^2) case); also the target of the ncu capture profiles/r01_physics_ncu.txt."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from physdock_b200.physics import PairEnergyField
from physdock_b200.synthetic import make_ligand_field
dev = torch.device("cuda")
B, Na, n_lig = 16, 2048, 32
f = make_ligand_field(Na, n_lig, seed=0)
x = (f["x0"][None] + 0.3 * torch.randn(B, Na, 3)).to(dev)
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for _ in range(n): fn()
    gr.replay(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); gr.replay(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / n * 1e3)
    return best
for name, rows in (("ligand rows (32)", f["rows"]), ("all rows (2048)", None)):
    fld = PairEnergyField(f["x_exists"].to(dev), f["sigma"], f["eps"], f["partner"], f["partner_r0"], f["partner_k"], rows=rows)
    n_rows = fld.n_rows
    t = timeit(lambda: fld.energy_grad(x))
    pairs = B * n_rows * Na
    print(f"energy+grad, {name}: {t:7.1f} us  = {pairs / t / 1e3:.1f} G pairs/s")
    if rows is not None:
        t = timeit(lambda: fld.descend(x, iters=5))
        print(f"descend x5,  {name}: {t:7.1f} us")
