"""Runs a few sampler steps at another BASELINE.json configuration and prints ms per step (sanity + context numbers).
   python tools/run_config.py C3   -> crop 384 / atom crop 3072, 64 samples, physics guidance on (configs[2])
   python tools/run_config.py C1   -> crop 64 / atom crop 512, 4 samples, physics off (configs[0])
   python tools/run_config.py C5   -> crop 256 / atom crop 2048, 5 samples per GPU (40 samples over 8 GPUs, configs[4])"""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from physdock_b200.dit import B200DiT
from physdock_b200.sampler import DiffusionSampler
from physdock_b200.synthetic import DiTDims, make_dit_state, make_complex, make_templates, make_ligand_field
from physdock_b200.physics import PairEnergyField
cfg = sys.argv[1] if len(sys.argv) > 1 else "C3"
Nt, Na, B, physics = {"C1": (64, 512, 4, False), "C3": (384, 3072, 64, True), "C5": (256, 2048, 5, False)}[cfg]
if len(sys.argv) > 2: B = int(sys.argv[2])            # optional batch override: python tools/run_config.py C5 8
dev = torch.device("cuda")
dims = DiTDims.named("medium")
dit = B200DiT.from_state_dict(make_dit_state(dims, seed=0), dims, device=dev)
cx = {k: v.to(dev) for k, v in make_complex(Nt, Na, dims, seed=1).items()}
kw = dict(align_ref_pos=False)
if physics:
    n_lig = int(cx["is_ligand"][cx["atom_id_to_token_id"]].sum())
    f = make_ligand_field(Na, n_lig, seed=2, missing=False)
    field = PairEnergyField(cx["a_mask"], f["sigma"], f["eps"], f["partner"], f["partner_r0"], f["partner_k"], rows=f["rows"])
    kw = dict(align_ref_pos=True, ref_mol_poses=make_templates(cx, 40), mmff_gamma_0_factor=6.0, physics_field=field, mmff_iters=5)
torch.manual_seed(0)
smp = DiffusionSampler(dit, cx, cx["a"], cx["ap"], cx["s"], cx["z"], num_sample=B, steps=40, karras_noise_schedule_power=1000, use_cuda_graph=True, **kw)
smp.begin()
steps = list(range(3)) + list(range(30, 33))           # early (stochastic, template projection) and late (descent) steps
for i in steps: smp.step(i)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for rep in range(3):
    for i in steps: x = smp.step(i)
e1.record(); torch.cuda.synchronize()
assert torch.isfinite(x).all()
ms = e0.elapsed_time(e1) / (3 * len(steps))
print(f"{cfg}: Nt={Nt} Na={Na} B={B} physics={'on' if physics else 'off'}: {ms:.3f} ms per step = {B / ms * 1e3:.0f} sample-steps/s, "
      f"peak memory {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB")
