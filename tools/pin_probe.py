import sys, time, torch
sys.path.insert(0, '.')
from oracle.ref_import import build_reference_dit, import_reference
from oracle import physdock_oracle as O
from physdock_b200.synthetic import DiTDims, make_dit_state, make_complex, dit_param_shapes, make_templates
import torch.nn as nn

dims = DiTDims.named("medium")
dit = build_reference_dit("medium")
ref_sd = dit.state_dict()
shapes = dit_param_shapes(dims)
assert list(ref_sd.keys()) == list(shapes.keys()), "key order mismatch"
assert all(tuple(ref_sd[k].shape) == shapes[k] for k in shapes)
sd = make_dit_state(dims, seed=0)
dit.load_state_dict(sd)
torch.set_num_threads(8)
cx = make_complex(64, 512, dims, seed=1)
B = 4
g = torch.Generator().manual_seed(3)
for t in [4608.0, 100.0, 10.0, 1.0, 0.2]:
    x_hat = torch.randn(B, 512, 3, generator=g) * (t**2 + 100)**0.5
    t_hat = torch.full([B], t)
    with torch.no_grad():
        y_ref = dit(cx, x_hat, t_hat, cx["a"], cx["ap"], cx["s"], cx["z"])
        y_or = O.af3dit_forward(sd, cx, x_hat, t_hat, cx["a"], cx["ap"], cx["s"], cx["z"])
        sd64 = {k: v.double() for k, v in sd.items()}
        cx64 = {k: (v.double() if v.is_floating_point() else v) for k, v in cx.items()}
        y64 = O.af3dit_forward(sd64, cx64, x_hat.double(), t_hat.double(), cx64["a"], cx64["ap"], cx64["s"], cx64["z"])
    print(t, "oracle==ref:", torch.equal(y_ref, y_or), float((y_ref-y_or).abs().max()),
          "rmsd fp32 vs fp64:", float(O.rmsd(y_or, y64).max()), "|y|", float(y_or.abs().mean()))

# sampler pin
PhysDock, PhysDockConfig, AF3DiT, tu = import_reference()
class RefSampler(PhysDock):
    def __init__(self, dit, cond):
        nn.Module.__init__(self)
        self.dit = dit
        self.diffusion_conditioning = cond
        self.sigma_data = 16.0
m = RefSampler(dit, lambda batch: (cx["a"], cx["ap"], cx["s"], cx["z"]))
tmpl = make_templates(cx, 12)
for kw in [dict(align_ref_pos=False), dict(align_ref_pos=True), dict(align_ref_pos=True, ref_mol_poses=tmpl, mmff_gamma_0_factor=6.0)]:
    torch.manual_seed(123)
    t0 = time.time()
    x_ref = m.sample_diffusion(cx, num_sample=B, steps=8, ref_mol=None, karras_noise_schedule_power=1000, **kw)
    torch.manual_seed(123)
    x_or = O.sample_diffusion(sd, cx, cx["a"], cx["ap"], cx["s"], cx["z"], num_sample=B, steps=8, karras_noise_schedule_power=1000, **kw)
    print(kw.keys(), "sampler equal:", torch.equal(x_ref, x_or), float((x_ref-x_or).abs().max()), time.time()-t0)
