"""Round-2 golden fixtures, generated FROM THE REAL REFERENCE (imported from /root/reference).  TEST INFRASTRUCTURE.
Run here (the reference tree does not exist on the GPU box):   python -m oracle.make_golden_r02 [trace40] [c1]

  trace40_nophys.npz / trace40_templates.npz
        FULL 40-step PhysDock.sample_diffusion traces (steps=40, rho=1000 exactly as redocking.py:39,53 run it) at
        Nt=64 / Na=512, B=2: all 29 stochastic steps (t_cur > 1) and the 11-step ODE tail; x_hat, t_hat, x_denoised per
        step + the recorded RNG tape (same layout as trace_*.npz of oracle/make_golden.py).
  c1_5sd5.npz
        BASELINE.json configs[0] on REAL data: `FeatureLoader.load` of demo/redocking/Posebusters_subset/5SD5_HWI_A_1.pkl.gz
        at crop_size=64 / atom_crop_size=512 (feature_loader.py:1004-1173; SURVEY.md Appendix A shims for rdkit), the
        reference trunk `DiffusionConditioning` with seeded random weights (params.pt is not shipped) -> a, ap, s, z
        (stored as fp16 to keep the fixture small; both the reference denoiser below and the kernels consume the SAME
        rounded values), the hot-path batch keys (real ragged chunk sizes, zero-size tokens, is_ligand layout), the
        reference AF3DiT output at 5 noise levels and a 4-sample 12-step `sample_diffusion` trace.
"""
import enum
import gzip
import os
import pickle
import sys
import types

import numpy as np
import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle.ref_import import REFERENCE_ROOT, build_reference_dit, import_reference, install_shims  # noqa: E402
from oracle.make_golden import npy, record_tape  # noqa: E402
from physdock_b200.synthetic import DiTDims, make_dit_state, make_complex, make_templates, checksum  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
T_LEVELS = [4608.0, 100.0, 10.0, 1.0, 0.2]
HOT_KEYS = ["atom_id_to_token_id", "token_id_to_chunk_sizes", "ap_mask", "z_mask", "a_mask", "is_ligand", "x_gt", "ref_pos"]


def ref_sampler(PhysDock, dit, cond):
    class RefSampler(PhysDock):     # the reference sampler with the trunk replaced by cached outputs
        def __init__(self):
            nn.Module.__init__(self)
            self.dit = dit
            self.diffusion_conditioning = cond
            self.sigma_data = 16.0
    return RefSampler()


def traced_sample(model, dit, batch, seed, **kw):
    steps = []
    orig_forward = dit.forward

    def wrapped(batch, x_hat, t_hat, a, ap, s, z):      # keyword names as model.py:153,221 passes them
        y = orig_forward(batch, x_hat, t_hat, a, ap, s, z)
        steps.append((x_hat.clone(), t_hat.clone(), y.clone()))
        return y

    dit.forward = wrapped
    try:
        x_final, tape = record_tape(lambda: model.sample_diffusion(batch, ref_mol=None, **kw), seed=seed)
    finally:
        dit.forward = orig_forward
    d = dict(n_tape=len(tape), n_steps=len(steps), x_final=npy(x_final))
    for i, t in enumerate(tape):
        d[f"tape_{i}"] = npy(t)
    for i, (xh, th, xd) in enumerate(steps):
        d[f"x_hat_{i}"], d[f"t_hat_{i}"], d[f"x_denoised_{i}"] = npy(xh), npy(th), npy(xd)
    return d


def make_trace40():
    dims = DiTDims.named("medium")
    PhysDock, _, _, _ = import_reference()
    dit = build_reference_dit("medium")
    sd = make_dit_state(dims, seed=0)
    dit.load_state_dict(sd)
    cx = make_complex(64, 512, dims, seed=1)
    model = ref_sampler(PhysDock, dit, lambda batch: (cx["a"], cx["ap"], cx["s"], cx["z"]))
    tmpl = make_templates(cx, 12)
    variants = {"nophys": dict(align_ref_pos=False),
                "templates": dict(align_ref_pos=True, ref_mol_poses=tmpl, mmff_gamma_0_factor=6.0)}
    for name, kw in variants.items():
        d = traced_sample(model, dit, cx, seed=321, num_sample=2, steps=40, karras_noise_schedule_power=1000, **kw)
        assert d["n_steps"] == 40
        d.update(sd_checksum=checksum(torch.cat([v.flatten() for v in sd.values()])), ap_checksum=checksum(cx["ap"]))
        np.savez_compressed(os.path.join(OUT, f"trace40_{name}.npz"), **d)
        print("trace40", name, "tape", d["n_tape"], os.path.getsize(os.path.join(OUT, f"trace40_{name}.npz")) // 1024, "KiB")


# ---------------------------------------------------------------------------------------------- real-data C1
def extend_rdkit_shims():
    """what FeatureLoader needs beyond oracle/ref_import.py's shims."""
    install_shims()
    rdchem, rdmolops, Chem = sys.modules["rdkit.Chem.rdchem"], sys.modules["rdkit.Chem.rdmolops"], sys.modules["rdkit.Chem"]

    class _AutoEnum:
        def __getattr__(self, k):
            return k
    for name in ("HybridizationType", "ChiralType", "BondType", "BondStereo", "BondDir"):
        setattr(rdchem, name, _AutoEnum())
    Chem.rdchem, Chem.rdmolops = rdchem, rdmolops
    Chem.MolFromSmarts = lambda s: None
    Chem.MolFromSmiles = lambda s: None
    Chem.Mol = object

    class _Stub:
        def __init__(self, *a, **k):
            pass

        def __setstate__(self, st):
            self._state = st

    class RdkitFreeUnpickler(pickle.Unpickler):
        def find_class(self, module, name):
            if module.startswith("rdkit"):
                return _Stub
            return super().find_class(module, name)

    def load_pkl(path, *a, **k):
        opener = gzip.open if str(path).endswith(".gz") else open
        with opener(path, "rb") as f:
            return RdkitFreeUnpickler(f).load()

    import PhysDock.utils.io_utils as io_utils
    io_utils.load_pkl = load_pkl
    import PhysDock.data.feature_loader as fl
    fl.load_pkl = load_pkl
    return fl


def make_c1():
    dims = DiTDims.named("medium")
    fl = extend_rdkit_shims()
    PhysDock, PhysDockConfig, _, _ = import_reference()
    loader = fl.FeatureLoader(
        msa_features_dir=os.path.join(REFERENCE_ROOT, "demo", "redocking", "features"),
        ccd_id_meta_data=os.path.join(REFERENCE_ROOT, "params", "ccd_id_meta_data.pkl.gz"),
        crop_size=64, atom_crop_size=512, inference_mode=True, infer_pocket_type="atom", infer_pocket_cutoff=6,
        infer_pocket_dist_type="ligand", infer_use_pocket=True, infer_use_key_res=True, key_res_random_mask_ratio=0.5,
        use_x_gt_ligand_as_ref_pos=False, num_recycles=2)
    import random
    random.seed(0); np.random.seed(0); torch.manual_seed(0)
    tensors, meta = loader.load(system_pkl_path=os.path.join(REFERENCE_ROOT, "demo", "redocking", "Posebusters_subset",
                                                             "5SD5_HWI_A_1.pkl.gz"))
    Nt, Na = tensors["s_mask"].shape[0], tensors["a_mask"].shape[0]
    print("5SD5_HWI_A_1 crop 64/512 ->", "Nt", Nt, "Na", Na, "ligand atoms",
          int(tensors["is_ligand"][tensors["atom_id_to_token_id"]].sum()),
          "zero-size tokens", int((tensors["token_id_to_chunk_sizes"] == 0).sum()))
    # reference trunk with seeded random weights (153 M parameters; params.pt is a Zenodo download)
    torch.manual_seed(2)
    model = PhysDock(PhysDockConfig(model_name="medium")).float().eval()
    sd = make_dit_state(dims, seed=0)
    model.dit.load_state_dict(sd)
    with torch.no_grad():
        a, ap, s, z = model.diffusion_conditioning(tensors)
    # random-init trunk outputs can be large; the fixture keeps them as fp16 (both sides consume the rounded values)
    scale = {k: float(v.abs().max()) for k, v in dict(a=a, ap=ap, s=s, z=z).items()}
    print("trunk output |max|:", scale)
    a, ap, s, z = (t.half().float() for t in (a, ap, s, z))
    assert all(torch.isfinite(t).all() for t in (a, ap, s, z))
    batch = {k: tensors[k] for k in HOT_KEYS}
    d = {f"batch_{k}": npy(v) for k, v in batch.items()}
    # what the physics parameter source reads (physics.field_from_features): elements and token bonds
    d["extra_ref_element"] = npy(tensors["ref_feat"][:, 4:132].argmax(-1) + 1)
    d["extra_token_bonds"] = npy(tensors["token_bonds"])
    d.update(a=npy(a.half()), ap=npy(ap.half()), s=npy(s.half()), z=npy(z.half()), Nt=Nt, Na=Na,
             sd_checksum=checksum(torch.cat([v.flatten() for v in sd.values()])))
    dit = model.dit
    g = torch.Generator().manual_seed(17)
    for t in T_LEVELS:
        x_hat = torch.randn(4, Na, 3, generator=g) * (t ** 2 + 100) ** 0.5
        t_hat = torch.full([4], t)
        with torch.no_grad():
            y = dit(batch, x_hat, t_hat, a, ap, s, z)
        d[f"x_hat_{t}"], d[f"x_denoised_{t}"] = npy(x_hat), npy(y)
    sampler = ref_sampler(PhysDock, dit, lambda b: (a, ap, s, z))
    tr = traced_sample(sampler, dit, batch, seed=99, num_sample=4, steps=12, karras_noise_schedule_power=1000,
                       align_ref_pos=True)
    d.update({f"trace_{k}": v for k, v in tr.items()})
    path = os.path.join(OUT, "c1_5sd5.npz")
    np.savez_compressed(path, **d)
    print("c1_5sd5.npz", os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    what = sys.argv[1:] or ["trace40", "c1"]
    if "trace40" in what:
        make_trace40()
    if "c1" in what:
        make_c1()
