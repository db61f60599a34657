"""BASELINE.json configs[3] (screening, 8 samples per ligand) on one GPU: ligands per second with the reference's serial
hand-off against physdock_b200.screen.screen_ligands (next ligand's staging + trunk overlapped with the current ligand's
sampling).  Trunk = the UNMODIFIED reference `DiffusionConditioning` from baseline/_ref with random weights (PyTorch eager,
TF32 as the reference sets it, model.py:5); features = synthetic tensors of FeatureLoader's shapes (tools/time_trunk.py).
    python tools/time_screening.py [n_ligands Nt Na]"""
import os, sys, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
os.environ.setdefault("PHYSDOCK_REFERENCE", os.path.join(ROOT, "baseline", "_ref"))
from oracle.ref_import import import_reference
from time_trunk import synthetic_features
from physdock_b200.dit import B200DiT
from physdock_b200.sampler import PhysDockB200
from physdock_b200.screen import screen_ligands
from physdock_b200.synthetic import DiTDims, make_dit_state

n_lig, Nt, Na = (int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (8, 256, 2048)
dev = torch.device("cuda")
PhysDock, PhysDockConfig, _, _ = import_reference()
torch.set_float32_matmul_precision("high")          # as PhysDock/models/model.py:5
torch.manual_seed(0)
ref = PhysDock(PhysDockConfig(model_name="medium")).float().eval().to(dev)
dims = DiTDims.named("medium")
dit = B200DiT.from_state_dict(make_dit_state(dims, seed=0), dims, device=dev)


def trunk(batch):
    with torch.inference_mode():
        return tuple(t.float() for t in ref.diffusion_conditioning(batch))


model = PhysDockB200(dit, diffusion_conditioning=trunk)
ligands = list(range(n_lig))
featurise = lambda i: synthetic_features(Nt, Na, seed=i)          # noqa: E731  (CPU tensors, like FeatureLoader.load)
for overlap in (False, True, False, True):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    out = screen_ligands(model, ligands, featurise, num_sample=8, steps=40, overlap=overlap, gather=False, align_ref_pos=False)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    assert all(o is not None and torch.isfinite(o).all() for o in out)
    print(f"{'overlapped (pinned staging + trunk prefetch)' if overlap else 'serial (reference order)            '}: "
          f"{n_lig} ligands x 8 samples x 40 steps at Nt={Nt}/Na={Na}: {dt * 1e3 / n_lig:7.1f} ms per ligand = {n_lig / dt:5.2f} ligands/s per GPU")
