"""2-GPU check of the sample-sharded sampler (skipped on boxes with one GPU): the union over ranks equals the
single-GPU run bit for bit (ShardedRNG), and the NCCL all_gather returns every sample on every rank."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from physdock_b200.dit import B200DiT
from physdock_b200.sampler import sample_diffusion
from physdock_b200.sharding import sample_diffusion_sharded
from physdock_b200.synthetic import DiTDims, make_dit_state, make_complex
rank = int(os.environ["RANK"]); torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
dims = DiTDims.named("medium")
dit = B200DiT.from_state_dict(make_dit_state(dims, seed=0), dims, device=dev)
cx = {k: v.to(dev) for k, v in make_complex(24, 100, dims, seed=9).items()}
kw = dict(steps=4, karras_noise_schedule_power=1000, align_ref_pos=True)
x_all = sample_diffusion_sharded(dit, cx, cx["a"], cx["ap"], cx["s"], cx["z"], num_sample=5, seed=3, **kw)
torch.manual_seed(3)
x_one = sample_diffusion(dit, cx, cx["a"], cx["ap"], cx["s"], cx["z"], num_sample=5, **kw)
assert x_all.shape == (5, 100, 3)
assert torch.equal(x_all, x_one), float((x_all - x_one).abs().max())
# the round driver in sharded mode (one all_gather_into_tensor per round) == the single-process driver, bit for bit
from physdock_b200.rounds import run_rounds
from physdock_b200.sampler import PhysDockB200
from physdock_b200.synthetic import make_templates
model = PhysDockB200(dit, diffusion_conditioning=lambda b: (b["a"], b["ap"], b["s"], b["z"]))
lig = torch.nonzero(cx["is_ligand"][cx["atom_id_to_token_id"]].bool()).flatten()
i0, i1, i2 = int(lig[0]), int(lig[1]), int(lig[2])
accept = lambda x: float(torch.linalg.cross(x[i1] - x[i0], x[i2] - x[i0])[2]) > 0.0
rk = dict(num_augmentation_sample=3, max_samples=5, max_rounds=3, steps=4, physics_correction=True, conformers=make_templates(cx, 9),
          accept_fn=accept, ranking=False, seed=11)
a = run_rounds(model, cx, sharded=True, **rk)
b = run_rounds(model, cx, sharded=False, **rk)
assert len(a.rounds) == len(b.rounds) and a.n_accepted == b.n_accepted
assert torch.equal(a.accept_samples, b.accept_samples)
for ra, rb in zip(a.rounds, b.rounds):
    assert ra.pass_flags == rb.pass_flags and torch.equal(ra.x_pred, rb.x_pred)
dist.barrier(); dist.destroy_process_group()
print("rank", rank, "ok")
'''


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_gpu_sharded_sampling_matches_single_gpu(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", str(script), ROOT],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("ok") == 2
