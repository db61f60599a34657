"""Per-complex fixed costs of a sampling call (matters for screening: one call per ligand): prepare_complex (pair-bias prepass),
schedule conditioning, CUDA-graph capture, against the 40 steps themselves.  B = 8 samples, Nt=256 / Na=2048."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from physdock_b200.dit import B200DiT
from physdock_b200.sampler import DiffusionSampler
from physdock_b200.synthetic import DiTDims, make_dit_state, make_complex
dev = torch.device("cuda")
dims = DiTDims.named("medium")
dit = B200DiT.from_state_dict(make_dit_state(dims, seed=0), dims, device=dev)
def T():
    torch.cuda.synchronize(); return time.perf_counter()
for rep in range(3):
    cx = {k: v.to(dev) for k, v in make_complex(256, 2048, dims, seed=10 + rep).items()}      # a NEW complex every time
    t0 = T()
    dit.prepare_complex(cx, cx["a"], cx["ap"], cx["s"], cx["z"]); dit._complex_sig = dit._complex_signature(cx, cx["a"], cx["ap"], cx["s"], cx["z"])
    t1 = T()
    smp = DiffusionSampler(dit, cx, cx["a"], cx["ap"], cx["s"], cx["z"], num_sample=8, steps=40, karras_noise_schedule_power=1000, align_ref_pos=False)
    t2 = T()
    smp.begin(); smp.step(0)
    t3 = T()
    for i in range(1, 40): smp.step(i)
    x = smp.x_next.cpu()
    t4 = T()
    print(f"rep {rep}: prepare_complex {1e3*(t1-t0):6.2f} ms | sampler init (schedule conditioning) {1e3*(t2-t1):6.2f} ms | first step (warm-up + graph capture) {1e3*(t3-t2):6.2f} ms | "
          f"39 steps + readback {1e3*(t4-t3):6.2f} ms | total {1e3*(t4-t0):6.2f} ms")

# eager launches (one C call enqueues the 121 kernels with PDL attributes) against the CUDA-graph replay, steady state
for B in (8, 16):
    cx = {k: v.to(dev) for k, v in make_complex(256, 2048, dims, seed=3).items()}
    for graph in (True, False, True, False):
        smp = DiffusionSampler(dit, cx, cx["a"], cx["ap"], cx["s"], cx["z"], num_sample=B, steps=40, karras_noise_schedule_power=1000,
                               align_ref_pos=False, use_cuda_graph=graph)
        smp.begin()
        for i in range(5): smp.step(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(40): smp.step(i)
        e1.record(); torch.cuda.synchronize()
        print(f"B={B} {'graph' if graph else 'eager'}: {e0.elapsed_time(e1) / 40:.3f} ms per step")
