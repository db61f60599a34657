#!/usr/bin/env python
"""Benchmark of the reverse-diffusion sampling step (BASELINE.json metric: denoising steps/sec at
crop_size=256 / atom_crop_size=2048).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One *step* = one iteration of PhysDock.sample_diffusion's loop (reference models/model.py:211-281) for a
batch of B samples of one complex: centre_random_augmentation + diffuse + AF3DiT + Euler update.
Workload = BASELINE.json configs[1]: redocking round 0 (align_ref_pos=False, redocking.py:284-299),
Nt=256 tokens, Na=2048 atoms, B=16 samples per GPU, 40-step rho=1000 schedule, synthetic trunk outputs and
weights from seeds (physdock_b200/synthetic.py).  Samples shard across GPUs with no
data-path collective (weak scaling); one all_gather collects the final coordinates after the timed loop.

Prints ONE JSON line (rank 0).  value = sample-steps/s with inputs resident in HBM (median of REPEATS timed
regions of K steps each); e2e = the same step with per-step inputs copied from pinned host memory and the
result read back.  Before anything is timed, rank 0 checks the benchmarked configuration itself (B=16, same
sampler object, same CUDA-graph path) against the CPU oracle ("parity" in the JSON line).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NT, NA, B_PER_GPU, SCHED_STEPS, RHO = 256, 2048, 16, 40, 1000
REPEATS = 5
C5_SAMPLES = 40          # BASELINE.json configs[4]: 40 samples of one complex spread over the GPUs (strong scaling)
METRIC = "denoising sample-steps/sec at crop=256, atom_crop=2048 (B=16 samples per GPU)"
UNIT = "sample-steps/s"
WORKLOAD = ("redocking Posebusters_subset, crop_size=256 atom_crop_size=2048, 16 samples, 1xB200 "
            "(BASELINE.json configs[1]); synthetic Nt=256 Na=2048, round 0: align_ref_pos=False, ref_mol=None")
ATTN_PROFILE = os.path.join("profiles", "r02_attn_atom_ncu.txt")


def flops_per_sample_step(Nt, Na, c_a=128, c_s=512, Ha=4, Hs=16, hid_a=384, hid_s=1408, na=3, nt=12):
    """F(Nt,Na) = 40.44 GFLOP at (256, 2048)."""
    atom = 8 * Na * c_a ** 2 + 4 * Ha * Na ** 2 * 32 + 6 * Na * c_a * hid_a
    tok = 8 * Nt * c_s ** 2 + 4 * Hs * Nt ** 2 * 32 + 6 * Nt * c_s * hid_s
    return 2 * na * atom + nt * tok + 2 * Na * c_a * c_s + 2 * Nt * c_s * c_a


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d["bf16_tflops_sustained"],
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm_gbs=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def cpu_oracle_steps(n_timed, n_warm, B_cpu, threads):
    """The reference's own CPU path for one step (oracle port, bit-identical to the reference on CPU:
    tests/test_oracle_pin.py), timed on the host cores: n_warm untimed + n_timed timed steps of the same 40-step
    schedule.  Returns (sample-steps/s, mean seconds per step)."""
    from oracle import physdock_oracle as O
    from physdock_b200.synthetic import DiTDims, make_dit_state, make_complex
    torch.set_num_threads(threads)
    dims = DiTDims.named("medium")
    sd = make_dit_state(dims, seed=0)
    cx = make_complex(NT, NA, dims, seed=1)
    sig = O.karras_noise_schedule(SCHED_STEPS, p=RHO)
    rng = O.TorchRNG("cpu")
    x = sig[0] * rng.normal((B_cpu, NA, 3))
    times = []
    with torch.no_grad():
        for k in range(n_warm + n_timed):
            i = k % SCHED_STEPS
            t0 = time.perf_counter()
            t_cur, t_next = sig[i], sig[i + 1]
            u = torch.stack([rng.rand((B_cpu,)) for _ in range(4)], -1)
            x_cur = O.centre_random_augmentation(x, cx["a_mask"], u, rng.normal((B_cpu, 3)))
            if t_cur > 1.0:                       # model.py:213-220
                t_hat = torch.full([B_cpu], float(t_cur * 1.8))
                x_hat, eta = O.diffuse(x_cur, t_hat, t_cur, rng.normal(x_cur.shape), 1.003), 1.5
            else:
                t_hat, x_hat, eta = torch.full([B_cpu], float(t_cur)), x_cur, 1.0
            x_den = O.af3dit_forward(sd, cx, x_hat, t_hat, cx["a"], cx["ap"], cx["s"], cx["z"])
            x = O.euler_update(x_hat, (x_hat - x_den) / t_hat[:, None, None], t_hat, t_next, eta)
            times.append(time.perf_counter() - t0)
    sec = statistics.mean(times[n_warm:])
    return B_cpu / sec, sec


def run_reference(args, rank):
    """--impl reference: the reference's CPU implementation of the step (oracle port, bit-identical to the reference:
    /root/reference does not exist on the GPU box), all host threads, the SAME configuration as the ours arm:
    B=16 samples, W warm-up + K timed steps of the 40-step rho=1000 schedule."""
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    W, K, B = max(1, args.warmup), max(1, args.steps), args.samples
    rate, sec = cpu_oracle_steps(K, W, B, threads)
    sample = (f"{K} timed steps (+{W} warm-up) of B={B} samples at Nt={NT}/Na={NA}, fp32 PyTorch CPU oracle "
              f"(bit-identical to the reference on CPU), {threads} threads")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": K,
        "warmup": W, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "Nt": NT, "Na": NA, "samples_per_gpu": B, "schedule": "40 steps rho=1000"},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def attention_dram_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the attention kernel, from the committed
    `ncu --set full` capture of the same shape (NOT measured in this run: a bench value is never taken under ncu)."""
    for rel in (ATTN_PROFILE, os.path.join("profiles", "r01_attn_atom_ncu.txt")):
        try:
            tot = 0.0
            for line in open(os.path.join(ROOT, rel)):
                f = line.split()
                if f and f[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                    tot += float(f[1]) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}[f[2]]
            if tot:
                return tot, rel
        except (OSError, ValueError, KeyError, IndexError):
            continue
    return None, None


def time_attention_kernel(dit, B, rounds=2):
    """Average launch duration of the dominant kernel (atom pair-bias attention) as the step runs it: launched back to
    back from a CUDA graph with Programmatic Dependent Launch, cycling the 6 cached bias blocks of the prepared complex
    (67 MB each, 403 MB > L2: no launch finds its bias in L2); CUDA events around the replay on the launching stream.
    Also returns the duration with an event recorded after EVERY launch (cold start of each launch exposed; round 1's
    method), reported as `launch_ms_isolated`."""
    from physdock_b200 import _lib
    lib = _lib.load()
    dev = torch.device("cuda", torch.cuda.current_device())
    H, S = 4, dit._complex_keep["Na"]
    S_pad = int(lib.pdk_pad_len(S))
    g = torch.Generator(device=dev).manual_seed(0)

    def planes(scale):            # [B,H,S_pad,64] rows [hi 32 | lo 32] of an fp32 tensor, as the QKV GEMM epilogue writes them
        x = torch.randn(B, H, S_pad, 32, generator=g, device=dev) * scale
        hi = x.half()
        return torch.cat([hi, (x - hi.float()).half()], -1).contiguous()
    # unit-RMS q and k (the model RMS-normalises both), q pre-scaled by log2(e) / sqrt(32) as the epilogue does
    q, k, v = planes(1.4427 / 32 ** 0.5), planes(1.0), planes(1.0)
    oh = torch.empty(B * S_pad, H * 32, dtype=torch.float16, device=dev)
    ol = torch.empty_like(oh)
    bias = dit._complex_keep["bias_a"].view(-1, H, S_pad, S_pad)
    n_blocks = bias.shape[0]

    def launch(l):
        _lib.check(lib.pdk_op_attention(q.data_ptr(), k.data_ptr(), v.data_ptr(), bias[l % n_blocks].data_ptr(),
                                        oh.data_ptr(), ol.data_ptr(), B, H, S_pad, _lib.stream_ptr(dev)), "pdk_op_attention")
    for l in range(3):
        launch(l)
    torch.cuda.synchronize()
    n = rounds * n_blocks
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        for i in range(n):
            launch(i)
    graph.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = []
    for _ in range(5):
        e0.record()
        graph.replay()
        e1.record()
        torch.cuda.synchronize()
        reps.append(e0.elapsed_time(e1) / n)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
    ev[0].record()
    for i in range(n):
        launch(i)
        ev[i + 1].record()
    torch.cuda.synchronize()
    isolated = statistics.mean(ev[i].elapsed_time(ev[i + 1]) for i in range(n))
    flops = B * H * 4 * S * S * 32          # QK^T + PV, 2 flop per MAC (algorithmic, not the 3x split)
    return statistics.median(reps) * 1e-3, flops, isolated * 1e-3


def parity_gate(smp, sd, cx_cpu, step_ids):
    """Teacher-forced parity of the BENCHMARKED configuration (same sampler object, B, CUDA-graph path): after
    `smp.step(i)` the x_hat the kernels consumed is fed to the CPU oracle (bit-identical to the reference) and
    x_denoised / x_next are compared per sample.  Bar: RMSD < 1e-3 A (BASELINE.json north_star)."""
    from oracle import physdock_oracle as O           # checker only
    worst_den, worst_next, per_step = 0.0, 0.0, []
    smp.begin()
    with torch.no_grad():
        for i in range(max(step_ids) + 1):
            smp.step(i)
            if i not in step_ids:
                continue
            t_cur, t_next, t_hat, stochastic, _ = smp.schedule(i)
            x_hat, x_den, x_next = smp.x_hat.cpu(), smp.x_den.cpu(), smp.x_next.cpu()
            th = torch.full([smp.B], float(t_hat))
            want_den = O.af3dit_forward(sd, cx_cpu, x_hat, th, cx_cpu["a"], cx_cpu["ap"], cx_cpu["s"], cx_cpu["z"])
            want_next = O.euler_update(x_hat, (x_hat - want_den) / th[:, None, None], th, t_next, 1.5 if stochastic else 1.0)
            r_den, r_next = float(O.rmsd(x_den, want_den).max()), float(O.rmsd(x_next, want_next).max())
            per_step.append({"step": i, "t_hat": float(t_hat), "x_denoised_rmsd": r_den, "x_next_rmsd": r_next})
            worst_den, worst_next = max(worst_den, r_den), max(worst_next, r_next)
    out = {"x_denoised_rmsd": worst_den, "x_next_rmsd": worst_next, "B": smp.B, "Nt": int(cx_cpu["s"].shape[0]),
           "Na": int(cx_cpu["a"].shape[0]), "tolerance_A": 1e-3,
           "oracle": "oracle/physdock_oracle.py on the host (fp32, bit-identical to the reference)", "steps": per_step,
           "path": "DiffusionSampler.step -> CUDA graph of pdk_dit_denoise_cond (the timed path)"}
    assert worst_den < 1e-3 and worst_next < 1e-3, f"parity gate failed: {out}"
    return out


def other_config(dit, sd, dims, dev, Nt, Na, Bc, W, physics, what):
    """One of the other BASELINE.json configurations on this GPU (context, not the headline): a 40-step run of the device-resident
    step at that shape; without physics also the parity gate of that shape (one stochastic step against the CPU oracle)."""
    from physdock_b200.sampler import DiffusionSampler
    from physdock_b200.synthetic import make_complex
    cx_cpu = make_complex(Nt, Na, dims, seed=1)
    cx = {k: v.to(dev) for k, v in cx_cpu.items()}
    kw = dict(align_ref_pos=False)
    if physics:
        from physdock_b200.synthetic import make_templates, make_ligand_field
        from physdock_b200.physics import PairEnergyField
        n_lig = int(cx["is_ligand"][cx["atom_id_to_token_id"]].sum())
        f = make_ligand_field(Na, n_lig, seed=2, missing=False)
        field = PairEnergyField(cx["a_mask"], f["sigma"], f["eps"], f["partner"], f["partner_r0"], f["partner_k"], rows=f["rows"])
        kw = dict(align_ref_pos=True, ref_mol_poses=make_templates(cx, 40), mmff_gamma_0_factor=6.0, physics_field=field, mmff_iters=5)
    torch.manual_seed(5)
    smp = DiffusionSampler(dit, cx, cx["a"], cx["ap"], cx["s"], cx["z"], num_sample=Bc, steps=SCHED_STEPS,
                           karras_noise_schedule_power=RHO, use_cuda_graph=True, **kw)
    out = {"workload": what, "Nt": Nt, "Na": Na, "samples": Bc}
    if physics:
        out["parity"] = "unpinned (pair-energy physics backend has no reference arithmetic)"
    else:
        par = parity_gate(smp, sd, cx_cpu, step_ids=(1,))
        out["parity"] = {"x_denoised_rmsd": par["x_denoised_rmsd"], "x_next_rmsd": par["x_next_rmsd"], "tolerance_A": 1e-3}
    smp.begin()
    for i in range(W):
        smp.step(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    smp.begin()
    e0.record()
    for i in range(SCHED_STEPS):
        x = smp.step(i)
    e1.record()
    torch.cuda.synchronize()
    assert torch.isfinite(x).all()
    ms = e0.elapsed_time(e1) / SCHED_STEPS
    out.update(ms_per_step=ms, sample_steps_per_s=Bc / (ms * 1e-3), gpu_launches_per_step=smp.launches_per_step(0))
    return out


def screening_ligand_cost(dit, dims, dev, n_ligands=4, samples=8):
    """BASELINE.json configs[3] (screening: 8 poses per ligand) as far as this path goes: wall time per ligand of the sampling
    call alone -- a NEW complex every call (pair-bias prepass, conditioning, 40 steps, poses to the host), trunk excluded
    (out of scope; measured against it in DESIGN.md section 7)."""
    from physdock_b200.sampler import PhysDockB200
    from physdock_b200.synthetic import make_complex
    model = PhysDockB200(dit, diffusion_conditioning=lambda b: (b["a"], b["ap"], b["s"], b["z"]))
    ligands = [{k: v.to(dev) for k, v in make_complex(NT, NA, dims, seed=100 + i).items()} for i in range(n_ligands + 1)]
    times = []
    for i, cx in enumerate(ligands):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        x = model.sample_diffusion(cx, num_sample=samples, steps=SCHED_STEPS, karras_noise_schedule_power=RHO, align_ref_pos=False).cpu()
        times.append(time.perf_counter() - t0)
        assert torch.isfinite(x).all()
    per = statistics.median(times[1:])                    # the first call warms allocations
    return {"workload": "BASELINE.json configs[3] per ligand: 8 poses x 40 steps at 256/2048 on one GPU, new complex per call, "
                        "trunk and featurisation excluded", "samples": samples, "ms_per_ligand": per * 1e3,
            "ligands_per_s_per_gpu": 1.0 / per, "sample_steps_per_s": samples * SCHED_STEPS / per}


def sharded_parity(dit, cx, world, rank):
    """world > 1: `sample_diffusion_sharded(exact=True)` must reproduce, slice for slice and bit for bit, the samples
    a single process draws (ShardedRNG over NCCL); 2 steps, 4 samples per rank."""
    from physdock_b200.sharding import sample_diffusion_sharded
    from physdock_b200.sampler import sample_diffusion
    n = 4 * world
    kw = dict(steps=SCHED_STEPS, karras_noise_schedule_power=RHO, align_ref_pos=False, max_steps=2)
    got = sample_diffusion_sharded(dit, cx, cx["a"], cx["ap"], cx["s"], cx["z"], num_sample=n, seed=77, exact=True, **kw)
    if rank != 0:
        return None
    torch.manual_seed(77)
    want = sample_diffusion(dit, cx, cx["a"], cx["ap"], cx["s"], cx["z"], num_sample=n, **kw)
    equal = bool(torch.equal(got, want))
    assert equal, f"sharded sampling differs from the single-process run: max abs {float((got - want).abs().max())}"
    return {"bit_equal": equal, "samples": n, "steps": 2, "backend": "nccl"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--samples", type=int, default=B_PER_GPU, help="samples per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle parity gate (tools/ A/B runs only)")
    ap.add_argument("--no-extras", action="store_true", help="skip the context measurements in config")
    ap.add_argument("--physics", action="store_true",
                    help="BASELINE.json configs[2]-style step: RDKit-free physics guidance on (40 synthetic conformer "
                         "templates, template selection + weighted Kabsch projection every guided step; late steps: "
                         "pair-energy descent, a stand-in for MMFF94 whose parity vs RDKit is UNPINNED)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, rank)

    if not torch.cuda.is_available():
        raise SystemExit("bench.py (impl=ours) needs a CUDA device: the sampling step has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    from physdock_b200.dit import B200DiT
    from physdock_b200.sampler import DiffusionSampler, PhysDockB200
    from physdock_b200.synthetic import DiTDims, make_dit_state, make_complex
    from physdock_b200.sharding import gather_samples, warm_communicator

    W, K, B = max(3, args.warmup), args.steps, args.samples
    dims = DiTDims.named("medium")
    sd = make_dit_state(dims, seed=0)
    cx_cpu = make_complex(NT, NA, dims, seed=1)
    dit = B200DiT.from_state_dict(sd, dims, device=dev)
    cx = {k: v.to(dev) for k, v in cx_cpu.items()}
    torch.manual_seed(123 + rank)                       # per-rank sampler seed
    if args.physics:
        from physdock_b200.synthetic import make_templates, make_ligand_field
        from physdock_b200.physics import PairEnergyField
        n_lig = int(cx["is_ligand"][cx["atom_id_to_token_id"]].sum())
        f = make_ligand_field(NA, n_lig, seed=2, missing=False)      # bonded chain on the ligand atoms (the last n_lig)
        field = PairEnergyField(cx["a_mask"], f["sigma"], f["eps"], f["partner"], f["partner_r0"], f["partner_k"],
                                rows=f["rows"])
        # early steps: template selection + Kabsch projection; late steps (t <= 6 A): 5 descent steps on the pair energy
        phys = dict(align_ref_pos=True, ref_mol_poses=make_templates(cx, 40), mmff_gamma_0_factor=6.0,
                    physics_field=field, mmff_iters=5)
    else:
        phys = dict(align_ref_pos=False)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world > 1:
            t = torch.tensor([x], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t)
        return x

    torch.cuda.synchronize()
    t0 = time.perf_counter()
    smp = DiffusionSampler(dit, cx, cx["a"], cx["ap"], cx["s"], cx["z"], num_sample=B, steps=SCHED_STEPS,
                           karras_noise_schedule_power=RHO, use_cuda_graph=True, **phys)     # steady state: graph replay
    torch.cuda.synchronize()
    prepare_ms = (time.perf_counter() - t0) * 1e3        # weight packing + pair-bias prepass + schedule conditioning

    # ---------------------------------------------------------------- parity of the benchmarked configuration
    parity = None
    if not args.no_parity and rank == 0 and not args.physics:
        parity = parity_gate(smp, sd, cx_cpu, step_ids=(1, 33))       # a stochastic step (t_hat ~ 3.7e3 A) and an ODE-tail step
    # the dominant kernel timed ALONE (roofline against the burst peak, which MEASURED_PEAKS.json also took on an idle chip):
    # before the sustained regions, which leave the chip in sw_power_cap at ~1750 MHz
    t_attn = f_attn = t_attn_iso = None
    if rank == 0:
        t_attn, f_attn, t_attn_iso = time_attention_kernel(dit, B)
    shard_par = None
    if world > 1:
        warm_communicator(dev)
        if not args.no_parity:
            shard_par = sharded_parity(dit, cx, world, rank)

    launches = [smp.launches_per_step(i % SCHED_STEPS) for i in range(K)]

    def timed_region(step_fn, before=None):
        """W warm-up steps once, then REPEATS regions of exactly K steps, each bracketed by barrier + synchronize and
        timed with CUDA events; returns the per-region times (ms, max over ranks)."""
        out = []
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for rep in range(REPEATS):
            smp.begin()
            if before is not None:
                before()
            if rep == 0:
                for i in range(W):
                    step_fn(i % SCHED_STEPS, warm=True)
                smp.begin()
                if before is not None:
                    before()
            barrier()
            e0.record()
            for i in range(K):
                step_fn(i % SCHED_STEPS, warm=False)
            e1.record()
            barrier()
            out.append(max_over_ranks(e0.elapsed_time(e1)))
        return out

    # ---------------------------------------------------------------- device-resident steps
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    ms_runs = timed_region(lambda i, warm: smp.step(i))
    ms_total = statistics.median(ms_runs)
    x_final = smp.x_next.clone()
    assert torch.isfinite(x_final).all(), "non-finite coordinates"

    # ---------------------------------------------------------------- end to end: pinned host in, host out
    host = {}

    def host_setup():
        host["x"] = smp.x_next.cpu().pin_memory()
        host["u4"], host["tr"] = torch.rand(B, 4).pin_memory(), torch.randn(B, 3).pin_memory()
        host["nz"], host["out"] = torch.randn(B, NA, 3).pin_memory(), torch.empty(B, NA, 3).pin_memory()
        torch.cuda.synchronize()
        host["up"] = None

    def host_step(i, warm):
        # every step: H2D of x and of the step's randoms, the step, D2H of x_next, host sync.  The randoms of step i+1 are
        # uploaded (copy stream) while step i computes; the result buffer of step i is the input buffer of step i+1.
        up = host["up"] if host["up"] is not None else smp.upload_randoms(i, host["u4"], host["tr"], host["nz"])
        smp.step_from_host(i, host["x"], host["u4"], host["tr"], host["nz"], host["out"], uploaded=up)
        host["up"] = smp.upload_randoms((i + 1) % SCHED_STEPS, host["u4"], host["tr"], host["nz"])
        torch.cuda.current_stream().synchronize()        # the caller consumes x_next on the host every step
        host["x"], host["out"] = host["out"], host["x"]

    ms_e2e_runs = timed_region(host_step, before=host_setup)
    ms_e2e = statistics.median(ms_e2e_runs)
    clk = clocks.stop() if rank == 0 else None

    # ---------------------------------------------------------------- the one collective: gather coordinates
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    gathered = gather_samples(x_final)
    e1.record()
    torch.cuda.synchronize()
    ms_gather = e0.elapsed_time(e1)
    assert gathered.shape[0] == B * world

    # ---------------------------------------------------------------- context: strong scaling of BASELINE configs[4]
    extras = {}
    if not args.no_extras and not args.physics:
        b5 = -(-C5_SAMPLES // world)                     # 40 samples of ONE complex spread over the GPUs
        smp5 = DiffusionSampler(dit, cx, cx["a"], cx["ap"], cx["s"], cx["z"], num_sample=b5, steps=SCHED_STEPS,
                                karras_noise_schedule_power=RHO, align_ref_pos=False, use_cuda_graph=True)
        smp5.begin()
        for i in range(W):
            smp5.step(i)
        barrier()
        e0.record()
        for i in range(SCHED_STEPS):
            smp5.step(i)
        e1.record()
        barrier()
        ms5 = max_over_ranks(e0.elapsed_time(e1))
        extras["strong_scaling_c5"] = {"workload": "BASELINE.json configs[4]: 40 samples of one 256/2048 complex over the GPUs",
                                       "samples_total": b5 * world, "samples_per_gpu": b5, "ms_per_step": ms5 / SCHED_STEPS,
                                       "sample_steps_per_s": b5 * world * SCHED_STEPS / (ms5 * 1e-3)}
        del smp5

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    if not args.no_extras and not args.physics:
        # full public-API call (redocking.py:284-299): PhysDockB200.sample_diffusion, 16 samples, 40 steps, wall clock
        model = PhysDockB200(dit, diffusion_conditioning=lambda batch: (batch["a"], batch["ap"], batch["s"], batch["z"]))
        for _ in range(2):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            x = model.sample_diffusion(cx, num_sample=B, steps=SCHED_STEPS, karras_noise_schedule_power=RHO,
                                       align_ref_pos=False)
            x_host = x.cpu()
            wall = time.perf_counter() - t0
        assert torch.isfinite(x_host).all()
        extras["sample_diffusion_call"] = {"api": "PhysDockB200.sample_diffusion(num_sample=16, steps=40) + .cpu() (one-shot: eager launches, no graph capture)",
                                           "wall_ms": wall * 1e3, "sample_steps_per_s": B * SCHED_STEPS / wall}
        # INTEGRATION.md section 1 path: model.dit = B200DiT..., the reference sampler calls dit.forward every step
        xh = torch.randn(B, NA, 3, device=dev) * 100
        th = torch.full([B], 100.0, device=dev)
        for _ in range(3):
            dit(cx, xh, th, cx["a"], cx["ap"], cx["s"], cx["z"])
        torch.cuda.synchronize()
        e0.record()
        for _ in range(20):
            dit(cx, xh, th, cx["a"], cx["ap"], cx["s"], cx["z"])
        e1.record()
        torch.cuda.synchronize()
        extras["dit_forward_call"] = {"api": "B200DiT.forward(batch, x_hat, t_hat, a, ap, s, z) (model.py:153,221 seam), CUDA-graph replay",
                                      "ms_per_call": e0.elapsed_time(e1) / 20}
        smp.begin()          # the sampler shares the denoiser's workspace/graphs: leave it in a defined state
        if world == 1:
            # the other single-complex configurations of BASELINE.json, one 40-step run each (these re-prepare the denoiser for
            # another complex: nothing below uses the benchmark complex's caches any more)
            extras["other_configs"] = {
                "C1": other_config(dit, sd, dims, dev, 64, 512, 4, W, False,
                                   "BASELINE.json configs[0]: crop 64 / atom crop 512, 4 samples, physics off (synthetic features)"),
                "C4": screening_ligand_cost(dit, dims, dev),
                "C3": other_config(dit, sd, dims, dev, 384, 3072, 64, W, True,
                                   "BASELINE.json configs[2]: crop 384 / atom crop 3072, 64 samples, RDKit-free physics guidance "
                                   "(40 templates, Kabsch projection, late steps: pair-energy descent; parity vs RDKit MMFF94 UNPINNED)"),
            }

    peaks = measured_peaks()
    traffic, traffic_src = attention_dram_traffic()
    value = world * B * K / (ms_total * 1e-3)
    e2e = world * B * K / (ms_e2e * 1e-3)
    F = flops_per_sample_step(NT, NA)
    cfg = {"workload": WORKLOAD + (" + physics guidance (40 templates, Kabsch projection; late steps: 5 pair-energy descent steps on the GPU; parity vs RDKit MMFF94 UNPINNED)" if args.physics else ""),
           "Nt": NT, "Na": NA, "samples_per_gpu": B, "schedule": "40 steps rho=1000",
           "timing": f"median of {REPEATS} regions of {K} steps (schedule indices 0..{K - 1} after begin()), CUDA events, max over ranks",
           "ms_per_step_runs": [m / K for m in ms_runs], "e2e_ms_per_step_runs": [m / K for m in ms_e2e_runs],
           "l2": "inputs larger than L2: 453 MB pair-bias cache + 203 MB weights streamed every step",
           "batch_steps_per_s": world * K / (ms_total * 1e-3), "gather_final_coords_ms": ms_gather,
           "prepare_complex_ms": prepare_ms,
           "gflop_per_sample_step": F / 1e9,
           "step_tensor_frac_of_sustained": (B * F / (ms_total / K * 1e-3)) / (peaks["tf_sustained"] * 1e12),
           "numerics": "fp32 data; tensor-core operands as split fp16 (hi+lo), 3 MMAs per product, fp32 accumulate"}
    cfg.update(extras)
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": cfg,
        "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": (2 * B * NA * 3 + B * 7) * 4,
                "d2h_bytes_per_step": B * NA * 3 * 4, "ms_per_step": ms_e2e / K},
        "gpu_launches": sum(launches),
        "clocks": clk,
        "parity": parity if parity is not None else ("unpinned (physics backend has no reference arithmetic)" if args.physics else "skipped"),
        "roofline": {"kernel": "attention_umma_kernel (atom pair-bias attention, S=2048 H=4 D=32)", "bound": "tensor",
                     "achieved": f_attn / t_attn / 1e12, "peak": peaks["tf_burst"], "unit": "TFLOP/s",
                     "frac": f_attn / t_attn / 1e12 / peaks["tf_burst"], "traffic": traffic,
                     "traffic_source": f"{traffic_src} (committed ncu --set full capture of this shape; not measured in this run)",
                     "issued_mma_tflops": 3 * f_attn / t_attn / 1e12,
                     "peak_source": peaks["source"] + ", burst bf16", "launch_ms": t_attn * 1e3,
                     "launch_ms_isolated": t_attn_iso * 1e3,
                     "timing": "kernel timed alone before the sustained regions: average launch duration over a CUDA-graph replay of 12 back-to-back launches cycling the 6 cached bias blocks (as in the step); launch_ms_isolated = event after every launch",
                     "algorithmic_gflop_per_launch": f_attn / 1e9},
    }
    if shard_par is not None:
        out["sharded_parity"] = shard_par
    if not args.no_cpu_baseline and world == 1:
        threads = os.cpu_count() or 1
        rate, sec = cpu_oracle_steps(5, 1, B, threads)       # a bounded sample of the same workload (~15-25 s of CPU work)
        out["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
                               "sample": f"5 timed steps (+1 warm-up) of B={B} samples at Nt={NT}/Na={NA}, fp32 PyTorch CPU "
                                         f"oracle (bit-identical to the reference on CPU), {sec:.2f} s/step"}
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
