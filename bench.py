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

Prints ONE JSON line (rank 0).  value = sample-steps/s with inputs resident in HBM; e2e = the same step
with per-step inputs copied from pinned host memory and the result read back.
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
METRIC = "denoising sample-steps/sec at crop=256, atom_crop=2048 (B=16 samples per GPU)"
UNIT = "sample-steps/s"
WORKLOAD = ("redocking Posebusters_subset, crop_size=256 atom_crop_size=2048, 16 samples, 1xB200 "
            "(BASELINE.json configs[1]); synthetic Nt=256 Na=2048, round 0: align_ref_pos=False, ref_mol=None")


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


def cpu_oracle_rate(n_calls, B_cpu, threads):
    """The reference's own CPU path for one step (oracle port, bit-identical to the reference on CPU:
    tests/test_oracle_pin.py), timed on the host cores.  Returns (sample-steps/s, seconds per step)."""
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
        for i in range(n_calls + 1):                   # first call is the warm-up
            t0 = time.perf_counter()
            t_cur, t_next = sig[i], sig[i + 1]
            u = torch.stack([rng.rand((B_cpu,)) for _ in range(4)], -1)
            x_cur = O.centre_random_augmentation(x, cx["a_mask"], u, rng.normal((B_cpu, 3)))
            t_hat = torch.full([B_cpu], float(t_cur * 1.8))
            x_hat = O.diffuse(x_cur, t_hat, t_cur, rng.normal(x_cur.shape), 1.003)
            x_den = O.af3dit_forward(sd, cx, x_hat, t_hat, cx["a"], cx["ap"], cx["s"], cx["z"])
            x = O.euler_update(x_hat, (x_hat - x_den) / t_hat[:, None, None], t_hat, t_next, 1.5)
            times.append(time.perf_counter() - t0)
    sec = statistics.mean(times[1:])
    return B_cpu / sec, sec


def run_reference(args, rank):
    """--impl reference: the reference's CPU implementation of the step (oracle port; /root/reference does not
    exist on the GPU box), all host threads, bounded sample B=8 per step."""
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    B_cpu = 8
    n = max(1, min(args.steps, 12))
    rate, sec = cpu_oracle_rate(n, B_cpu, threads)
    sample = f"{n} timed steps (+1 warm-up) of B={B_cpu} samples at Nt={NT}/Na={NA}, fp32 PyTorch CPU, {threads} threads"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": n,
        "warmup": 1, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": {"workload": WORKLOAD, "Nt": NT, "Na": NA, "B": B_cpu},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def attention_dram_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the attention kernel, from the committed
    `ncu --set full` capture summarised in profiles/r01_attn_atom_ncu.txt (same shape as the bench)."""
    p = os.path.join(ROOT, "profiles", "r01_attn_atom_ncu.txt")
    try:
        tot = 0.0
        for line in open(p):
            f = line.split()
            if f and f[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                tot += float(f[1]) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}[f[2]]
        return tot or None
    except (OSError, ValueError, KeyError, IndexError):
        return None


def time_attention_kernel(dit, B, iters=12):
    """Average duration of the dominant kernel (atom pair-bias attention) measured with CUDA events on the
    launching stream; the 6 cached bias blocks (67 MB each) are cycled so no launch finds its bias in L2."""
    from physdock_b200 import ops
    dev = torch.device("cuda")
    H, S = 4, dit._complex_keep["Na"]
    S_pad = ops.pad_len(S)
    g = torch.Generator(device=dev).manual_seed(0)
    planes = [torch.randn(B, H, S_pad, 64, generator=g, device=dev).half() for _ in range(3)]
    bias = dit._complex_keep["bias_a"].view(-1, H, S_pad, S_pad)
    for l in range(3):
        ops.attention(*planes, bias[l % bias.shape[0]])
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
    ev[0].record()
    for i in range(iters):
        ops.attention(*planes, bias[i % bias.shape[0]])
        ev[i + 1].record()
    torch.cuda.synchronize()
    ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(iters)]
    flops = B * H * 4 * S * S * 32          # QK^T + PV, 2 flop per MAC (algorithmic, not the 3x split)
    return statistics.mean(ms) * 1e-3, flops


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--samples", type=int, default=B_PER_GPU, help="samples per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--physics", action="store_true",
                    help="BASELINE.json configs[2]-style step: RDKit-free physics guidance on (40 synthetic conformer "
                         "templates, template selection + weighted Kabsch projection every guided step)")
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
    from physdock_b200.sampler import DiffusionSampler
    from physdock_b200.synthetic import DiTDims, make_dit_state, make_complex
    from physdock_b200.sharding import gather_samples

    W, K, B = max(3, args.warmup), args.steps, args.samples
    dims = DiTDims.named("medium")
    dit = B200DiT.from_state_dict(make_dit_state(dims, seed=0), dims, device=dev)
    cx = {k: v.to(dev) for k, v in make_complex(NT, NA, dims, seed=1).items()}
    torch.manual_seed(123 + rank)                       # per-rank sampler seed
    phys = {}
    if args.physics:
        from physdock_b200.synthetic import make_templates
        from physdock_b200.physics import PairEnergyField
        from physdock_b200.synthetic import make_ligand_field
        n_lig = int(cx["is_ligand"][cx["atom_id_to_token_id"]].sum())
        f = make_ligand_field(NA, n_lig, seed=2, missing=False)      # bonded chain on the ligand atoms (the last n_lig)
        field = PairEnergyField(cx["a_mask"], f["sigma"], f["eps"], f["partner"], f["partner_r0"], f["partner_k"],
                                rows=f["rows"])
        # early steps: template selection + Kabsch projection; late steps (t <= 6 A): 5 descent steps on the pair energy
        phys = dict(align_ref_pos=True, ref_mol_poses=make_templates(cx, 40), mmff_gamma_0_factor=6.0,
                    physics_field=field, mmff_iters=5)
    else:
        phys = dict(align_ref_pos=False)
    smp = DiffusionSampler(dit, cx, cx["a"], cx["ap"], cx["s"], cx["z"], num_sample=B, steps=SCHED_STEPS,
                           karras_noise_schedule_power=RHO, **phys)
    smp.begin()
    launches_per_step = dit.launches_per_denoise() + 2 + (3 if args.physics else 0)   # + centre_augment + euler (+ physics)

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

    # ---------------------------------------------------------------- device-resident steps
    for i in range(W):
        smp.step(i % SCHED_STEPS)
    barrier()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(K):
        smp.step((W + i) % SCHED_STEPS)
    e1.record()
    barrier()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    x_final = smp.x_next
    assert torch.isfinite(x_final).all(), "non-finite coordinates"

    # ---------------------------------------------------------------- end to end: pinned host in, host out
    x_h = x_final.cpu().pin_memory()
    u4_h = torch.rand(B, 4).pin_memory()
    tr_h = torch.randn(B, 3).pin_memory()
    nz_h = torch.randn(B, NA, 3).pin_memory()
    out_h = torch.empty(B, NA, 3).pin_memory()
    for i in range(W):
        smp.step_from_host(i % SCHED_STEPS, x_h, u4_h, tr_h, nz_h, out_h)
    barrier()
    barrier()
    e0.record()
    # every step: H2D of x and of the step's randoms, the step, D2H of x_next, host sync.  The randoms of step i+1 are
    # uploaded (copy stream) while step i computes; the result buffer of step i is the input buffer of step i+1.
    up = smp.upload_randoms(W % SCHED_STEPS, u4_h, tr_h, nz_h)
    for i in range(K):
        smp.step_from_host((W + i) % SCHED_STEPS, x_h, u4_h, tr_h, nz_h, out_h, uploaded=up)
        if i + 1 < K:
            up = smp.upload_randoms((W + i + 1) % SCHED_STEPS, u4_h, tr_h, nz_h)
        torch.cuda.current_stream().synchronize()        # the caller consumes x_next on the host every step
        x_h, out_h = out_h, x_h
    e1.record()
    barrier()
    ms_e2e = max_over_ranks(e0.elapsed_time(e1))
    clk = clocks.stop() if rank == 0 else None

    # ---------------------------------------------------------------- the one collective: gather coordinates
    barrier()
    e0.record()
    gathered = gather_samples(x_final)
    e1.record()
    torch.cuda.synchronize()
    ms_gather = e0.elapsed_time(e1)
    assert gathered.shape[0] == B * world

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks = measured_peaks()
    t_attn, f_attn = time_attention_kernel(dit, B)
    value = world * B * K / (ms_total * 1e-3)
    e2e = world * B * K / (ms_e2e * 1e-3)
    F = flops_per_sample_step(NT, NA)
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD + (" + physics guidance (40 templates, Kabsch projection; late steps: 5 pair-energy descent steps on the GPU)" if args.physics else ""),
                   "Nt": NT, "Na": NA, "samples_per_gpu": B, "schedule": "40 steps rho=1000",
                   "l2": "inputs larger than L2: 453 MB pair-bias cache + 203 MB weights streamed every step",
                   "batch_steps_per_s": world * K / (ms_total * 1e-3), "gather_final_coords_ms": ms_gather,
                   "gflop_per_sample_step": F / 1e9,
                   "step_tensor_frac_of_sustained": (B * F / (ms_total / K * 1e-3)) / (peaks["tf_sustained"] * 1e12),
                   "numerics": "fp32 data; tensor-core operands as split fp16 (hi+lo), 3 MMAs per product, fp32 accumulate"},
        "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": (2 * B * NA * 3 + B * 7) * 4,
                "d2h_bytes_per_step": B * NA * 3 * 4, "ms_per_step": ms_e2e / K},
        "gpu_launches": K * launches_per_step,
        "clocks": clk,
        "roofline": {"kernel": "attention_umma_kernel (atom pair-bias attention, S=2048 H=4 D=32)", "bound": "tensor",
                     "achieved": f_attn / t_attn / 1e12, "peak": peaks["tf_burst"], "unit": "TFLOP/s",
                     "frac": f_attn / t_attn / 1e12 / peaks["tf_burst"], "traffic": attention_dram_traffic(),
                     "issued_mma_tflops": 3 * f_attn / t_attn / 1e12,
                     "peak_source": peaks["source"] + ", burst bf16", "launch_ms": t_attn * 1e3,
                     "algorithmic_gflop_per_launch": f_attn / 1e9},
    }
    if not args.no_cpu_baseline and world == 1:
        threads = os.cpu_count() or 1
        rate, sec = cpu_oracle_rate(10, 8, threads)       # ~10 s of CPU work: a bounded sample of the same workload
        out["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
                               "sample": f"10 timed steps (+1 warm-up) of B=8 samples at Nt={NT}/Na={NA}, fp32 PyTorch CPU "
                                         f"oracle (bit-identical to the reference on CPU), {sec:.2f} s/step"}
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
