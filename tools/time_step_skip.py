"""In-graph attribution of the step: times the device-resident sampler step with groups of launches dropped
(debug library built with -DPDK_DBG_SKIP, env PDK_SKIP=<labels>; results are wrong by construction, only the time counts).
The difference to the full step is what that group costs INSIDE the CUDA graph with PDL overlap -- unlike the ncu launch
list, which serialises kernels and runs them cold."""
import sys, os, subprocess, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GROUPS = [("full step", ""), ("attention (atom + token)", "attention"), ("QKV GEMMs", "gemm(qkv)"), ("out-proj GEMMs", "gemm(out)"),
          ("token SwiGLU GEMM", "gemm(w13)"), ("token w2 GEMM", "gemm(w2)"), ("atom fused transition", "transition"),
          ("AdaLN kernels", "adaln"),
          ("glue (precond, down/upscale, pooling, output + fused Euler)", "precond,segment_mean,gather_add,denoise_out,split,gemm(down),gemm(up)"),
          ("coordinate kernel (centre + augment + noise)", "centre_augment")]
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch
    sys.path.insert(0, ROOT)
    from physdock_b200.dit import B200DiT
    from physdock_b200.sampler import DiffusionSampler
    from physdock_b200.synthetic import DiTDims, make_dit_state, make_complex
    dev = torch.device("cuda")
    dims = DiTDims.named("medium")
    dit = B200DiT.from_state_dict(make_dit_state(dims, seed=0), dims, device=dev)
    cx = {k: v.to(dev) for k, v in make_complex(256, 2048, dims, seed=1).items()}
    torch.manual_seed(0)
    smp = DiffusionSampler(dit, cx, cx["a"], cx["ap"], cx["s"], cx["z"], num_sample=16, steps=40, karras_noise_schedule_power=1000, align_ref_pos=False, use_cuda_graph=True)
    smp.begin()
    for i in range(4): smp.step(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(20): smp.step(4 + i)
    e1.record(); torch.cuda.synchronize()
    print(json.dumps({"ms": e0.elapsed_time(e1) / 20}))
    sys.exit(0)
lib = os.path.join(ROOT, "build", "dbg", "libpdk_SKIP.so")
base = None
for name, skip in GROUPS:
    env = dict(os.environ, PHYSDOCK_B200_LIB=lib, PDK_SKIP=skip)
    if not skip: env.pop("PDK_SKIP")
    out = subprocess.run([sys.executable, __file__, "child"], env=env, capture_output=True, text=True)
    try:
        ms = json.loads(out.stdout.strip().splitlines()[-1])["ms"]
    except Exception:
        print(name, "FAILED", out.stderr[-300:]); continue
    if base is None: base = ms
    print(f"{name:55s} step {ms:6.3f} ms   group costs {1e3 * (base - ms):7.1f} us = {100 * (base - ms) / base:5.1f} %")
