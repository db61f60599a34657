"""Turns the scratch captures of tools/gpu_profile.sh (gpurun_out/) into the tracked summaries under profiles/ (round tag = argv[1]).
   python tools/summarise_profiles.py r02"""
import collections, csv, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
SPECS = [("attn_atom", 0, "attn_atom", "attention_umma_kernel, atom shape B=16 H=4 S=2048 (tools/prof_attention.py)"),
         ("attn_token", 0, "attn_token", "attention_umma_kernel, token shape B=16 H=16 S=256"),
         ("transition", 0, "transition", "transition_umma_kernel, atom DiTTransition M=32768 (tools/prof_transition.py)"),
         ("gemm", 0, "gemm_token_swiglu", "gemm_umma_kernel EPI_SWIGLU token shape 4096x2816x512 (pair tiling <2,1,0>; tools/prof_gemm.py)"),
         ("gemm", 1, "gemm_atom_swiglu", "gemm_umma_kernel EPI_SWIGLU atom shape 32768x768x128 (single-CTA tiling <2,0,0>)"),
         ("qkv", 0, "gemm_atom_qkv", "gemm_umma_kernel EPI_QKV atom shape 32768x384x128 (tools/prof_qkv.py)"),
         ("qkv", 1, "gemm_token_qkv", "gemm_umma_kernel EPI_QKV token shape 4096x1536x512 (pair tiling)"),
         ("physics", 0, "physics", "pair_energy_grad_kernel (tools/prof_physics.py)")]
for rep, idx, name, desc in SPECS:
    src = os.path.join(G, rep + ".ncu-rep")
    if not os.path.exists(src):
        continue
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), src, str(idx), "12"], capture_output=True, text=True).stdout
    with open(os.path.join(P, f"{tag}_{name}_ncu.txt"), "w") as f:
        f.write(f"# {tag}: ncu --set full --clock-control none --import-source on, kernel: {desc}\n# summarised by tools/ncu_summary.py from gpurun_out/{rep}.ncu-rep (kernel index {idx})\n" + out)
# coordinate / physics-guidance kernels: one line per captured launch
src = os.path.join(G, "coords.ncu-rep")
if os.path.exists(src):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    h = rows[0]
    kn, t, g, dr, dw = (h.index(c) for c in ("Kernel Name", "gpu__time_duration.sum", "Grid Size", "dram__bytes_read.sum", "dram__bytes_write.sum"))
    with open(os.path.join(P, f"{tag}_coords_ncu.txt"), "w") as f:
        f.write(f"# {tag}: ncu --set full --clock-control none of the coordinate / physics-guidance kernels inside `bench.py --physics` (B=16, Na=2048)\n")
        f.write("# kernel, grid, gpu__time_duration (us), dram read, dram write  [latency-bound: < 1.5 MB of traffic each]\n")
        for r in rows[2:]:
            f.write(f"{re.sub(r'.*::', '', r[kn].split('(')[0]):28s} {r[g]:14s} {r[t]:>10s} us   {r[dr]:>10s} {rows[1][dr]}  {r[dw]:>10s} {rows[1][dw]}\n")
# launch list of one device-resident step
src = os.path.join(G, "launches.csv")
if os.path.exists(src):
    rows = [r for r in csv.reader(l for l in open(src) if not l.startswith("=="))]
    h = rows[0]
    kn, mv, gs, bs, idc, mu = (h.index(c) for c in ("Kernel Name", "Metric Value", "Grid Size", "Block Size", "ID", "Metric Unit"))
    L = []
    for r in rows[1:]:
        if len(r) <= mv:
            continue
        v = float(r[mv].replace(",", "")) / (1e3 if r[mu] == "ns" else 1.0)
        raw = r[kn]
        name = raw.replace("void ", "").replace("pdk::<unnamed>::", "").split("(")[0]
        if raw.startswith("void at::") or "at_cuda_detail" in raw or raw.startswith("std::enable_if"):
            name = "torch:" + name.replace("at::", "").replace("<unnamed>::", "")[:44]
        L.append((int(r[idc]), name, r[gs], r[bs], v))
    idx = [i for i, x in enumerate(L) if x[1].startswith("centre_augment")]
    def start(k):
        s = idx[k]
        while s > 0 and L[s - 1][1].startswith("torch:"):
            s -= 1
        return s
    best = None
    for k in range(len(idx) - 1):
        st = L[start(k):start(k + 1)]
        if sum("distribution" in x[1] for x in st) >= 6 and sum(x[1].startswith("torch:") for x in st) <= 9:
            best = st
    if best:
        with open(os.path.join(P, f"{tag}_launches_step.csv"), "w") as f:
            f.write(f"# {tag}: ncu launch list of `python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity --no-extras`\n"
                    "# ncu --metrics gpu__time_duration.sum --clock-control none; ONE device-resident sampler step (B=16, Nt=256, Na=2048); the kernels of the CUDA-graph replay are listed individually\n"
                    f"# per-launch durations are cold-cache/serialised: compare SHARES (in-graph costs: {tag}_step_attribution.txt)\n"
                    "id,kernel,grid,block,duration_us\n")
            for x in best:
                f.write(f'{x[0]},"{x[1]}","{x[2]}","{x[3]}",{x[4]:.2f}\n')
        tot = sum(x[4] for x in best)
        att = sum(x[4] for x in best if x[1].startswith("attention") and x[2].startswith("(29"))
        print(f"launch list: {len(best)} kernels ({sum(not x[1].startswith('torch:') for x in best)} ours), {tot:.0f} us serialised, atom attention {100 * att / tot:.1f} %")
for a, b, hdr in (("step_attribution.txt", f"{tag}_step_attribution.txt", None), ("attention_timeline.txt", f"{tag}_attention_timeline.txt", None)):
    if os.path.exists(os.path.join(G, a)):
        open(os.path.join(P, b), "w").write(open(os.path.join(G, a)).read())
if os.path.exists(os.path.join(G, "time_gemm.log")):
    with open(os.path.join(P, f"{tag}_kernel_times_graph.txt"), "w") as f:
        f.write("# CUDA-graph-timed launches (tools/time_gemm.py, tools/time_attention.py): us per launch, back to back, B=16 shapes\n")
        f.write("".join(l for l in open(os.path.join(G, "time_gemm.log")) if "product build" not in l))
print("done")
