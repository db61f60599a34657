"""Where the out-of-scope trunk spends its GPU time (context for SURVEY.md section 8 row f2): torch.profiler over one call of the
reference's `DiffusionConditioning` (from baseline/_ref, random weights, synthetic FeatureLoader-shaped inputs, TF32 as the
reference sets it) at Nt=256 / Na=2048 on the B200.   python tools/prof_trunk.py > profiles/r02_trunk_profile.txt"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
os.environ.setdefault("PHYSDOCK_REFERENCE", os.path.join(ROOT, "baseline", "_ref"))
from oracle.ref_import import import_reference
from time_trunk import synthetic_features
from torch.profiler import profile, ProfilerActivity, record_function

Nt, Na = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (256, 2048)
dev = torch.device("cuda")
PhysDock, PhysDockConfig, _, _ = import_reference()
torch.set_float32_matmul_precision("high")
torch.manual_seed(0)
model = PhysDock(PhysDockConfig(model_name="medium")).float().eval().to(dev)
dc = model.diffusion_conditioning
batch = synthetic_features(Nt, Na, device=dev)
# wrap the trunk's direct sub-modules so the table also shows where the time goes structurally
for name, child in dc.named_children():
    orig = child.forward
    def wrapped(*a, _orig=orig, _name=name, **k):
        with record_function("trunk." + _name):
            return _orig(*a, **k)
    child.forward = wrapped
with torch.inference_mode():
    for _ in range(2):
        dc(batch)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        dc(batch)
        torch.cuda.synchronize()
ev = prof.key_averages()
ops = [e for e in ev if e.key.startswith("aten::") and e.self_device_time_total > 0]      # kernels are attributed to their aten op
tot = sum(e.self_device_time_total for e in ops)
print(f"# reference trunk DiffusionConditioning, Nt={Nt} Na={Na}, TF32 on: {tot / 1e3:.1f} ms of GPU kernel time in one call, "
      f"{sum(e.count for e in ops)} operator calls that launch kernels")
print("# by sub-module (device time incl. children)")
mods = {}
for e in ev:
    if e.key.startswith("trunk."):
        mods[e.key] = max(mods.get(e.key, 0.0), e.device_time_total)
for k, v in sorted(mods.items(), key=lambda kv: -kv[1]):
    print(f"  {k:40s} {v / 1e3:9.2f} ms")
print("# by operator (self device time)")
for e in sorted(ops, key=lambda e: -e.self_device_time_total)[:18]:
    print(f"  {e.key[:40]:40s} {e.self_device_time_total / 1e3:9.2f} ms  {100 * e.self_device_time_total / tot:5.1f} %  calls={e.count}")
gemm = sum(e.self_device_time_total for e in ops if e.key in ("aten::mm", "aten::bmm", "aten::addmm", "aten::matmul", "aten::linear"))
print(f"# GEMM operators (mm / bmm / addmm): {gemm / 1e3:.1f} ms = {100 * gemm / tot:.0f} %; everything else is elementwise / reduction / softmax / copy traffic over the pair tensors")
