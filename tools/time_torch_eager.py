"""Context number (not a bench leg): the reference's arithmetic (oracle port = the reference's PyTorch modules, fp32) run
EAGERLY ON THE B200 at the benchmark shape, TF32 off and on -- "the bar to beat on the same box" of SURVEY.md section 8d."""
import sys, os, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import physdock_oracle as O
from physdock_b200.synthetic import DiTDims, make_dit_state, make_complex
dev = torch.device("cuda")
dims = DiTDims.named("medium")
sd = {k: v.to(dev) for k, v in make_dit_state(dims, seed=0).items()}
cx = {k: v.to(dev) for k, v in make_complex(256, 2048, dims, seed=1).items()}
B = int(os.environ.get("B", 16))
x_hat = torch.randn(B, 2048, 3, device=dev) * 30
t_hat = torch.full([B], 25.0, device=dev)
for tf32 in (False, True):
    torch.backends.cuda.matmul.allow_tf32 = tf32
    torch.backends.cudnn.allow_tf32 = tf32
    with torch.no_grad():
        for _ in range(2):
            O.af3dit_forward(sd, cx, x_hat, t_hat, cx["a"], cx["ap"], cx["s"], cx["z"])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        n = 5
        for _ in range(n):
            O.af3dit_forward(sd, cx, x_hat, t_hat, cx["a"], cx["ap"], cx["s"], cx["z"])
        e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(f"PyTorch eager denoiser on the B200, B={B}, TF32 {'on ' if tf32 else 'off'}: {ms:8.2f} ms per call = {B / ms * 1e3:8.1f} sample-steps/s")
