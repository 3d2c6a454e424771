"""Times the x3 tcgen05 GEMM / wgrad kernels against cuBLAS fp32 and TF32 at update-chunk size."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "2048-ppo_b200")]
from g2048 import linear

M = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 22
dev = torch.device("cuda")
x = torch.randn(M, 196, device=dev)
dy = torch.randn(M, 196, device=dev) * 1e-2
w = torch.randn(196, 196, device=dev) * 0.1
x48 = torch.randn(M, 48, device=dev)
w48 = torch.randn(196, 48, device=dev)

def t(f, n=5):
    f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

img, imgT, img48 = linear.pack_weight(w), linear.pack_weight(w, True), linear.pack_weight(w48)
rows = []
rows.append(("x3 gemm 196x196", t(lambda: linear.gemm(x, img, 196)), 2 * 196 * 4))
rows.append(("x3 gemm K=48", t(lambda: linear.gemm(x48, img48, 196)), (48 + 196) * 4))
rows.append(("x3 wgrad 196x196", t(lambda: linear.wgrad(dy, x)), 2 * 196 * 4))
rows.append(("x3 wgrad K=48", t(lambda: linear.wgrad(dy, x48)), (48 + 196) * 4))
for tf in (False, True):
    torch.backends.cuda.matmul.allow_tf32 = tf
    tag = "tf32" if tf else "fp32"
    rows.append((f"cublas {tag} fwd", t(lambda: x @ w.T), 2 * 196 * 4))
    rows.append((f"cublas {tag} wgrad", t(lambda: dy.T @ x), 2 * 196 * 4))
for name, ms, bytes_per in rows:
    print(f"{name:22s} {ms:8.3f} ms  {M * bytes_per / ms / 1e6:8.1f} GB/s  {2 * M * 196 * 196 / ms / 1e9:7.1f} TFLOP/s-equiv")
