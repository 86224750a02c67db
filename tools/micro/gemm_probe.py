"""Which formulation of the skinny first-layer GEMMs of the PPO MLP (K = 13) does cuBLAS run fastest?  Run under gpurun."""
import torch

torch.backends.cuda.matmul.allow_tf32 = True
B = 32768
dev = "cuda"
x13, x16 = torch.randn(B, 13, device=dev), torch.randn(B, 16, device=dev)
w13, w16 = torch.randn(512, 13, device=dev), torch.randn(512, 16, device=dev)
dy = torch.randn(B, 512, device=dev)


def t(fn, n=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


print("fwd  x13 @ w13.T            ", round(t(lambda: x13 @ w13.t()), 1), "us")
print("fwd  x16 @ w16.T            ", round(t(lambda: x16 @ w16.t()), 1), "us")
print("dW   dy.T @ x13  [512,13]   ", round(t(lambda: dy.t() @ x13), 1), "us")
print("dW   (x13.T @ dy).T         ", round(t(lambda: (x13.t() @ dy).t()), 1), "us")
print("dW   dy.T @ x16  [512,16]   ", round(t(lambda: dy.t() @ x16), 1), "us")
print("dW   (x16.T @ dy).T         ", round(t(lambda: (x16.t() @ dy).t()), 1), "us")
print("db   dy.sum(0)              ", round(t(lambda: dy.sum(0)), 1), "us")
ones = torch.ones(1, B, device=dev)
print("db   ones @ dy              ", round(t(lambda: ones @ dy), 1), "us")
onesv = torch.ones(B, device=dev)
print("db   mv(dy.T, ones)         ", round(t(lambda: torch.mv(dy.t(), onesv)), 1), "us")
h = torch.randn(B, 512, device=dev)
w = torch.randn(512, 512, device=dev)
print("fwd  512x512                ", round(t(lambda: h @ w.t()), 1), "us")
print("dW   dy.T @ h [512,512]     ", round(t(lambda: dy.t() @ h), 1), "us")
print("dX   dy @ w                 ", round(t(lambda: dy @ w), 1), "us")
print("tanh                        ", round(t(lambda: torch.tanh(h)), 1), "us")
hb, wb, dyb = h.bfloat16(), w.bfloat16(), dy.bfloat16()
print("bf16 fwd 512x512            ", round(t(lambda: hb @ wb.t()), 1), "us")
print("bf16 dW 512x512             ", round(t(lambda: dyb.t() @ hb), 1), "us")
