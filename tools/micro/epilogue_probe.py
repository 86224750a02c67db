"""Attribution of the forward / dgrad contraction kernels' time: the same launch with parts switched off (DN_MLP_DBG bits:
1 no tanh, 2 no staging / TMA stores, 4 no MMAs, 8 no TMEM prefetch).  One process per variant (the flag is read at plan time)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
CHILD = r'''
import sys, os, ctypes as C, torch
sys.path.insert(0, %r)
from drl_dronenavigation_b200 import _lib as L
lib = L.lib()
kind, M, N, K, passes = [int(x) for x in sys.argv[1:6]]
g = torch.Generator(device="cuda").manual_seed(0)
def planes(r, c): return (torch.randn(2, r, c, device="cuda", generator=g) * 0.1).to(torch.bfloat16)
a = planes(M, K); b = planes(N, K) if kind == 0 else planes(K, N); h = planes(M, N); out = torch.zeros(2, M, N, dtype=torch.bfloat16, device="cuda")
bias = torch.zeros(N, device="cuda")
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
def run():
    L.check(lib.dn_mlp_gemm(kind, passes, M, N, K, 1, a.data_ptr(), b.data_ptr(), bias.data_ptr(), 1, h.data_ptr(), out.data_ptr(), None, st))
for _ in range(5): run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50): run()
e1.record(); torch.cuda.synchronize()
print(f"{e0.elapsed_time(e1) * 20:.1f}")
''' % ROOT
def t(kind, M, N, K, passes, dbg):
    env = dict(os.environ, DN_MLP_DBG=str(dbg))
    return float(subprocess.run([sys.executable, "-c", CHILD, str(kind), str(M), str(N), str(K), str(passes)], env=env, capture_output=True, text=True, check=True).stdout.strip())
for name, kind, M, N, K in (("fwd L1 (K=64)", 0, 32768, 512, 64), ("fwd L2 (K=512)", 0, 32768, 512, 512), ("dgrad L2 (K=512)", 1, 32768, 512, 512)):
    for passes in (3, 1):
        row = {dbg: t(kind, M, N, K, passes, dbg) for dbg in (0, 1, 2, 3, 4, 6, 7, 8)}
        print(f"{name:18s} passes={passes}  us per launch by DN_MLP_DBG: " + "  ".join(f"{k}:{v:6.1f}" for k, v in row.items()), flush=True)
