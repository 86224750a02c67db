"""Where does the host-buffer step spend its time?  dn_step (the plain launch) is timed with CUDA events and with a
wall clock + stream synchronise for every placement of the kernel's inputs and outputs: device memory or pinned,
device-mapped host memory (what the zero-copy form of dn_step_host uses).  Run under gpurun."""
import ctypes as C
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from drl_dronenavigation_b200 import _lib as L


class A:
    substeps, track, actions = 8, "circle", "saturating"


dev = torch.device("cuda", 0)
for N in (12, 4096, 16384):
    env = bench.make_env(N, A, dev)
    env.reset()
    D = env.obs_dim

    def bufs(host):
        mk = (lambda *s, dtype: torch.empty(*s, dtype=dtype).pin_memory()) if host else (lambda *s, dtype: torch.empty(*s, dtype=dtype, device=dev))
        return mk
    out = {}
    for a_host in (False, True):
        for o_host in ("none", "small", "obs", "all"):     # which outputs live in host memory
            act = bufs(a_host)(8, N, 4, dtype=torch.float32)
            act.copy_(torch.rand(8, N, 4) * 2 - 1)
            obs = bufs(o_host in ("obs", "all"))(N, D, dtype=torch.float32)
            sm = o_host in ("small", "all")
            rew, done, found = bufs(sm)(N, dtype=torch.float32), bufs(sm)(N, dtype=torch.uint8), bufs(sm)(N, dtype=torch.int32)
            ios = [env._make_io(act[k], obs, rew, done, None, found) for k in range(8)]
            st = env._stream()
            for k in range(30):
                L.check(env._lib.dn_step(env._handle, C.byref(ios[k % 8]), st))
            torch.cuda.synchronize()
            K = 300
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for k in range(K):
                L.check(env._lib.dn_step(env._handle, C.byref(ios[k % 8]), st))
            e1.record()
            torch.cuda.synchronize()
            back_to_back = 1e3 * e0.elapsed_time(e1) / K
            t0 = time.perf_counter()
            for k in range(K):
                L.check(env._lib.dn_step(env._handle, C.byref(ios[k % 8]), st))
                torch.cuda.synchronize()
            sync_each = 1e6 * (time.perf_counter() - t0) / K
            out[("host" if a_host else "dev", o_host)] = (round(back_to_back, 2), round(sync_each, 2))
    print(f"N={N}: (actions, host outputs) -> (us per launch back to back, us per launch + stream sync)")
    for k, v in out.items():
        print("   ", k, v)
    # the product call for comparison
    mk = bufs(True)
    h_act = mk(8, N, 4, dtype=torch.float32); h_act.uniform_(-1, 1)
    h_obs, h_rew, h_done, h_found = mk(N, D, dtype=torch.float32), mk(N, dtype=torch.float32), mk(N, dtype=torch.uint8), mk(N, dtype=torch.int32)
    ios = [env._make_io(h_act[k], h_obs, h_rew, h_done, None, h_found) for k in range(8)]
    for k in range(50):
        env.step_host(ios[k % 8])
    t0 = time.perf_counter()
    for k in range(500):
        env.step_host(ios[k % 8])
    print("    dn_step_host (zero copy):", round(1e6 * (time.perf_counter() - t0) / 500, 2), "us")
    env.close()
