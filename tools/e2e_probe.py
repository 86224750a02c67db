"""Where do the ~27 us of a 4096-env dn_step_host call go?  Times the host-buffer call at several batch sizes (fixed
cost vs transfer cost) in both modes (zero-copy pinned buffers / staged pageable buffers).  Run under gpurun."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench


class A:
    substeps, track, actions = 8, "circle", "saturating"


dev = torch.device("cuda", 0)
for N in (128, 1024, 4096, 16384, 65536):
    env = bench.make_env(N, A, dev)
    env.reset()
    D = env.obs_dim
    res = {}
    for mode in ("pinned", "pageable"):
        mk = (lambda *s, dtype: torch.empty(*s, dtype=dtype).pin_memory()) if mode == "pinned" else (lambda *s, dtype: torch.empty(*s, dtype=dtype))
        h_act = mk(8, N, 4, dtype=torch.float32); h_act.uniform_(-1, 1)
        h_obs, h_rew = mk(N, D, dtype=torch.float32), mk(N, dtype=torch.float32)
        h_done, h_found = mk(N, dtype=torch.uint8), mk(N, dtype=torch.int32)
        ios = [env._make_io(h_act[k], h_obs, h_rew, h_done, None, h_found) for k in range(8)]
        for k in range(50):
            env.step_host(ios[k % 8])
        K = 500
        t0 = time.perf_counter()
        for k in range(K):
            env.step_host(ios[k % 8])
        res[mode] = 1e6 * (time.perf_counter() - t0) / K
    # device-only step + sync for comparison (same Python path, no host traffic)
    acts = bench.make_actions(8, N, "saturating", dev, seed=1)
    for k in range(50):
        env.step(acts[k % 8])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k in range(500):
        env.step(acts[k % 8]); torch.cuda.synchronize()
    res["device_step_plus_sync"] = 1e6 * (time.perf_counter() - t0) / 500
    print(N, {k: round(v, 2) for k, v in res.items()}, "bytes", N * 16, N * (D * 4 + 9))
    env.close()
