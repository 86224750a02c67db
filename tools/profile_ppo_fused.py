"""Per-kernel time of the fused PPO minibatch step at the bench shape (torch.profiler / CUPTI sees the library's kernels)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from drl_dronenavigation_b200.ppo import PPOConfig, PPOLearner

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
mb = int(sys.argv[2]) if len(sys.argv) > 2 else 32768
prec = sys.argv[3] if len(sys.argv) > 3 else "bf16x3"
L = PPOLearner(13, 4, PPOConfig(batch_size=mb, n_epochs=1, target_kl=None, mlp_precision=prec), device="cuda")
g = torch.Generator(device="cuda").manual_seed(0)
obs = torch.randn(B, 13, device="cuda", generator=g)
act = torch.rand(B, 4, device="cuda", generator=g) * 2 - 1
with torch.no_grad():
    a, logp, v = L.act(obs[:65536])
logp = torch.randn(B, device="cuda", generator=g) * 0.1 - 5.5
v = torch.randn(B, device="cuda", generator=g)
adv, ret = torch.randn(B, device="cuda", generator=g), torch.randn(B, device="cuda", generator=g)
gen = torch.Generator(device="cuda").manual_seed(1)
L.update(obs, act, logp, v, adv, ret, generator=gen)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
out = L.update(obs, act, logp, v, adv, ret, generator=gen)
e1.record()
torch.cuda.synchronize()
n = out["minibatches"]
print(f"update: {e0.elapsed_time(e1):.2f} ms for {n} minibatches of {mb} -> {1e3 * e0.elapsed_time(e1) / n:.1f} us per minibatch ({prec})")
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    L.update(obs, act, logp, v, adv, ret, generator=gen)
    torch.cuda.synchronize()
rows = sorted(prof.key_averages(), key=lambda r: -r.device_time_total)
tot = sum(r.device_time_total for r in rows)
print(f"{'kernel':90s} {'calls':>6s} {'total us':>10s} {'us/call':>9s} {'share':>6s}")
for r in rows[:25]:
    print(f"{r.key[:90]:90s} {r.count:6d} {r.device_time_total:10.0f} {r.device_time_total / max(r.count, 1):9.1f} {100 * r.device_time_total / tot:5.1f}%")
print("sum of kernel time per minibatch (us):", tot / n)

# per-launch sequence of one minibatch (kernel order within dn_ppo_minibatch_grad / apply)
evs = [e for e in prof.events() if e.device_time_total > 0 and ("dnmma" in e.name or "dnppo" in e.name)]
evs.sort(key=lambda e: e.time_range.start)
per_mb = len(evs) // n
print("launch sequence of the second minibatch:")
for e in evs[per_mb:2 * per_mb]:
    print(f"  {e.name[:70]:70s} {e.device_time_total:8.1f} us")
