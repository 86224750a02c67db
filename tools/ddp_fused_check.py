"""torchrun --nproc-per-node 2 tools/ddp_fused_check.py -- the fused PPO update under NCCL: one all-reduce of the flat bucket per
minibatch (gradients + early-stop vote), parameters bit-identical across ranks, ranks stop together."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from drl_dronenavigation_b200.ppo import PPOConfig, PPOLearner

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)


def rollout(L, B, seed):
    g = torch.Generator(device=dev).manual_seed(seed)
    obs = torch.randn(B, 13, device=dev, generator=g)
    a, logp, v = L.act(obs, generator=g)
    logp = logp + 0.1 * torch.randn(B, device=dev, generator=g)
    adv, ret = torch.randn(B, device=dev, generator=g), v + torch.randn(B, device=dev, generator=g)
    return obs, a, logp, v, adv, ret


def gathered(t):
    out = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(out, t)
    return out

# (1) plain update: different data per rank, identical parameters afterwards, one collective per minibatch
L = PPOLearner(13, 4, PPOConfig(batch_size=2048, n_epochs=2, target_kl=None, update_impl="fused"), device=dev)
p0 = gathered(L.flat_parameters())
assert all(torch.equal(p0[0], p) for p in p0), "initial parameters differ"
ro = rollout(L, 8192, seed=100 + rank)
out = L.update(*ro, generator=torch.Generator(device=dev).manual_seed(7))
p1 = gathered(L.flat_parameters())
assert all(torch.equal(p1[0], p) for p in p1), "parameters diverged across ranks"
assert not torch.equal(p0[0], p1[0])
assert L.allreduce_calls == out["minibatches"] == 8 and out["optimizer_steps"] == 8, (L.allreduce_calls, out)
# the step is Adam on the MEAN of the ranks' gradients: compare with a single-rank learner fed both shards' gradients
# (checked through the bucket: after the all-reduce every rank holds the same summed bucket)
b = gathered(L._bucket.clone())
assert all(torch.equal(b[0], x) for x in b)

# (2) early stop: only the last rank exceeds its KL threshold, every rank must stop at the same minibatch
# (target_kl differs per rank only to make ONE rank vote; it must not be None on any rank: that would skip the stop checks there)
L2 = PPOLearner(13, 4, PPOConfig(batch_size=1024, n_epochs=20, learning_rate=1e-2, target_kl=(1e-6 if rank == world - 1 else 1e9), update_impl="fused"), device=dev)
ro = rollout(L2, 8192, seed=200 + rank)
out2 = L2.update(*ro, generator=torch.Generator(device=dev).manual_seed(9))
stats = gathered(torch.tensor([float(out2["minibatches"]), float(out2["optimizer_steps"]), float(out2["early_stop"])], device=dev))
assert all(torch.equal(stats[0], s) for s in stats), stats
assert out2["early_stop"] and out2["minibatches"] == out2["optimizer_steps"] + 1, out2
calls = gathered(torch.tensor([float(L2.allreduce_calls), float(out2["launched_minibatches"])], device=dev))
assert all(torch.equal(calls[0], c) for c in calls), calls          # every rank issued the same number of collectives
p2 = gathered(L2.flat_parameters())
assert all(torch.equal(p2[0], p) for p in p2)
# (3) the peer-memory all-reduce itself against ncclAllReduce on the same bucket
fu = L.fused
impl = L.allreduce_impl
if impl == "peer":
    g = torch.Generator(device=dev).manual_seed(1000 + rank)
    for trial in range(5):
        L._bucket.copy_(torch.randn(L._bucket.shape, device=dev, generator=g))
        want = L._bucket.clone()
        dist.all_reduce(want, op=dist.ReduceOp.SUM)
        fu.allreduce()
        torch.cuda.synchronize()
        got = gathered(L._bucket.clone())
        assert all(torch.equal(got[0], x) for x in got), "peer all-reduce: ranks disagree"
        if world == 2:
            assert torch.equal(got[0], want), float((got[0] - want).abs().max())       # two addends: the sum does not depend on the order
        else:
            assert torch.allclose(got[0], want, rtol=1e-5, atol=1e-5), float((got[0] - want).abs().max())

# (3b) the collective alone, back to back
if impl == "peer":
    def loop(fn, n=200):
        for _ in range(20):
            fn()
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) * 1e3 / n
    L._bucket.zero_()
    t_peer = loop(fu.allreduce)
    t_nccl = loop(lambda: dist.all_reduce(L._bucket, op=dist.ReduceOp.SUM))
    g_ar = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g_ar):
        for _ in range(10):
            fu.allreduce()
    t_peer_graph = loop(g_ar.replay, n=40) / 10
    if rank == 0:
        print(f"all-reduce of {L._bucket.numel()} floats alone, us per call: peer memory {t_peer:.1f} (launched), {t_peer_graph:.1f} (graph replay); ncclAllReduce {t_nccl:.1f}")

# (4) time per optimiser step at the bench shape, peer memory vs ncclAllReduce (same data, same graphs otherwise)
def timed(env_value):
    os.environ["DN_PPO_ALLREDUCE"] = env_value
    Lt = PPOLearner(13, 4, PPOConfig(batch_size=32768, n_epochs=2, target_kl=None, update_impl="fused"), device=dev)
    rt = rollout(Lt, 1 << 19, seed=300 + rank)
    gen = torch.Generator(device=dev).manual_seed(11)
    Lt.update(*rt, generator=gen)
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    o = Lt.update(*rt, generator=gen)
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) * 1e3 / o["minibatches"]], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    pf = gathered(Lt.flat_parameters())
    assert all(torch.equal(pf[0], p) for p in pf)
    return float(t.item()), Lt.allreduce_impl, pf[0]
us_peer, impl_peer, p_peer = timed("peer")
us_nccl, impl_nccl, p_nccl = timed("nccl")
assert impl_nccl == "nccl"
rel = float((p_peer - p_nccl).abs().max() / p_nccl.abs().max())
assert rel < 1e-4, rel          # same update up to the order of the cross-rank sums
if rank == 0:
    print(f"ddp fused check ok: world {world}, all-reduce impl {impl}, update {out['minibatches']} minibatches / {L.allreduce_calls} all-reduces, "
          f"early stop after {out2['minibatches']} minibatches ({out2['optimizer_steps']} applied), launched {out2['launched_minibatches']}")
    print(f"us per optimiser step (32768-sample minibatch, max over ranks): {impl_peer} {us_peer:.1f}, {impl_nccl} {us_nccl:.1f}; "
          f"parameters after 32 steps differ by {rel:.1e} (order of the sums)")
dist.barrier()
dist.destroy_process_group()
