"""The fused actor-critic forward (include/hlynr_policy.h, csrc/hlynr_policy.cu: tcgen05 GEMMs + LayerNorm/ReLU epilogues + heads
in one kernel) against the fp32 PyTorch definition of the reference's network (policy.ReferenceActorCritic:
rl_system/scripts/train_flat_ppo.py:37-85 CustomMLP + SB3 action_net / value_net).

Tolerances (stated, bf16 tensor-core operands with fp32 accumulation):
  * vs a torch forward that rounds weights / input / hidden activations to bf16 at the same points: |diff| <= 5e-3 + 5e-3 |ref|
    (what is left is summation order and the rare 1-ulp bf16 flip of a hidden activation, which the test's x40 action-head gain
    amplifies: worst observed 4.0e-3 over 38 k rows; typical differences are ~1e-7);
  * vs the plain fp32 torch forward: |diff| <= 3e-2 + 3e-2 |ref| (bf16 has 8 mantissa bits; three 512-wide layers).
"""
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _nets(seed=0, scale_heads=True):
    import torch
    from hlynr_intercept_b200.policy import ReferenceActorCritic

    torch.manual_seed(seed)
    net = ReferenceActorCritic(device="cuda")
    with torch.no_grad():   # non-trivial biases, LayerNorm affine parameters, log_std and head gains, so that every operand matters
        for m in net.network:
            if m.__class__.__name__ == "Linear":
                m.bias.normal_(0, 0.2)
            if m.__class__.__name__ == "LayerNorm":
                m.weight.uniform_(0.5, 1.5); m.bias.normal_(0, 0.3)
        if scale_heads:
            net.action_net.weight.mul_(40.0); net.action_net.bias.normal_(0, 0.1)
            net.value_net.bias.fill_(0.37)
        net.log_std.copy_(torch.tensor([-0.5, -0.2, 0.0, 0.1, 0.3, -1.0], device="cuda"))
    return net


@pytest.mark.parametrize("cluster", [1, 2, 4])
@pytest.mark.parametrize("n", [1, 127, 128, 129, 1000, 148 * 128 * 2 + 77])
def test_fused_forward_matches_torch(n, cluster):
    import torch
    from hlynr_intercept_b200.policy import FusedActorCritic, bf16_emulated_forward

    net = _nets()
    fused = FusedActorCritic(net, device=0, seed=5)
    fused.set_option("cluster", cluster)   # thread-block clusters: the CTAs of a cluster share the weight tiles by TMA multicast
    g = torch.Generator(device="cuda"); g.manual_seed(n)
    obs = (torch.randn(n, 104, device="cuda", generator=g) * 1.5).clamp(-10, 10).contiguous()
    obs[:, 7] = 10.0   # a clipped VecNormalize channel
    with torch.no_grad():
        mean_b, value_b = bf16_emulated_forward(net, obs)
        mean_f, value_f = net.mean(obs), net.value(obs)
    m, v = fused.mean_and_value(obs)
    torch.cuda.synchronize()
    m, v = m.clone(), v.clone()
    for got, ref_b, ref_f in ((m, mean_b, mean_f), (v, value_b, value_f)):
        assert torch.isfinite(got).all()
        assert ((got - ref_b).abs() <= 5e-3 + 5e-3 * ref_b.abs()).all(), float((got - ref_b).abs().max())
        assert float((got - ref_b).abs().mean()) < 1e-4
        assert ((got - ref_f).abs() <= 3e-2 + 3e-2 * ref_f.abs()).all(), float((got - ref_f).abs().max())
    # sampling: a = mean + exp(log_std) * eps with eps ~ N(0, 1) from Philox; log pi(a) of the diagonal Gaussian
    a, v2, lp = fused(obs)
    torch.cuda.synchronize()
    assert torch.equal(v2, v)
    eps = (a - m) / net.log_std.exp()
    want_lp = (-0.5 * eps ** 2 - net.log_std - 0.5 * math.log(2 * math.pi)).sum(-1)
    assert ((lp - want_lp).abs() <= 1e-3 + 1e-4 * want_lp.abs()).all()
    a1 = a.clone()
    a2, _, _ = fused(obs)      # a new call counter: new noise
    torch.cuda.synchronize()
    assert not torch.equal(a1, a2)
    if n >= 1000:
        e = eps.detach().flatten().double()
        assert abs(float(e.mean())) < 5 / math.sqrt(e.numel()) and abs(float(e.var()) - 1.0) < 0.05
        assert abs(float((e ** 4).mean()) - 3.0) < 0.3
    ad, _, _ = fused(obs, deterministic=True)
    torch.cuda.synchronize()
    assert torch.equal(ad, m)
    # caller-provided outputs (rows of a rollout buffer) and the clipped action SB3 passes to env.step
    big = torch.zeros(3, n, 6, device="cuda"); vals = torch.zeros(2, n, device="cuda"); clipped = torch.zeros(n, 6, device="cuda")
    lp_out = torch.zeros(n, device="cuda")
    a3, v3, lp3 = fused.forward(obs, out=dict(actions=big[1], values=vals[1], logp=lp_out, clipped=clipped))
    torch.cuda.synchronize()
    assert a3.data_ptr() == big[1].data_ptr() and (big[0] == 0).all() and (big[2] == 0).all() and (vals[0] == 0).all()
    assert torch.equal(vals[1], v) and torch.equal(clipped, big[1].clamp(-1.0, 1.0)) and (clipped.abs() <= 1).all()
    assert not torch.equal(clipped, big[1])   # the test's head gain makes some actions leave [-1, 1]
    fused.close()


def test_device_side_row_count_and_weight_sync():
    import torch
    from hlynr_intercept_b200.policy import FusedActorCritic

    net = _nets(seed=1)
    fused = FusedActorCritic(net, device=0)
    fused.set_option("cluster", 2)
    obs = torch.randn(5000, 104, device="cuda")
    full = fused.value(obs).clone()
    buf = fused._out("values", (5000,))
    for count in (0, 1, 128, 300, 5000, 9999):
        buf.fill_(-7.0)
        cnt = torch.tensor([count], dtype=torch.int32, device="cuda")
        v = fused.value_rows(obs, cnt)
        torch.cuda.synchronize()
        k = min(count, 5000)
        assert torch.equal(v[:k], full[:k])
        done_tiles = -(-k // 128) * 128
        assert (v[min(done_tiles, 5000):] == -7.0).all()   # tiles beyond the device-side count were never touched
    with torch.no_grad():
        net.value_net.bias.add_(1.0)
    fused.sync_weights()
    torch.testing.assert_close(fused.value(obs), full + 1.0, rtol=0, atol=1e-5)
    fused.close()


def test_collector_with_the_fused_policy_matches_torch_policy_path():
    """DeviceRolloutCollector driven by the fused kernel vs the same collector driven by the torch module: same actions noise is
    not shared, so the comparison is on what does not depend on it -- with log_std = -20 both are deterministic up to bf16."""
    import torch
    from hlynr_intercept_b200 import config
    from hlynr_intercept_b200.policy import FusedActorCritic
    from hlynr_intercept_b200.post import HlynrObsPipeline
    from hlynr_intercept_b200.rollout import DeviceRolloutCollector
    from hlynr_intercept_b200.sim import HlynrSim

    net = _nets(seed=2, scale_heads=False)
    with torch.no_grad():
        net.log_std.fill_(-20.0)
    cfg = config.baseline_config("cfg4")
    cfg["max_steps"] = 9     # mass time-outs: every row of the done list needs its bootstrap value
    n, T = 1500, 12
    cols = []
    for pol in (FusedActorCritic(net, device=0), net):
        sim = HlynrSim(cfg, n_envs=n, seed=3, warn_dead=False)
        pipe = HlynrObsPipeline(sim, n_stack=4, training=False)
        col = DeviceRolloutCollector(pipe, pol, T)
        col.collect()
        torch.cuda.synchronize()
        cols.append(col)
    a, b = cols
    assert int(a.overflow.item()) == 0 and int(b.overflow.item()) == 0
    # the first step's observation is identical; values agree to bf16 accuracy there
    assert torch.equal(a.obs[0], b.obs[0])
    assert ((a.values[0] - b.values[0]).abs() <= 3e-2 + 3e-2 * b.values[0].abs()).all()
    # truncated episodes were bootstrapped in both: rewards at the time-limit ticks carry gamma * V(terminal obs)
    assert torch.isfinite(a.advantages).all() and torch.isfinite(a.returns).all()
    t_lim = 8   # ticks 0..8 -> step count 9 = max_steps at index 8
    assert (a.episode_starts[t_lim + 1] == 1).all() and (b.episode_starts[t_lim + 1] == 1).all()
    d = (a.rewards[t_lim] - b.rewards[t_lim]).abs()
    assert float(d.max()) < 0.1 and float((a.rewards[t_lim] - a.rewards[t_lim - 1]).abs().mean()) > 1e-3
