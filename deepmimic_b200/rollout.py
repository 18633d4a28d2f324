"""Batched policy rollout over the batched environment: what the reference does per env and per MPI worker
in RLWorld.update_agents -> RLAgent._update_new_action (R/learning/rl_world.py:94-132, R/learning/rl_agent.py:319-343), done for all
N environments of a rank at once with observations, actions and rewards staying on the device.

  DeviceNormalizer   R/learning/normalizer.py (mean / std / clip, group-wise running statistics), torch tensors
  GaussianMLPPolicy  actor of PPOAgent: fc_2layers_1024units -> Gaussian mean (+ state-independent log-std bias = log(noise))
                     (R/learning/ppo_agent.py:52-90, pg_agent.py:140-160, nets/fc_2layers_1024units.py, tf_util.py:27-39)
  BatchedRollout     record_state -> normalise -> actor -> un-normalise -> set_action -> 20 x update -> reward / flags -> masked reset,
                     collecting [T, N, .] trajectory tensors for a learner

The MLP runs as plain torch matmuls (cuBLAS): a library GEMM, not part of the hand-written hot path.  The reference's TF1 checkpoints
are read by deepmimic_b200/tf_checkpoint.py (TensorBundle reader, no TensorFlow) and loaded with load_actor_weights; without a
checkpoint the weights are random-initialised the way the reference initialises them."""
import math

import numpy as np


class DeviceNormalizer:
    NORM_GROUP_SINGLE = 0
    NORM_GROUP_NONE = -1

    def __init__(self, size, group_ids=None, eps=0.02, clip=float("inf"), device="cpu"):
        import torch
        self.torch = torch
        self.eps, self.clip = eps, clip
        self.mean = torch.zeros(size, device=device)
        self.mean_sq = torch.zeros(size, device=device)
        self.std = torch.ones(size, device=device)
        self.count = 0
        g = np.zeros(size, dtype=np.int64) if group_ids is None else np.asarray(group_ids, dtype=np.int64)
        self.group_ids = g
        self.new_count = 0
        self.new_sum = torch.zeros(size, device=device)
        self.new_sum_sq = torch.zeros(size, device=device)

    def set_mean_std(self, mean, std):
        t = self.torch
        self.mean = t.as_tensor(np.asarray(mean), dtype=t.float32, device=self.mean.device).clone()
        self.std = t.as_tensor(np.asarray(std), dtype=t.float32, device=self.mean.device).clone()
        self.mean_sq = self.std * self.std + self.mean * self.mean

    def normalize(self, x):
        y = (x - self.mean) / self.std
        return y if math.isinf(self.clip) else y.clamp(-self.clip, self.clip)

    def unnormalize(self, y):
        return y * self.std + self.mean

    def record(self, x):
        x = x.reshape(-1, self.mean.numel())
        self.new_count += x.shape[0]
        self.new_sum += x.sum(dim=0)
        self.new_sum_sq += (x * x).sum(dim=0)

    def _process_group_data(self, new, old):
        out = new.clone()
        for gid in np.unique(self.group_ids):
            idx = self.torch.as_tensor(np.nonzero(self.group_ids == gid)[0], device=new.device)
            if gid == self.NORM_GROUP_NONE:
                out[idx] = old[idx]
            elif gid != self.NORM_GROUP_SINGLE:
                out[idx] = new[idx].mean()
        return out

    def update(self, all_reduce=None):
        """Fold the recorded samples into the running statistics; `all_reduce(tensor)` sums over ranks (torch.distributed) if given."""
        t = self.torch
        cnt = t.tensor([float(self.new_count)], device=self.mean.device)
        s, sq = self.new_sum.clone(), self.new_sum_sq.clone()
        if all_reduce is not None:
            for x in (cnt, s, sq):
                all_reduce(x)
        n = int(cnt.item())
        if n > 0:
            total = self.count + n
            new_mean = self._process_group_data(s / n, self.mean)
            new_mean_sq = self._process_group_data(sq / n, self.mean_sq)
            w_old, w_new = self.count / total, n / total
            self.mean = w_old * self.mean + w_new * new_mean
            self.mean_sq = w_old * self.mean_sq + w_new * new_mean_sq
            self.count = total
            self.std = t.sqrt((self.mean_sq - self.mean * self.mean).clamp_min(0)).clamp_min(self.eps)
        self.new_count = 0
        self.new_sum.zero_(); self.new_sum_sq.zero_()


def build_policy(state_size, action_size, init_output_scale=0.01, noise=0.05, hidden=(1024, 512)):
    import torch

    class GaussianMLPPolicy(torch.nn.Module):
        def __init__(self):
            super().__init__()
            dims = [state_size] + list(hidden)
            self.hidden = torch.nn.ModuleList([torch.nn.Linear(a, b) for a, b in zip(dims[:-1], dims[1:])])
            for l in self.hidden:
                torch.nn.init.xavier_uniform_(l.weight); torch.nn.init.zeros_(l.bias)
            self.mean = torch.nn.Linear(dims[-1], action_size)
            torch.nn.init.uniform_(self.mean.weight, -init_output_scale, init_output_scale); torch.nn.init.zeros_(self.mean.bias)
            self.logstd = torch.nn.Parameter(torch.full((action_size,), math.log(noise)))

        def forward(self, norm_s):
            h = norm_s
            for l in self.hidden:
                h = torch.relu(l(h))      # fc_net leaves the last layer linear, build_net applies the activation afterwards
            return self.mean(h)

        def sample(self, norm_s, explore_mask=None, generator=None):
            """Normalised action and its log-probability; rows with explore_mask False take the mode."""
            mu = self.forward(norm_s)
            std = self.logstd.exp()
            eps = torch.randn(mu.shape, device=mu.device, generator=generator)
            if explore_mask is not None:
                eps = eps * explore_mask[:, None].to(eps.dtype)
            a = mu + std * eps
            logp = (-0.5 * eps * eps - self.logstd - 0.5 * math.log(2 * math.pi)).sum(dim=-1)
            return a, logp

    return GaussianMLPPolicy()


def build_gated_policy(state_size, goal_size, action_size, init_output_scale=0.01, noise=0.05, hidden=(1024, 512), gate_common=128, gate_hidden=64):
    """The goal-conditioned actor of the reference's AMP task agents, `fc_2layers_gated_1024units`
    (R/learning/nets/fc_2layers_gated_1024units.py:6-58): the trunk sees [norm_s, norm_g]; every hidden layer's pre-activation is scaled by
    2*sigmoid(.) and shifted by a bias, both computed from the normalised goal through gate_common (128, relu) and a 64-unit relu layer."""
    import torch

    class GatedGaussianMLPPolicy(torch.nn.Module):
        def __init__(self):
            super().__init__()
            dims = [state_size + goal_size] + list(hidden)
            lin = torch.nn.Linear
            self.goal_size = goal_size
            self.hidden = torch.nn.ModuleList([lin(a, b) for a, b in zip(dims[:-1], dims[1:])])
            self.gate_common = lin(goal_size, gate_common)
            self.gate_hidden = torch.nn.ModuleList([lin(gate_common, gate_hidden) for _ in hidden])
            self.gate_bias = torch.nn.ModuleList([lin(gate_hidden, h) for h in hidden])
            self.gate_scale = torch.nn.ModuleList([lin(gate_hidden, h) for h in hidden])
            for l in list(self.hidden) + [self.gate_common] + list(self.gate_hidden) + list(self.gate_bias) + list(self.gate_scale):
                torch.nn.init.xavier_uniform_(l.weight); torch.nn.init.zeros_(l.bias)
            self.mean = lin(dims[-1], action_size)
            torch.nn.init.uniform_(self.mean.weight, -init_output_scale, init_output_scale); torch.nn.init.zeros_(self.mean.bias)
            self.logstd = torch.nn.Parameter(torch.full((action_size,), math.log(noise)))

        def forward(self, norm_s, norm_g):
            gc = torch.relu(self.gate_common(norm_g))
            h = torch.cat([norm_s, norm_g], dim=-1)
            for l, gh, gb, gs in zip(self.hidden, self.gate_hidden, self.gate_bias, self.gate_scale):
                gate = torch.relu(gh(gc))
                h = torch.relu(2.0 * torch.sigmoid(gs(gate)) * l(h) + gb(gate))
            return self.mean(h)

        def sample(self, norm_s, norm_g, explore_mask=None, generator=None):
            mu = self.forward(norm_s, norm_g)
            std = self.logstd.exp()
            eps = torch.randn(mu.shape, device=mu.device, generator=generator)
            if explore_mask is not None:
                eps = eps * explore_mask[:, None].to(eps.dtype)
            a = mu + std * eps
            logp = (-0.5 * eps * eps - self.logstd - 0.5 * math.log(2 * math.pi)).sum(dim=-1)
            return a, logp

    return GatedGaussianMLPPolicy()


def load_actor_weights(policy, actor):
    """Copies a reference actor (deepmimic_b200.tf_checkpoint.load_actor, or the tests/golden fixture keys w0 b0 w1 b1 wm bm logstd) into a
    GaussianMLPPolicy.  TF dense kernels are [in, out]; torch Linear weights are [out, in]."""
    import torch
    if "hidden" in actor:
        hidden, mean, logstd = actor["hidden"], actor["mean"], actor["logstd"]
    else:
        hidden, mean, logstd = [(actor["w0"], actor["b0"]), (actor["w1"], actor["b1"])], (actor["wm"], actor["bm"]), actor["logstd"]
    with torch.no_grad():
        for layer, (w, b) in zip(policy.hidden, hidden):
            layer.weight.copy_(torch.as_tensor(np.asarray(w, dtype=np.float32)).t()); layer.bias.copy_(torch.as_tensor(np.asarray(b, dtype=np.float32)))
        policy.mean.weight.copy_(torch.as_tensor(np.asarray(mean[0], dtype=np.float32)).t()); policy.mean.bias.copy_(torch.as_tensor(np.asarray(mean[1], dtype=np.float32)))
        policy.logstd.copy_(torch.as_tensor(np.asarray(logstd, dtype=np.float32)))
        if "gate_common" in actor:
            put = lambda layer, wb: (layer.weight.copy_(torch.as_tensor(np.asarray(wb[0], dtype=np.float32)).t()), layer.bias.copy_(torch.as_tensor(np.asarray(wb[1], dtype=np.float32))))
            put(policy.gate_common, actor["gate_common"])
            for i, g in enumerate(actor["gates"]):
                put(policy.gate_hidden[i], g["hidden"]); put(policy.gate_bias[i], g["bias"]); put(policy.gate_scale[i], g["scale"])
    return policy


class _nullcontext:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


class BatchedRollout:
    """backend "torch": the policy is a torch module (cuBLAS GEMMs, eager normalisers) -- needed for training and for the gated task actor.
    backend "tcgen05": inference of the plain 2-layer actor on the library's own tensor-core kernels (dm_mlp_*, kernels/dm_mlp.cu): normaliser,
    three GEMMs, bias / ReLU and the action un-normalisation in four launches (operand preparation + one per layer) on the environment's stream; the weights and the normaliser
    statistics are snapshotted by refresh_tensor_core_policy() (call it again after a learner update)."""

    def __init__(self, env, policy=None, exp_rate=1.0, noise=0.05, seed=0, backend="torch"):
        import torch
        self.torch, self.env = torch, env
        dev = env.device
        S, A, G = env.get_state_size(), env.get_action_size(), env.get_goal_size()
        self.goal_size = G
        # goal-conditioned scenes (AMP tasks) use the reference's gated actor; everything else the plain 1024-512 MLP
        self.policy = (policy or (build_gated_policy(S, G, A, noise=noise) if G > 0 else build_policy(S, A, noise=noise))).to(dev)
        if G > 0:
            self.g_norm = DeviceNormalizer(G, env.build_goal_norm_groups(), device=dev)
            self.g_norm.set_mean_std(-env.build_goal_offset(), 1.0 / env.build_goal_scale())
        self.s_norm = DeviceNormalizer(S, env.build_state_norm_groups(), device=dev)
        self.s_norm.set_mean_std(-env.build_state_offset(), 1.0 / env.build_state_scale())
        self.a_norm = DeviceNormalizer(A, device=dev)
        self.a_norm.set_mean_std(-env.build_action_offset(), 1.0 / env.build_action_scale())
        self.exp_rate = exp_rate
        self.gen = torch.Generator(device=dev); self.gen.manual_seed(seed)
        self.backend, self._tc = backend, None
        if backend not in ("torch", "tcgen05"):
            raise ValueError("backend must be 'torch' or 'tcgen05'")
        if backend == "tcgen05" and G > 0:
            raise ValueError("the tcgen05 backend implements the plain 2-layer actor; goal-conditioned (gated) actors run on the torch backend")

    def refresh_tensor_core_policy(self):
        """(re)builds the dm_mlp handle from the current torch policy and normalisers"""
        from .capi import TensorCoreMLP
        pol, env = self.policy, self.env
        if len(pol.hidden) != 2:
            raise ValueError("the tcgen05 backend implements exactly two hidden layers")
        g = lambda t: t.detach().float().cpu().numpy()
        if self._tc is not None:
            self._tc.close()
        self._tc = TensorCoreMLP(g(pol.hidden[0].weight).T, g(pol.hidden[0].bias), g(pol.hidden[1].weight).T, g(pol.hidden[1].bias), g(pol.mean.weight).T, g(pol.mean.bias),
                                 in_mean=g(self.s_norm.mean), in_std=g(self.s_norm.std), in_clip=self.s_norm.clip, out_mean=g(self.a_norm.mean), out_std=g(self.a_norm.std),
                                 max_rows=env.num_envs, device=env.device.index or 0)
        self._tc_act = self.torch.empty(env.num_envs, env.get_action_size(), device=env.device)
        return self._tc

    def _act_tensor_core(self, s, explore):
        """un-normalised actions and log-probabilities from the tensor-core actor (exploration noise is drawn in torch, added in the kernel's epilogue)"""
        t = self.torch
        if self._tc is None:
            self.refresh_tensor_core_policy()
        std = self.policy.logstd.detach().exp()
        eps = t.randn(s.shape[0], std.shape[0], device=s.device, generator=self.gen) * explore[:, None].to(s.dtype)
        noise = (std * eps).contiguous()
        cur = t.cuda.current_stream(s.device)
        self._tc.forward(s.contiguous(), self._tc_act, noise=noise, stream=cur.cuda_stream)
        logp = (-0.5 * eps * eps - self.policy.logstd.detach() - 0.5 * math.log(2 * math.pi)).sum(dim=-1)
        return self._tc_act, logp

    @property
    def stream(self):
        return self.env.stream

    def collect(self, num_steps, record_stats=True):
        """num_steps policy steps of all environments; returns dict of [T, N, .] tensors (states, actions, logps, rewards, dones)."""
        t, env = self.torch, self.env
        N, S, A = env.num_envs, env.get_state_size(), env.get_action_size()
        out = dict(states=t.empty(num_steps, N, S, device=env.device), actions=t.empty(num_steps, N, A, device=env.device),
                   logps=t.empty(num_steps, N, device=env.device), rewards=t.empty(num_steps, N, device=env.device),
                   dones=t.empty(num_steps, N, dtype=t.bool, device=env.device), terminate=t.empty(num_steps, N, dtype=t.int32, device=env.device))
        G = self.goal_size
        if G > 0:
            out["goals"] = t.empty(num_steps, N, G, device=env.device)
        # the whole loop runs on the environment's stream: with the actor and the bookkeeping on another stream every env call is a pair of
        # cross-stream event waits (measured: 0.4 ms of bubbles per policy step); the caller's stream waits for the trajectory at the end
        caller = t.cuda.current_stream(env.device) if env.device.type == "cuda" else None
        if caller is not None:
            env.stream.wait_stream(caller)
        with t.no_grad(), (t.cuda.stream(env.stream) if caller is not None else _nullcontext()):
            s = env.record_state()
            for k in range(num_steps):
                out["states"][k] = s
                if record_stats:
                    self.s_norm.record(s)
                explore = t.rand(N, device=env.device, generator=self.gen) < self.exp_rate
                if G > 0:   # RLAgent._update_new_action records the goal next to the state (R/learning/rl_agent.py:319-343)
                    g = env.record_goal()
                    out["goals"][k] = g
                    if record_stats:
                        self.g_norm.record(g)
                    na, logp = self.policy.sample(self.s_norm.normalize(s), self.g_norm.normalize(g), explore, self.gen)
                elif self.backend == "tcgen05":
                    a, logp = self._act_tensor_core(s, explore)
                else:
                    na, logp = self.policy.sample(self.s_norm.normalize(s), explore, self.gen)
                if not (G == 0 and self.backend == "tcgen05"):
                    a = self.a_norm.unnormalize(na).contiguous()
                s, r, done, term = env.step(a)
                out["actions"][k] = a; out["logps"][k] = logp; out["rewards"][k] = r; out["dones"][k] = done; out["terminate"][k] = term
                env.reset()                # restarts exactly the finished episodes
                # the restarted environments need the observation of their new state.  Unconditional (one more ~10 us observation kernel) instead of
                # `if done.any()`: that test is a host synchronisation per policy step, which leaves the GPU idle while the host launches the
                # next step's small kernels (measured: 1.52 M -> see tests/test_mlp_gpu.py for the current rates)
                s = env.record_state()
        if caller is not None:
            caller.wait_stream(env.stream)
        return out
