"""Batched IDHP agent with the API of the reference's DHPAgent / DHPActor / DHPCritic
(/root/reference/agents.py:8-209), backed by libheli3dof.so.

All six weight matrices of an agent live in ONE `(11*nh, N)` struct-of-arrays CUDA tensor (`agent.net`,
layout in include/heli3dof.h); `actor.fc1.weight` etc. are views of shape (out, in, N).  The update is
the closed form of agents.py:152-196 (no autograd).
"""
import numpy as np
import torch

from . import _lib as L
from ._config import as_rows, build_batch, build_config, byref, is_per_episode, resolve_device


class _Layer:
    def __init__(self, weight):
        self.weight = weight


class _Net:
    """Two bias-free layers (agents.py:19-22, :42-45) as views into the agent's SoA block."""

    def __init__(self, fc1, fc2):
        self.fc1 = _Layer(fc1)
        self.fc2 = _Layer(fc2)

    def parameters(self):
        return iter([self.fc1.weight, self.fc2.weight])

    def state_dict(self):
        return {"fc1.weight": self.fc1.weight.detach().cpu().clone(), "fc2.weight": self.fc2.weight.detach().cpu().clone()}

    def load_state_dict(self, sd):
        for name, layer in (("fc1.weight", self.fc1), ("fc2.weight", self.fc2)):
            w = torch.as_tensor(sd[name])
            if w.dim() == 2:        # a reference-format checkpoint (no episode axis): broadcast
                w = w.unsqueeze(-1)
            layer.weight.copy_(w.to(layer.weight).expand_as(layer.weight))

    def zero_grad(self):
        pass


class DHPCritic(_Net):
    """agents.py:8-27"""


class DHPActor(_Net):
    """agents.py:30-63"""

    def __init__(self, fc1, fc2, scaling, action_network_final_layer):
        super().__init__(fc1, fc2)
        self.scale = np.deg2rad(scaling) if scaling is not None else 1     # agents.py:46-49
        if action_network_final_layer not in L.ACTIVATIONS:
            raise NotImplementedError(f"Activation function {action_network_final_layer} not supported. ")
        self.final_layer = action_network_final_layer


class DHPAgent:
    """agents.py:66-209.  Extra keyword arguments: n_envs, device, dtype (network precision), init
    ("reference": consume torch's global CPU RNG exactly as the reference constructor does, so
    torch.manual_seed(s) reproduces its weights; "batched": one torch.randn per matrix), generator."""

    def __init__(self, control_channel: str, discount_factor,
                 n_hidden_actor: int, nn_stdev_actor, learning_rate_actor, action_scaling, action_network_final_layer,
                 n_hidden_critic: int, nn_stdev_critic, learning_rate_critic, tau_target_critic,
                 tracked_state: int, ac_states: list,
                 reward_weight=1, n_envs=1, device=None, dtype=torch.float64, init="auto", generator=None):
        self.control_channel_str = control_channel
        self.control_channel_mapping = {'col': 0, 'lon': 1, 'lat': 2, 'ped': 3}
        self.control_channel_num = self.control_channel_mapping[control_channel]
        if self.control_channel_num > 1:
            raise NotImplementedError("the 3-DOF path has the channels 'col' and 'lon' only")
        self.tracked_state = tracked_state
        self.ac_states = ac_states
        self.n_inputs = len(ac_states) + 1
        if self.n_inputs != 2 or n_hidden_actor != n_hidden_critic:
            raise NotImplementedError("libheli3dof implements the 3-DOF agents: one auxiliary state + the tracked "
                                      "error (main_3dof.py:31-46) and equal actor / critic widths")
        self.n_hidden = nh = int(n_hidden_actor)
        self.n_envs = n = int(n_envs)
        self.device = resolve_device(device)
        self.dtype = dtype
        L.dtype_code(dtype)
        self.final_layer = action_network_final_layer

        self.net = torch.zeros(11 * nh, n, dtype=dtype, device=self.device)
        v = self.net
        self.actor = DHPActor(v[0:2 * nh].view(nh, 2, n), v[2 * nh:3 * nh].view(1, nh, n), action_scaling,
                              action_network_final_layer)
        self.critic = DHPCritic(v[3 * nh:5 * nh].view(nh, 2, n), v[5 * nh:7 * nh].view(2, nh, n))
        self.target_critic = DHPCritic(v[7 * nh:9 * nh].view(nh, 2, n), v[9 * nh:11 * nh].view(2, nh, n))
        self.networks = [self.actor, self.critic, self.target_critic]
        self.nn_stdev_actor, self.nn_stdev_critic = nn_stdev_actor, nn_stdev_critic
        self.hyper_rows_device = None       # optional (6, N) f64 device rows (HYPER_NAMES order): used as they are
        # whether target critic == critic is KNOWN (True after construction, agents.py:107; None after a load: the fused
        # kernel's tau = 1 shortcut does not carry the target critic, so ClosedLoop3DOF checks once before relying on it)
        self.target_synced = True
        if init == "auto":
            init = "reference" if n == 1 else "batched"
        self._init_weights(init, nn_stdev_actor, nn_stdev_critic, generator)

        self.learning_rate_actor = learning_rate_actor
        self.learning_rate_critic = learning_rate_critic
        self.gamma = discount_factor
        self.tau_target_critic = tau_target_critic
        self.reward_weight = reward_weight

    # -- initialisation ---------------------------------------------------------------------------
    def _init_weights(self, init, std_a, std_c, generator):
        nh, n = self.n_hidden, self.n_envs
        shapes = ((self.actor.fc1, (nh, 2), std_a), (self.actor.fc2, (1, nh), std_a),
                  (self.critic.fc1, (nh, 2), std_c), (self.critic.fc2, (2, nh), std_c))
        if init == "reference":
            # agents.py:19-22 / :42-45: nn.Linear's own init draws first, then normal_ overwrites; both
            # consume the global CPU generator.  fp32 draws as in the original stack, cast afterwards.
            for ep in range(n):
                for layer, (o, i), std in shapes:
                    lin = torch.nn.Linear(i, o, bias=False, dtype=torch.float32)
                    torch.nn.init.normal_(lin.weight, mean=0, std=std)
                    layer.weight[:, :, ep] = lin.weight.detach().to(self.dtype)
        elif init == "batched":
            for layer, (o, i), std in shapes:
                w = torch.randn(o, i, n, generator=generator, dtype=torch.float64) * std
                layer.weight.copy_(w.to(self.dtype))
        elif init == "device":
            return                          # ClosedLoop3DOF.init_episodes draws the weights on the device
        else:
            raise ValueError("init must be 'auto', 'reference', 'batched' or 'device'")
        self.target_critic.fc1.weight.copy_(self.critic.fc1.weight)      # copy.deepcopy, agents.py:107
        self.target_critic.fc2.weight.copy_(self.critic.fc2.weight)

    # -- C-ABI plumbing ----------------------------------------------------------------------------
    def _hyper_values(self):
        return {"gamma": self.gamma, "lr_actor": self.learning_rate_actor, "lr_critic": self.learning_rate_critic,
                "tau": self.tau_target_critic, "scale": self.actor.scale, "reward_weight": self.reward_weight}

    def _hyper_rows(self):
        """(6, N) f64 per-episode hyper-parameter rows, or None when every value is a scalar."""
        if self.hyper_rows_device is not None:
            return self.hyper_rows_device
        vals = self._hyper_values()
        if not any(is_per_episode(v) for v in vals.values()):
            return None
        # the rows are rebuilt only when one of the attributes was re-bound or modified in place
        key = tuple((id(v), v._version if torch.is_tensor(v) else None) for v in vals.values())
        cacheable = not any(isinstance(v, np.ndarray) for v in vals.values())   # in-place numpy edits are invisible
        cached = getattr(self, "_hyper_cache", None)
        if cacheable and cached is not None and cached[0] == key:
            return cached[1]
        rows = []
        for k in L.HYPER_NAMES:
            v = vals[k]
            v = v.detach().to("cpu", torch.float64) if torch.is_tensor(v) else torch.as_tensor(np.asarray(v, np.float64))
            rows.append(v.reshape(-1).expand(self.n_envs))
        out = torch.stack(rows).to(self.device).contiguous()
        self._hyper_cache = (key, out, list(vals.values()))      # keep the keyed objects alive (ids stay unique)
        return out

    def _which(self):
        return self.control_channel_num

    def _call_setup(self, env_dtype, rls=None):
        kw = {"col": self} if self._which() == 0 else {"lon": self}
        cfg = build_config(rls=rls, env_dtype=env_dtype, net_dtype=self.dtype, **kw)
        hyper = self._hyper_rows()
        bk = {"col_net": self.net, "col_hyper": hyper} if self._which() == 0 else {"lon_net": self.net, "lon_hyper": hyper}
        if rls is not None:
            bk["rls_W"] = rls._W
        return cfg, build_batch(self.n_envs, **bk), hyper

    def _ref2(self, ref, dtype):
        """(7,N) reference vector (heli_models.py:236) or its (2,N) [z, theta] slots -> contiguous (2,N)."""
        n = self.n_envs
        if not torch.is_tensor(ref):
            ref = torch.as_tensor(np.asarray(ref, dtype=np.float64))
        if ref.dim() == 1:
            ref = ref.reshape(-1, 1).expand(-1, n)
        if ref.shape[0] == 7:
            ref = ref[[1, 4]]
        return as_rows(ref, 2, n, dtype, self.device, "ref")

    # -- reference API ----------------------------------------------------------------------------------
    def get_action(self, state, ref):
        """agents.py:117-121 -> (N,) action in radians (network precision)"""
        state = as_rows(state, 7, self.n_envs, state.dtype if torch.is_tensor(state) else torch.float64,
                        self.device, "state")
        ref2 = self._ref2(ref, state.dtype)
        out = torch.empty(self.n_envs, dtype=self.dtype, device=self.device)
        cfg, b, _keep = self._call_setup(state.dtype)
        L.check(L.lib().h3_agent_action(byref(cfg), byref(b), self._which(), L.ptr(state), L.ptr(ref2), L.ptr(out),
                                        L.stream_ptr(self.device)))
        return out

    def get_reward(self, next_state, ref):
        """agents.py:123-128 -> (reward (N,), dr_ds (1, 2, N))"""
        next_state = as_rows(next_state, 7, self.n_envs,
                             next_state.dtype if torch.is_tensor(next_state) else torch.float64, self.device,
                             "next_state")
        ref2 = self._ref2(ref, next_state.dtype)
        reward = torch.empty(self.n_envs, dtype=next_state.dtype, device=self.device)
        dr_ds = torch.empty(2, self.n_envs, dtype=self.dtype, device=self.device)
        cfg, b, _keep = self._call_setup(next_state.dtype)
        L.check(L.lib().h3_agent_reward(byref(cfg), byref(b), self._which(), L.ptr(next_state), L.ptr(ref2),
                                        L.ptr(reward), L.ptr(dr_ds), L.stream_ptr(self.device)))
        return reward, dr_ds.unsqueeze(0)

    def get_transition_matrices(self, rls_estimator):
        """agents.py:130-140 -> F (2, 2, N), G (2, 1, N) in network precision"""
        F = torch.empty(4, self.n_envs, dtype=self.dtype, device=self.device)
        G = torch.empty(2, self.n_envs, dtype=self.dtype, device=self.device)
        cfg, b, _keep = self._call_setup(rls_estimator.dtype, rls=rls_estimator)
        L.check(L.lib().h3_agent_transition(byref(cfg), byref(b), self._which(), L.ptr(F), L.ptr(G),
                                            L.stream_ptr(self.device)))
        return F.view(2, 2, self.n_envs), G.view(2, 1, self.n_envs)

    def augment_state(self, state, reference):
        """agents.py:142-150 -> (2, N)"""
        ri = self.tracked_state
        return torch.stack([state[x] for x in self.ac_states] + [state[ri] - reference[ri]]).to(self.dtype)

    def update_networks(self, state, next_state, ref, next_ref, dr_ds, F, G):
        """agents.py:152-196 (forward passes, actor backward, critic error, SGD, soft target update)."""
        n, dev = self.n_envs, self.device
        edt = state.dtype if torch.is_tensor(state) else torch.float64
        state = as_rows(state, 7, n, edt, dev, "state")
        next_state = as_rows(next_state, 7, n, edt, dev, "next_state")
        ref2, nref2 = self._ref2(ref, edt), self._ref2(next_ref, edt)
        dr = as_rows(dr_ds.reshape(2, -1), 2, n, self.dtype, dev, "dr_ds")
        Fm = as_rows(F.reshape(4, -1), 4, n, self.dtype, dev, "F")
        Gm = as_rows(G.reshape(2, -1), 2, n, self.dtype, dev, "G")
        cfg, b, _keep = self._call_setup(edt)
        L.check(L.lib().h3_agent_update(byref(cfg), byref(b), self._which(), L.ptr(state), L.ptr(next_state),
                                        L.ptr(ref2), L.ptr(nref2), L.ptr(dr), L.ptr(Fm), L.ptr(Gm),
                                        L.stream_ptr(dev)))

    def save(self, path):
        """agents.py:198-203.  With n_envs == 1 the file has the reference's format (no episode axis)."""
        def sd(net):
            d = net.state_dict()
            return {k: (v.squeeze(-1) if self.n_envs == 1 else v) for k, v in d.items()}
        torch.save({"actor_state_dict": sd(self.actor), "critic_state_dict": sd(self.critic),
                    "target_critic_state_dict": sd(self.target_critic)}, path)

    def load(self, path):
        """agents.py:205-209"""
        checkpoint = torch.load(path, map_location="cpu")
        self.actor.load_state_dict(checkpoint["actor_state_dict"])
        self.critic.load_state_dict(checkpoint["critic_state_dict"])
        self.target_critic.load_state_dict(checkpoint["target_critic_state_dict"])
        self.target_synced = None
