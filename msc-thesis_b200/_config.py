"""Builds the C-ABI structs (H3Config / H3Batch) from the host objects."""
import ctypes as C

import numpy as np
import torch

from . import _lib as L


def resolve_device(device=None):
    if device is not None:
        return torch.device(device)
    if torch.cuda.is_available():
        return torch.device("cuda", torch.cuda.current_device())
    # objects can be built on the CPU for host-side plumbing (config, checkpoints, sharding);
    # every compute call then fails loudly in _lib.ptr()
    return torch.device("cpu")


def as_rows(x, rows, n, dtype, device, name="argument"):
    """Coerce x to a contiguous (rows, n) tensor.  Accepts (rows,), (rows, n), lists of per-row values
    (floats, 1-element tensors -- main_3dof.py:133-137 builds such ragged lists)."""
    if isinstance(x, (list, tuple)):
        x = [v if torch.is_tensor(v) else torch.as_tensor(np.asarray(v, dtype=np.float64)) for v in x]
        x = torch.stack([v.to(device=device, dtype=dtype).reshape(-1).expand(n) for v in x])
    elif not torch.is_tensor(x):
        x = torch.as_tensor(np.asarray(x, dtype=np.float64))
    x = x.to(device=device, dtype=dtype)
    if x.dim() == 1 and x.shape[0] == rows:
        x = x.reshape(rows, 1).expand(rows, n)
    if tuple(x.shape) != (rows, n):
        raise ValueError("%s: expected shape (%d,) or (%d, %d), got %s" % (name, rows, rows, n, tuple(x.shape)))
    return x.contiguous()


def as_vec(x, n, dtype, device):
    if not torch.is_tensor(x):
        x = torch.as_tensor(np.asarray(x, dtype=np.float64))
    return x.to(device=device, dtype=dtype).reshape(-1).expand(n).contiguous()


def is_per_episode(v):
    return (torch.is_tensor(v) and v.numel() > 1) or (isinstance(v, np.ndarray) and v.size > 1)


def scalar(v):
    if torch.is_tensor(v):
        return float(v.reshape(-1)[0])
    return float(np.asarray(v).reshape(-1)[0])


def fill_agent(h, agent):
    """H3AgentHyper from a DHPAgent (attributes read at call time: train.py:120-121 mutates them)."""
    vals = agent._hyper_values()
    for k in L.HYPER_NAMES:
        setattr(h, k, 0.0 if is_per_episode(vals[k]) else scalar(vals[k]))
    h.n_hidden = agent.n_hidden
    h.tracked_state = int(agent.tracked_state)
    h.ac_state = int(agent.ac_states[0])
    h.channel = int(agent.control_channel_num)
    h.final_activation = L.ACTIVATIONS[agent.final_layer]


def default_agent(h, channel):
    h.gamma, h.lr_actor, h.lr_critic, h.tau, h.scale, h.reward_weight = 0.9, 0.1, 0.1, 1.0, 1.0, 1.0
    h.n_hidden, h.tracked_state, h.ac_state, h.channel, h.final_activation = 6, 4 if channel else 1, 5 if channel else 3, channel, 0


def build_config(env=None, pid=None, rls=None, col=None, lon=None, env_dtype=torch.float64, net_dtype=None,
                 switch_step=6000, ref_mode=L.REF_KERNEL, window=(0, 0), store_policy=0, dt=0.01):
    cfg = L.H3Config()
    if env is not None:
        cfg.env = env._params()
        cfg.task = L.TASKS[env.task]
        kp, ki, kd = env.pid_weights if len(env.pid_weights) == 3 else (0.0, 0.0, 0.0)
        cfg.env_kp, cfg.env_ki, cfg.env_kd = float(kp), float(ki), float(kd)
        cfg.env_ref = float(env.ref) if env.ref is not None and not torch.is_tensor(env.ref) else 0.0
        cfg.ref_amp = 10.0 if is_per_episode(env.ref_amplitude) else scalar(env.ref_amplitude)
        cfg.ref_period = 10.0 if is_per_episode(env.ref_period) else scalar(env.ref_period)
        env_dtype = env.dtype
    else:
        cfg.env = L.default_env_params(dt)
        cfg.ref_amp, cfg.ref_period = 10.0, 10.0
    cfg.env_dtype = L.dtype_code(env_dtype)
    if net_dtype is None:
        net_dtype = lon.dtype if lon is not None else (col.dtype if col is not None else env_dtype)
    cfg.net_dtype = L.dtype_code(net_dtype)
    cfg.ref_mode = ref_mode
    if pid is not None:
        cfg.pid_kp, cfg.pid_ki, cfg.pid_kd, cfg.pid_href = float(pid.Kp), float(pid.Ki), float(pid.Kd), float(pid.h_ref)
    else:
        cfg.pid_kp, cfg.pid_ki, cfg.pid_kd, cfg.pid_href = 2.0, 0.2, 0.1, 5.0
    cfg.rls_gamma = float(rls.gamma) if rls is not None else 1.0
    cfg.rls_symmetric = int(bool(rls.assume_symmetric)) if rls is not None else 0
    cfg.store_policy = int(store_policy)
    cfg.switch_step = int(switch_step)
    cfg.update_lon = 1
    cfg.win_lo, cfg.win_hi = int(window[0]), int(window[1])
    for h, agent, chan in ((cfg.col, col, 0), (cfg.lon, lon, 1)):
        if agent is not None:
            fill_agent(h, agent)
        else:
            default_agent(h, chan)
    return cfg


def build_batch(n, n_active=None, **tensors):
    """H3Batch with the given tensors (all (rows, n) or (n,) contiguous on one CUDA device); `n_active` <= n
    launches only the first n_active columns (ld stays n)."""
    b = L.H3Batch()
    b.n = n if n_active is None else n_active
    b.ld = n
    for k, v in tensors.items():
        setattr(b, k, L.ptr(v).value if v is not None else None)
    return b


def byref(x):
    return C.byref(x)
