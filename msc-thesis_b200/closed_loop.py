"""The fused fast path: the tick loop of the reference's main_3dof.py:122-197 for N episodes in lockstep,
n_ticks per kernel launch, every episode's state resident on chip inside a launch.

ClosedLoop3DOF owns nothing but bookkeeping: it borrows the SoA tensors of the five reference-API objects
(Helicopter3DOF, CollectivePID3DOF, RecursiveLeastSquares, 2 x DHPAgent) and launches h3_closed_loop on
them, so the per-object methods stay usable between runs.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib as L
from ._config import build_batch, build_config, byref

TRAJ_COLUMNS = ("x", "z", "u", "w", "theta", "q", "lambda_i", "reference", "collective", "cyclic", "r1", "r2",
                "alive")


def sinusoid_ref_table(t0, dt, n_ticks, amp=10.0, period=10.0):
    """Host-made reference for a lockstep launch (heli_models.py:218-237 evaluated with numpy, so the values
    are the reference's own bit for bit).  Entry j is the reference that becomes current after local tick
    j: get_ref() at t_{j+1} = t0 + (j+1) dt (accumulated like heli_models.py:206), which evaluates
    t_{j+1} + dt.  Returns (table (n_ticks,), t after n_ticks)."""
    t = float(t0)
    out = np.empty(n_ticks, dtype=np.float64)
    for j in range(n_ticks):
        t = t + dt
        x = np.pi * (t + dt) / period
        out[j] = np.deg2rad(amp / 1.76 * (np.sin(x) + np.sin(2 * x)))
    return out, t


class ClosedLoop3DOF:
    def __init__(self, env, pid, rls, agent_col, agent_lon, excitation=None, switch_step=6000,
                 score_window=(0, 0), ticks_per_launch=1000, store_policy=0, ref=None):
        self.env, self.pid, self.rls = env, pid, rls
        self.agent_col, self.agent_lon = agent_col, agent_lon
        self.n = n = env.n_envs
        self.device = env.device
        for obj in (rls, agent_col, agent_lon):
            if obj.n_envs != n or obj.device != env.device:
                raise ValueError("all objects of a closed loop must share n_envs and device")
        if rls.dtype != env.dtype or agent_col.dtype != agent_lon.dtype:
            raise ValueError("env/RLS share one dtype, the two agents share one dtype")
        if (env.dtype, agent_lon.dtype) not in ((torch.float64, torch.float64), (torch.float64, torch.float32),
                                                 (torch.float32, torch.float32)):
            raise ValueError("supported precision pairs (env, nets): (f64,f64), (f64,f32), (f32,f32)")
        self.switch_step = int(switch_step)
        self.score_window = tuple(int(x) for x in score_window)
        self.ticks_per_launch = int(ticks_per_launch)
        self.store_policy = int(store_policy)
        self.excitation = None
        if excitation is not None:
            e = np.asarray(excitation, dtype=np.float64)[:, :2].copy()      # columns [col, lon] of excitation.npy
            self.excitation = torch.as_tensor(e, device=self.device).contiguous()
        pid._ensure(n, env.dtype, self.device)
        # main_3dof.py:62: ref = env.get_ref() before the loop (velocity task: integrates corr_u once)
        if ref is None:
            ref = env.get_ref()
        self.ref2 = ref[[1, 4]].to(env.dtype).contiguous() if ref.shape[0] == 7 else ref.to(env.dtype).contiguous()
        self.acc = torch.zeros(3, n, dtype=torch.float64, device=self.device)
        self.stats_dev = torch.zeros(L.N_STATS, dtype=torch.float64, device=self.device)
        nbytes = int(L.lib().h3_stats_scratch_bytes(n))
        self._scratch = torch.zeros((nbytes + 7) // 8, dtype=torch.float64, device=self.device)
        self.k = 0                  # host copy of the lockstep tick index (None: heterogeneous batch)
        self.launches = 0
        # compaction (long horizons): live episodes are kept in the first n_active columns; `order[i]` is the
        # original index of the episode now stored in column i; retired episodes keep their frozen state in
        # the columns behind n_active and their share of the statistics in _retired
        self.n_active = n
        self.order = torch.arange(n, device=self.device)
        self._retired = torch.zeros(L.N_STATS, dtype=torch.float64, device=self.device)

    # -- state ------------------------------------------------------------------------------------------
    def init_episodes(self, seed, episode0=0, v_trim=None, sigma=None, episode_id=None, covariance=None):
        """(Re-)initialise every episode on the device: what main_3dof.py:58-71 does before its loop -- trim and
        reset of the environment, the first reference, RLS reset, N(0, nn_stdev) network weights with the target
        critic a copy of the critic -- as ONE kernel (h3_init_episodes).  The draws of an episode depend on `seed`
        and its GLOBAL index `episode0 + i` (or `episode_id[i]`) only, so a sharded job makes the same episodes as
        an unsharded one.  v_trim: scalar or (N,) tensor; sigma: None (the agents' nn_stdev) or a (4, N) device
        tensor [col actor, col critic, lon actor, lon critic]."""
        env, rls = self.env, self.rls
        n = self.n
        ini = L.H3Init()
        ini.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        ini.episode0 = int(episode0)
        keep = []
        if v_trim is None:
            v_trim = 10.0
        if torch.is_tensor(v_trim) and v_trim.numel() > 1:
            vt = v_trim.to(device=self.device, dtype=torch.float64).contiguous()
            keep.append(vt)
            ini.v_trim_ep = vt.data_ptr()
        else:
            ini.v_trim = float(v_trim)
        sg = (self.agent_col.nn_stdev_actor, self.agent_col.nn_stdev_critic, self.agent_lon.nn_stdev_actor,
              self.agent_lon.nn_stdev_critic)
        if sigma is not None and torch.is_tensor(sigma) and sigma.dim() == 2:
            sg_t = sigma.to(device=self.device, dtype=torch.float64).contiguous()
            if tuple(sg_t.shape) != (4, n):
                raise ValueError("sigma: expected a (4, N) tensor")
            keep.append(sg_t)
            ini.sigma_ep = sg_t.data_ptr()
            sg = (0.0, 0.0, 0.0, 0.0)
        elif sigma is not None:
            sg = tuple(float(x) for x in sigma)
        for i in range(4):
            ini.sigma[i] = float(sg[i])
        ini.rls_cov0 = float(rls.covariance if covariance is None else covariance)
        if episode_id is not None:
            ids = episode_id.to(device=self.device, dtype=torch.int64).contiguous()
            keep.append(ids)
            ini.episode_id = ids.data_ptr()
        self.n_active = n
        cfg = self._config(L.REF_KERNEL)
        b = self._batch(self.agent_col._hyper_rows(), self.agent_lon._hyper_rows())
        L.check(L.lib().h3_init_episodes(byref(cfg), byref(b), C.byref(ini), L.stream_ptr(self.device)))
        # host-side bookkeeping of the objects whose tensors were just rewritten
        theta = env.trimmed_state[4].clone()
        env.trimmed_state.copy_(env.state)
        env.trimmed_state[4].copy_(theta)
        env.t = 0
        env.uniform_clock = True
        rls.assume_symmetric = True
        rls._sym_dirty = False
        self.agent_col.target_synced = self.agent_lon.target_synced = True
        self.k = 0
        self.order = torch.arange(n, device=self.device)
        self._retired.zero_()
        self._keep_init = keep              # inputs of an asynchronous launch stay alive until the next call
        return self

    def set_heterogeneous(self):
        """Episodes no longer share tick index / clock (e.g. after loading per-episode snapshots): launches use
        each episode's own step counter and evaluate the sinusoid reference in the kernel."""
        self.k = None
        self.env.uniform_clock = False

    def _batch(self, col_hyper, lon_hyper):
        env, rls = self.env, self.rls
        return build_batch(self.n, n_active=self.n_active, state=env.state, t=env.t_ep, ref=self.ref2, corr_u=env.corr_u,
                           trim=env.trimmed_controls, theta_trim=env._theta_trim(), ref_par=env.ref_params,
                           pid_corr=self.pid.hdot_corr, rls_W=rls._W, rls_cov=rls._cov, rls_x=rls.x, rls_u=rls.u,
                           rls_skip=rls._skip, col_net=self.agent_col.net, lon_net=self.agent_lon.net,
                           col_hyper=col_hyper, lon_hyper=lon_hyper, step=env.step_count, flags=env.flags,
                           term_step=env.term_step, acc=self.acc)

    def _config(self, ref_mode):
        return build_config(env=self.env, pid=self.pid, rls=self.rls, col=self.agent_col, lon=self.agent_lon,
                            switch_step=self.switch_step, ref_mode=ref_mode, window=self.score_window,
                            store_policy=self.store_policy)

    def _check_shortcuts(self):
        """The fused kernels take two shortcuts that rest on invariants of the reference's own updates: the upper
        triangle of the covariance stands for the whole matrix (model.py:59 preserves symmetry), and with tau = 1 the
        cyclic agent's target critic is the critic (agents.py:196) and is not carried.  Both hold by construction; they
        are re-verified here (one synchronising comparison) only after something that can break them: a literal
        RecursiveLeastSquares.update() call, or a checkpoint load."""
        rls, lon = self.rls, self.agent_lon
        if rls.assume_symmetric and getattr(rls, "_sym_dirty", False):
            c = rls.cov
            rls.assume_symmetric = bool(torch.equal(c, c.transpose(0, 1)))
            rls._sym_dirty = False
        tau = lon.tau_target_critic
        if lon.target_synced is not True and not torch.is_tensor(tau) and float(tau) == 1.0 and lon.hyper_rows_device is None:
            nh = lon.n_hidden
            lon.target_synced = bool(torch.equal(lon.net[7 * nh:], lon.net[3 * nh:7 * nh]))

    def _uniform_rows(self, agent):
        """(6, N) hyper-parameter rows holding the agent's scalar values (forces the per-episode-hyper kernel variant)."""
        vals = agent._hyper_values()
        rows = torch.empty(L.HYPER_ROWS, self.n, dtype=torch.float64, device=self.device)
        for i, k in enumerate(L.HYPER_NAMES):
            rows[i] = float(vals[k]) if not torch.is_tensor(vals[k]) else vals[k].to(self.device, torch.float64)
        return rows

    def _reference_table(self, n_ticks, min_horizon=16384):
        """Device-resident sinusoid reference for the lockstep clock, built ONCE for a whole horizon (the table is a
        function of the tick index only) and re-used by every launch: run() then issues no host -> device copy.
        Entry i belongs to tick k0 + i; `t[i]` is env.t before that tick, accumulated like heli_models.py:206.  The
        table is rebuilt when the clock does not continue it (reset, a different dt / amplitude / period)."""
        env = self.env
        key = (float(env.dt), float(env.ref_amplitude), float(env.ref_period), str(self.device))
        tab = getattr(self, "_tab", None)
        k, need = self.k, self.k + n_ticks
        ok = (tab is not None and tab["key"] == key and tab["k0"] <= k and need <= tab["k0"] + len(tab["host"])
              and float(tab["t"][k - tab["k0"]]) == float(env.t))
        if not ok:
            if tab is not None and tab["key"] == key and tab["k0"] <= k <= tab["k0"] + len(tab["host"]) and \
                    float(tab["t"][k - tab["k0"]]) == float(env.t):
                k0, t0 = tab["k0"], float(tab["t"][0])          # same clock: extend the horizon
            else:
                k0, t0 = k, float(env.t)
            n = max(need - k0, min_horizon, 2 * len(tab["host"]) if tab is not None and k0 == tab["k0"] else 0)   # grow geometrically
            host, _ = sinusoid_ref_table(t0, env.dt, n, env.ref_amplitude, env.ref_period)
            t = np.empty(n + 1)
            acc = t0
            for i in range(n + 1):
                t[i] = acc
                acc = acc + env.dt
            pinned = torch.from_numpy(host).pin_memory()
            tab = {"key": key, "k0": k0, "host": host, "t": t, "pinned": pinned,
                   "dev": pinned.to(self.device, non_blocking=True)}
            self._tab = tab
        return tab["dev"], tab["t"]

    def plan(self, n_ticks):
        """Launch plan [(first tick, ticks)] : split at the phase switch so that phase-1 launches never load
        the collective agent's networks, and at ticks_per_launch."""
        k0 = self.k if self.k is not None else 0
        cuts = [0, n_ticks]
        if self.k is not None and k0 < self.switch_step < k0 + n_ticks:
            cuts.append(self.switch_step - k0)
        cuts = sorted(set(cuts))
        out = []
        for a, b in zip(cuts[:-1], cuts[1:]):
            j = a
            while j < b:
                m = min(self.ticks_per_launch, b - j)
                out.append((j, m))
                j += m
        return out

    def run(self, n_ticks, record=0, want_stats=True):
        """Advance every live episode by n_ticks ticks.  record = R > 0 returns the per-tick log rows of
        main_3dof.py:164-175 for the first R episodes as a (n_ticks, 13, R) tensor (TRAJ_COLUMNS)."""
        env = self.env
        n_ticks = int(n_ticks)
        lockstep = self.k is not None and env.uniform_clock
        per_ep_ref = env.ref_params is not None
        table_mode = lockstep and env.task == "sinusoid" and not per_ep_ref
        self._check_shortcuts()
        cfg = self._config(L.REF_TABLE if table_mode else L.REF_KERNEL)
        col_hyper, lon_hyper = self.agent_col._hyper_rows(), self.agent_lon._hyper_rows()
        if lon_hyper is None and self.agent_lon.target_synced is False:
            lon_hyper = self._uniform_rows(self.agent_lon)      # carries the target critic explicitly (no tau = 1 shortcut)
        b = self._batch(col_hyper, lon_hyper)
        traj = None
        self._t_at_record_start = env.t
        if record:
            traj = torch.zeros(n_ticks, L.TRAJ_ROWS, record, dtype=env.dtype, device=self.device)
        stream = L.stream_ptr(self.device)
        lib = L.lib()
        plan = self.plan(n_ticks)
        tab_dev, tab_t = (None, None)
        if table_mode:
            tab_dev, tab_t = self._reference_table(n_ticks)
        for j0, m in plan:
            table = None
            if table_mode:
                # a slice of the device-resident table of the whole horizon: no host -> device copy, no sync
                off = self._tab["k0"]
                table = tab_dev[self.k + j0 - off:self.k + j0 - off + m]
                t_end = float(tab_t[self.k + j0 - off + m])
            else:
                t_end = env.t
                for _ in range(m):
                    t_end = t_end + env.dt
            tr = None
            if traj is not None:
                tr = L.H3Traj()
                tr.out = traj[j0:j0 + m].data_ptr()
                tr.n_traj = record
                tr.ld_traj = record
            last = (j0 + m == n_ticks)
            stats = self.stats_dev if (want_stats and last) else None
            L.check(lib.h3_closed_loop(byref(cfg), byref(b), self.k + j0 if lockstep else -1, L.ptr(table),
                                       L.ptr(self.excitation),
                                       0 if self.excitation is None else self.excitation.shape[0], m,
                                       byref(tr) if tr is not None else None, L.ptr(stats),
                                       L.ptr(self._scratch) if stats is not None else None, stream))
            env.t = t_end
            self.launches += 1
        if want_stats and self.n_active < self.n:
            self.stats_dev += self._retired         # episodes retired by compact() are not launched any more
        if self.k is not None:
            self.k += n_ticks
        return traj

    def run_logged(self, n_ticks, every=10, episodes=1, as_reference=False):
        """run() with BOTH logs of main_3dof.py: the per-tick rows (:164-175, returned like run(record=episodes)) and the
        network / incremental-model history that the script appends after the ticks 0, every, 2 every, ... (:178-186):
        for each agent the six weight matrices (actor, critic, target critic; input and output layer, raveled in
        torch's (out, in) order) and F, G from get_transition_matrices.  Returns (traj, weight_stats) with
        weight_stats = {'t': (n_ticks,), 'col': {'nn': {'a': {'i': (T, 2 nh, R), 'o': (T, nh, R)}, 'c': ..., 'tc': ...},
        'rls': {'F': (T, 4, R), 'G': (T, 2, R)}}, 'lon': {...}} as CPU tensors (R = episodes, T = logged ticks).
        as_reference: main_3dof.py:181 walks `agent_col.networks` for BOTH agents, so its 'lon' entries hold the
        collective agent's weights; True reproduces that, False (default) logs each agent's own networks.
        The history is taken between launches (no extra kernel): launches are cut at the logged ticks."""
        n_ticks, every, R = int(n_ticks), int(every), int(episodes)
        k_start = self.k if self.k is not None else 0
        t_start = float(self.env.t)
        agents = {"col": self.agent_col, "lon": self.agent_lon}
        hist = {a: {"nn": {n: {"i": [], "o": []} for n in ("a", "c", "tc")}, "rls": {"F": [], "G": []}} for a in agents}
        parts, done = [], 0

        def snapshot():
            for name, ag in agents.items():
                src = self.agent_col if as_reference else ag
                for key, net in zip(("a", "c", "tc"), src.networks):
                    hist[name]["nn"][key]["i"].append(net.fc1.weight[..., :R].reshape(-1, R).clone())
                    hist[name]["nn"][key]["o"].append(net.fc2.weight[..., :R].reshape(-1, R).clone())
                F, G = ag.get_transition_matrices(self.rls)
                hist[name]["rls"]["F"].append(F[..., :R].reshape(4, R).clone())
                hist[name]["rls"]["G"].append(G[..., :R].reshape(2, R).clone())

        while done < n_ticks:
            k = k_start + done
            m = 1 if k % every == 0 else min(every - k % every, n_ticks - done)     # stop right after a logged tick
            parts.append(self.run(m, record=R, want_stats=(done + m == n_ticks)))
            if k % every == 0:
                snapshot()
            done += m
        traj = torch.cat(parts) if parts else None
        t = np.empty(n_ticks)
        acc = t_start
        for i in range(n_ticks):                    # env.t after every tick, accumulated like heli_models.py:206
            acc = acc + self.env.dt
            t[i] = acc
        out = {"t": torch.as_tensor(t)}
        for a in agents:
            out[a] = {"nn": {n: {l: torch.stack(v).cpu() for l, v in d.items()} for n, d in hist[a]["nn"].items()},
                      "rls": {k2: torch.stack(v).cpu() for k2, v in hist[a]["rls"].items()}}
        self._t_at_record_start = t_start
        return traj, out

    def _episode_tensors(self):
        env, rls = self.env, self.rls
        ts = [env.state, env.t_ep, self.ref2, env.corr_u, env.trimmed_controls, env.trimmed_state, self.pid.hdot_corr,
              rls._W, rls._cov, rls.x, rls.u, rls._skip, rls.X, self.agent_col.net, self.agent_lon.net,
              env.step_count, env.flags, env.term_step, self.acc, self.order]
        if env.ref_params is not None:
            ts.append(env.ref_params)
        for ag in (self.agent_col, self.agent_lon):
            if ag.hyper_rows_device is not None:
                ts.append(ag.hyper_rows_device)
        return ts

    def compact(self, sort_key=None):
        """Move the live episodes to the front so that launches only cover them: terminated episodes otherwise keep
        occupying lanes of their warp.  Returns the number of live episodes.  Per-episode tensors of all five objects
        are permuted IN PLACE; `self.order` maps columns to original episode indices (`self.original_order(t)` undoes
        the permutation for reporting).  Per-episode hyper-parameter tensors held by the agents are permuted as well.

        sort_key: optional (n_active,) tensor; the live episodes are additionally ordered by it (ascending).  Episodes
        are independent, so any order gives the same per-episode results bit for bit; grouping episodes that take the
        same data-dependent branches (tanh / sincos range) into the same warps removes the divergence between them."""
        na = self.n_active
        alive = self.env.flags[:na] == 0
        n_alive = int(alive.sum().item())
        if n_alive == na and sort_key is None:
            return na
        if sort_key is None:
            perm = torch.cat([torch.nonzero(alive).reshape(-1), torch.nonzero(~alive).reshape(-1)])
        else:
            key = sort_key[:na].to(torch.float64)
            key = torch.where(alive, torch.nan_to_num(key, nan=1e300, posinf=1e300, neginf=-1e300),
                              torch.full_like(key, float("inf")))
            perm = torch.argsort(key, stable=True)
        for t in self._episode_tensors():
            head = t[..., :na]
            head.copy_(head.index_select(-1, perm))
        for ag in (self.agent_col, self.agent_lon):
            for name in ("gamma", "learning_rate_actor", "learning_rate_critic", "tau_target_critic", "reward_weight"):
                v = getattr(ag, name)
                if torch.is_tensor(v) and v.numel() == self.n:
                    v = v.clone()
                    v[:na] = v[:na].index_select(0, perm.to(v.device))
                    setattr(ag, name, v)
            if torch.is_tensor(ag.actor.scale) and ag.actor.scale.numel() == self.n:
                v = ag.actor.scale.clone()
                v[:na] = v[:na].index_select(0, perm.to(v.device))
                ag.actor.scale = v
        # statistics share of the episodes retired now (same field order as the kernel epilogue)
        f = self.env.flags[n_alive:na]
        r = torch.zeros(L.N_STATS, dtype=torch.float64, device=self.device)
        r[0] = na - n_alive
        r[2] = ((f & L.FLAG_DONE) != 0).sum()
        r[3] = ((f & L.FLAG_NAN) != 0).sum()
        r[4:7] = self.acc[:, n_alive:na].sum(dim=1)
        r[7] = self.env.step_count[n_alive:na].sum()
        self._retired += r
        self.n_active = n_alive
        return n_alive

    def sort_key(self):
        """Per-episode key for compact(sort_key=...): the current pitch-loop agitation |theta - theta_ref| + 0.1 |q|.
        The tick has no data-dependent branch (tanh and sin / cos are one formula for every argument), so the order no
        longer decides which lanes diverge; it still matters because agitated episodes are the ones that terminate next:
        sorted, the lanes that die inside the next launch sit in the same warps (2-3 % over the launches)."""
        st = self.env.state
        return (st[4] - self.ref2[1]).abs().double() + 0.1 * st[5].abs().double()

    def original_order(self, t):
        """Per-episode tensor `t` (episode axis last) re-ordered to the original episode indices."""
        out = torch.empty_like(t)
        out.index_copy_(-1, self.order.to(t.device), t)
        return out

    # -- results ----------------------------------------------------------------------------------------------
    def stats(self):
        """This rank's batch statistics of the last run() as a dict (synchronises)."""
        v = self.stats_dev.cpu().tolist()
        return dict(zip(L.STATS_FIELDS, v))

    def allreduce_stats(self, group=None):
        """Sum the statistics vector over all ranks (NCCL all-reduce on the compute stream; gloo on CPU)."""
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.stats_dev, op=dist.ReduceOp.SUM, group=group)
        return self.stats_dev

    def to_dataframe(self, traj, episode=0, t0=None):
        """The per-tick log of main_3dof.py:164-175 for one recorded episode as a pandas DataFrame with the
        reference's column names (t, x, z, u, w, theta, q, reference, collective, cyclic, r1, r2): what
        plotting.plot_stats_3dof consumes.  `t` is env.t AFTER the step, accumulated like heli_models.py:206."""
        import pandas as pd
        rows = traj[:, :, episode].detach().cpu().double().numpy()
        rows = rows[rows[:, 12] > 0]                     # ticks this episode actually executed
        t = np.empty(len(rows))
        acc = float(self._t_at_record_start if t0 is None else t0)
        for i in range(len(rows)):
            acc = acc + self.env.dt
            t[i] = acc
        names = ("x", "z", "u", "w", "theta", "q", "lambda_i", "reference", "collective", "cyclic", "r1", "r2")
        df = pd.DataFrame({n: rows[:, i] for i, n in enumerate(names) if n != "lambda_i"})
        df.insert(0, "t", t)
        return df

    def info(self, phase2=True):
        """Occupancy facts of the kernel variant run() would launch (block, smem, blocks/SM, registers)."""
        cfg = self._config(L.REF_TABLE)
        if self.agent_col._hyper_rows() is not None or self.agent_lon._hyper_rows() is not None:
            cfg.lon.tau = 0.5           # the query has no batch: any tau != 1 selects the per-episode-hyper variant
        vals = [C.c_int32() for _ in range(5)]
        L.check(L.lib().h3_closed_loop_info(byref(cfg), 1 if phase2 else 0, *[C.byref(v) for v in vals]))
        return dict(zip(("block", "smem_bytes", "blocks_per_sm", "regs", "lanes_per_warp"), (v.value for v in vals)))
