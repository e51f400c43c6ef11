"""Whole episodes from per-episode parameters: the batched form of running the reference's script once per episode.

One `EpisodeRunner.run()` call is N runs of main_3dof.py (/root/reference/main_3dof.py:58-203): the host supplies, per
episode, what distinguishes one run of the script from another -- trim speed, reference-signal amplitude and period,
initial-weight standard deviations, the agents' hyper-parameters -- plus one seed for the job; the device does the
rest: h3_init_episodes (trim, reset, weight draws), n_ticks closed-loop ticks with the phase switch, and hands back
per-episode returns, score accumulator, termination flags / ticks and the batch statistics.  Per episode 152 bytes go
up and 36 bytes come back; the 2.7 KB of mutable episode state never leaves the GPU."""
import torch

from . import _lib as L
from .PID import CollectivePID3DOF
from .agents import DHPAgent
from .closed_loop import ClosedLoop3DOF
from .excitation import excitation_signal
from .heli_models import Helicopter3DOF
from .model import RecursiveLeastSquares

# rows of the (19, N) float64 parameter block
PARAM_ROWS = ("v_trim", "ref_amplitude", "ref_period", "sigma_col_actor", "sigma_col_critic", "sigma_lon_actor",
              "sigma_lon_critic") + tuple("col_" + k for k in L.HYPER_NAMES) + tuple("lon_" + k for k in L.HYPER_NAMES)
N_PARAMS = len(PARAM_ROWS)
ROW = {k: i for i, k in enumerate(PARAM_ROWS)}

# main_3dof.py:20-46 (scale = deg2rad(action_scaling), agents.py:46-47)
DEFAULTS = {"v_trim": 10.0, "ref_amplitude": 10.0, "ref_period": 10.0, "sigma_col_actor": 0.1, "sigma_col_critic": 0.1,
            "sigma_lon_actor": 0.2, "sigma_lon_critic": 0.2, "col_gamma": 0.9, "col_lr_actor": 0.1, "col_lr_critic": 0.1,
            "col_tau": 0.01, "col_scale": 0.08726646259971647, "col_reward_weight": 0.01, "lon_gamma": 0.99,
            "lon_lr_actor": 0.2, "lon_lr_critic": 0.2, "lon_tau": 1.0, "lon_scale": 0.2617993877991494,
            "lon_reward_weight": 1.0}

COL_KW = dict(control_channel="col", discount_factor=0.9, n_hidden_actor=10, nn_stdev_actor=0.1, learning_rate_actor=0.1,
              action_scaling=5, action_network_final_layer="tanh", n_hidden_critic=10, nn_stdev_critic=0.1,
              learning_rate_critic=0.1, tau_target_critic=0.01, tracked_state=1, ac_states=[3], reward_weight=0.01)
LON_KW = dict(control_channel="lon", discount_factor=0.99, n_hidden_actor=6, nn_stdev_actor=0.2, learning_rate_actor=0.2,
              action_scaling=15, action_network_final_layer="tanh", n_hidden_critic=6, nn_stdev_critic=0.2,
              learning_rate_critic=0.2, tau_target_critic=1, tracked_state=4, ac_states=[5])


def default_params(n, pin=False, **overrides):
    """(19, N) float64 host block filled with the main_3dof.py values; overrides: row name -> scalar or (N,) values."""
    p = torch.empty(N_PARAMS, n, dtype=torch.float64, pin_memory=pin)
    for k, i in ROW.items():
        p[i] = DEFAULTS[k]
    for k, v in overrides.items():
        p[ROW[k]] = torch.as_tensor(v, dtype=torch.float64)
    return p


def make_loop(n, device, env_dtype=torch.float64, net_dtype=torch.float64, excitation_ticks=1000, switch_step=6000,
              score_window=(5000, 6000), ticks_per_launch=1000, col_kw=None, lon_kw=None):
    """The five objects of main_3dof.py:58-71 for N episodes with NO host-side initialisation work (weights are drawn
    by ClosedLoop3DOF.init_episodes), wrapped in a ClosedLoop3DOF."""
    env = Helicopter3DOF(dt=0.01, t_max=120, n_envs=n, device=device, dtype=env_dtype)
    env.task, env.pid_weights, env.ref = "sinusoid", (0, 0, 0), 20          # config_3dof.json:34
    rls = RecursiveLeastSquares(state_size=7, action_size=2, gamma=1, covariance=10 ** 4, constant=False, n_envs=n,
                                device=device, dtype=env_dtype)
    pid = CollectivePID3DOF(h_ref=5, dt=0.01)
    col = DHPAgent(**(col_kw or COL_KW), n_envs=n, device=device, dtype=net_dtype, init="device")
    lon = DHPAgent(**(lon_kw or LON_KW), n_envs=n, device=device, dtype=net_dtype, init="device")
    ref0 = torch.zeros(2, n, dtype=env_dtype, device=device)                # overwritten by init_episodes
    return ClosedLoop3DOF(env, pid, rls, col, lon, excitation=excitation_signal(excitation_ticks) if excitation_ticks else None,
                          switch_step=switch_step, score_window=score_window, ticks_per_launch=ticks_per_launch, ref=ref0)


class EpisodeRunner:
    """Device-side state for N whole episodes, reused from call to call."""

    def __init__(self, n, device=None, precision="f64", per_episode_hyper=True, **loop_kw):
        edt = torch.float32 if precision == "f32" else torch.float64
        ndt = torch.float64 if precision == "f64" else torch.float32
        self.n = int(n)
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.loop = make_loop(self.n, self.device, edt, ndt, **loop_kw)
        self.per_episode_hyper = bool(per_episode_hyper)
        self.params = torch.empty(N_PARAMS, self.n, dtype=torch.float64, device=self.device)
        self.res_f = torch.empty(3, self.n, dtype=torch.float64, device=self.device)     # views set in run()
        self.res_i = torch.empty(3, self.n, dtype=torch.int32, device=self.device)
        lp = self.loop
        lp.env.ref_params = self.params[ROW["ref_amplitude"]:ROW["ref_period"] + 1]
        if self.per_episode_hyper:
            lp.agent_col.hyper_rows_device = self.params[ROW["col_gamma"]:ROW["col_gamma"] + L.HYPER_ROWS]
            lp.agent_lon.hyper_rows_device = self.params[ROW["lon_gamma"]:ROW["lon_gamma"] + L.HYPER_ROWS]
        else:
            lp.env.ref_params = None        # launch-uniform reference: the host table path (H3_REF_TABLE)

    def result_buffers(self, pin=True):
        """Host buffers for run(): ((3, N) f64 [return col, return lon, score-window sum], (3, N) i32 [flags, term_step,
        step], (8,) f64 statistics)."""
        return (torch.empty(3, self.n, dtype=torch.float64, pin_memory=pin),
                torch.empty(3, self.n, dtype=torch.int32, pin_memory=pin),
                torch.empty(L.N_STATS, dtype=torch.float64, pin_memory=pin))

    def run(self, params_host, results_host, seed, n_ticks=12000, episode0=0, compact_every=0, sort=True):
        """One job: parameters up (H2D), init, n_ticks ticks, results down (D2H) -- all enqueued on the current
        stream; synchronise before reading results_host.  compact_every = K > 0 moves the live episodes to the front
        every K ticks (ClosedLoop3DOF.compact(); it synchronises), so terminated episodes stop occupying lanes; with
        `sort` they are also ordered by ClosedLoop3DOF.sort_key(), which takes the divergence between calm and
        agitated episodes out of the warps.  Results do not depend on either."""
        lp = self.loop
        self.params.copy_(params_host, non_blocking=True)
        lp.init_episodes(seed, episode0=episode0, v_trim=self.params[ROW["v_trim"]],
                         sigma=self.params[ROW["sigma_col_actor"]:ROW["sigma_lon_critic"] + 1])
        if compact_every and compact_every > 0:
            done = 0
            while done < n_ticks:
                m = min(compact_every, n_ticks - done)
                lp.run(m)
                done += m
                if done < n_ticks:
                    lp.compact(sort_key=lp.sort_key() if sort else None)
        else:
            lp.run(n_ticks)
        out_f, out_i, out_s = results_host
        if lp.n_active < lp.n or compact_every:
            acc, fl, ts, sc = (lp.original_order(t) for t in (lp.acc, lp.env.flags, lp.env.term_step, lp.env.step_count))
        else:
            acc, fl, ts, sc = lp.acc, lp.env.flags, lp.env.term_step, lp.env.step_count
        out_f.copy_(acc, non_blocking=True)
        out_i[0].copy_(fl, non_blocking=True)
        out_i[1].copy_(ts, non_blocking=True)
        out_i[2].copy_(sc, non_blocking=True)
        out_s.copy_(lp.stats_dev, non_blocking=True)
        return results_host

    @staticmethod
    def h2d_bytes(n):
        return N_PARAMS * 8 * n

    @staticmethod
    def d2h_bytes(n):
        return (3 * 8 + 3 * 4) * n + L.N_STATS * 8
