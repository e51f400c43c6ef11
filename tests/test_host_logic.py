"""CPU: host-side logic and the C-ABI boundary (no GPU compute): the library loads and exports every symbol
include/heli3dof.h declares, the ctypes mirrors match the C structs, host-made tables are the reference's
values, object plumbing / checkpoints work on CPU tensors, compute calls fail loudly without a GPU."""
import ctypes
import json
import os
import re
import subprocess

import numpy as np
import pytest
import torch

import msc_thesis_b200 as h3
from msc_thesis_b200 import _lib as L
from tests import helpers as H

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "heli3dof.h")


def test_library_exports_every_declared_symbol():
    text = open(HEADER).read()
    declared = set(re.findall(r"\b(h3_[a-z0-9_]+)\s*\(", text))
    assert declared == set(L.EXPORTS), declared ^ set(L.EXPORTS)
    lib = h3.lib()
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.h3_version() == 100
    assert lib.h3_stats_scratch_bytes(64) == 16 + 8 * 8 * 8


def test_ctypes_structs_match_the_header(tmp_path):
    src = tmp_path / "layout.c"
    src.write_text('''
#include <stdio.h>
#include <stddef.h>
#include "heli3dof.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu\\n", sizeof(H3EnvParams), sizeof(H3AgentHyper), sizeof(H3Config), sizeof(H3Batch), sizeof(H3Traj), sizeof(H3Init));
  printf("%zu %zu %zu %zu %zu\\n", offsetof(H3Config, env_dtype), offsetof(H3Config, rls_gamma), offsetof(H3Config, switch_step), offsetof(H3Config, col), offsetof(H3Config, lon));
  printf("%zu %zu %zu %zu %zu %zu %zu\\n", offsetof(H3Batch, state), offsetof(H3Batch, rls_skip), offsetof(H3Batch, step), offsetof(H3Batch, acc), offsetof(H3Batch, heat), offsetof(H3Init, sigma), offsetof(H3Init, episode_id));
  return 0;
}''')
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    lines = subprocess.check_output([str(exe)]).decode().split("\n")
    sizes = [int(x) for x in lines[0].split()]
    assert sizes == [ctypes.sizeof(c) for c in (L.H3EnvParams, L.H3AgentHyper, L.H3Config, L.H3Batch, L.H3Traj, L.H3Init)]
    offs = [int(x) for x in lines[1].split()]
    assert offs == [getattr(L.H3Config, k).offset for k in ("env_dtype", "rls_gamma", "switch_step", "col", "lon")]
    offs = [int(x) for x in lines[2].split()]
    assert offs == [getattr(L.H3Batch, k).offset for k in ("state", "rls_skip", "step", "acc", "heat")] + [
        L.H3Init.sigma.offset, L.H3Init.episode_id.offset]


def test_6dof_exports_and_struct_layout(tmp_path):
    """include/heli6dof.h: every declared h6_* symbol is exported, and the ctypes mirrors match the C structs."""
    text = open(os.path.join(ROOT, "include", "heli6dof.h")).read()
    declared = set(re.findall(r"\b(h6_[a-z0-9_]+)\s*\(", text))
    assert declared == set(L.H6_EXPORTS), declared ^ set(L.H6_EXPORTS)
    lib = h3.lib()
    for name in declared:
        assert getattr(lib, name) is not None
    out = (ctypes.c_double * 3)()
    lib.h6_engine_power(out)
    assert list(out) == [2 * 313000 * 0.95 * 0.8643, 0.95 * 313000 * 0.95, 1.10 * 313000 * 0.95]    # heli_models.py:397-401
    assert lib.h6_stats_scratch_bytes(1000) == 4 * 10 * 8
    src = tmp_path / "layout6.c"
    src.write_text('''
#include <stdio.h>
#include <stddef.h>
#include "heli6dof.h"
int main(void) {
  printf("%zu %zu %zu\\n", sizeof(H6Config), sizeof(H6Batch), sizeof(H6Traj));
  printf("%zu %zu %zu %zu %zu %zu\\n", offsetof(H6Config, lat_gains), offsetof(H6Config, amp), offsetof(H6Config, run_mode), offsetof(H6Config, win_hi), offsetof(H6Config, col), offsetof(H6Config, lon));
  printf("%zu %zu %zu %zu %zu %zu\\n", offsetof(H6Batch, state), offsetof(H6Batch, trim_act), offsetof(H6Batch, rls_skip), offsetof(H6Batch, acc), offsetof(H6Batch, tile), offsetof(H6Batch, rls_f32));
  return 0;
}''')
    exe = tmp_path / "layout6"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    lines = subprocess.check_output([str(exe)]).decode().split("\n")
    assert [int(x) for x in lines[0].split()] == [ctypes.sizeof(c) for c in (L.H6Config, L.H6Batch, L.H6Traj)]
    assert [int(x) for x in lines[1].split()] == [getattr(L.H6Config, k).offset for k in
                                                  ("lat_gains", "amp", "run_mode", "win_hi", "col", "lon")]
    assert [int(x) for x in lines[2].split()] == [getattr(L.H6Batch, k).offset for k in
                                                  ("state", "trim_act", "rls_skip", "acc", "tile", "rls_f32")]
    assert len(L.H6_BATCH_FIELDS) == 25 and L.H6Batch.tile.offset == L.H6Batch.work.offset + 24


def test_6dof_abi_rejects_bad_arguments_before_touching_the_gpu():
    """Error convention of include/heli6dof.h: negative status + h3_last_error() text, never an exception or a launch,
    for null buffers, inconsistent sizes and unsupported configurations."""
    lib = h3.lib()
    cfg, b = L.H6Config(), L.H6Batch()
    b.n, b.ld = 8, 8
    assert lib.h6_closed_loop(ctypes.byref(cfg), ctypes.byref(b), None, 0, 10, None, None, None, None) == -1
    assert b"null field" in lib.h3_last_error()
    assert lib.h6_closed_loop(None, ctypes.byref(b), None, 0, 10, None, None, None, None) == -1
    assert lib.h6_trim(None, None, None, 1.0, None, None, None, None, 4, 4, None) == -1
    assert lib.h6_env_derivatives(None, None, None, None, 0, 0.01, None, 4, 4, None) == -1
    assert lib.h6_env_step(ctypes.byref(cfg), ctypes.byref(b), None, None, 0, 5, None, 0, 0, None) == -1
    assert lib.h6_rls_update(ctypes.byref(cfg), ctypes.byref(b), None, None, None, None) == -1
    assert lib.h6_tile_pack(None, None, 4, 8, 8, 1, None) == -1
    fake = ctypes.c_void_p(256)                     # non-null placeholder: validation must stop before any access
    for k in L.H6_BATCH_FIELDS:
        if k not in ("work",):
            setattr(b, k, fake)
    b.ld = 4                                        # ld < n
    assert lib.h6_closed_loop(ctypes.byref(cfg), ctypes.byref(b), None, 0, 10, None, None, None, None) == -1
    b.ld, b.tile = 8, 16                            # only 0 or 32
    assert lib.h6_closed_loop(ctypes.byref(cfg), ctypes.byref(b), None, 0, 10, None, None, None, None) == -1
    b.tile = 0
    cfg.col.n_hidden, cfg.lon.n_hidden = 6, 10      # compiled for 10 / 10
    assert lib.h6_closed_loop(ctypes.byref(cfg), ctypes.byref(b), None, 0, 10, None, None, None, None) == -2
    assert b"n_hidden" in lib.h3_last_error()
    cfg.col.n_hidden = cfg.lon.n_hidden = 10
    cfg.col.tracked_state = 15                      # outside the 15-state observation
    assert lib.h6_closed_loop(ctypes.byref(cfg), ctypes.byref(b), None, 0, 10, None, None, None, None) == -2
    # n == 0 / n_ticks == 0: nothing to do, no launch
    cfg.col.tracked_state, b.n = 11, 0
    assert lib.h6_closed_loop(ctypes.byref(cfg), ctypes.byref(b), None, 0, 10, None, None, None, None) == 0


def test_6dof_host_objects_fail_loudly_without_gpu():
    """The 6-DOF classes can be constructed on the CPU (plumbing), every compute call raises: no fallback."""
    import importlib
    H6 = importlib.import_module("msc-thesis_b200.heli6dof")
    env = H6.Helicopter6DOF(dt=0.01, t_max=120, n_envs=4, device="cpu")
    assert env.state.shape == (15, 4) and env.Paeo == 2 * 313000 * 0.95 * 0.8643
    env.set_engine_status(1, transient=True)
    assert float(env.P_available[0]) == env.Poeitr
    with pytest.raises(ValueError):
        env.set_engine_status(3)
    with pytest.raises(L.H3Error):
        env.trim(15, 0, 0)
    pid = H6.LatPedPID(0.0, 0.5, 0.5, 0.01, H6.pid_params)
    assert pid.gains[4] == pid.gains[3] == np.rad2deg(5)              # PID.py:51 uses Kphi twice
    gen = H6.RefGenerator(T=10, dt=0.01, A=10, u_ref=0, t_switch=60, filter_tau=2.5, n_envs=4, device="cpu")
    with pytest.raises(NotImplementedError):
        gen.get_ref(env.state, 0.0)                                   # util.py:163: no task set
    with pytest.raises(NotImplementedError):
        H6.RecursiveLeastSquares6(state_size=15, action_size=2, constant=True, device="cpu")
    rls = H6.RecursiveLeastSquares6(**H6.rls_params, n_envs=4, device="cpu")
    assert rls.W.shape == (17, 15, 4) and float(rls.cov[3, 3, 0]) == 1e8 and rls.skip_update
    assert H6.env_params_train["step_switch"] == 6000 and H6.env_params_test2["n_steps"] == 3000
    rls.W[3, 2, 1] = 0.25
    rls.cov[1, 4, 2] = 7.0
    rls.skip_update = False
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        rls.save(os.path.join(d, "rls.pkl"))                          # train.py:238
        back = H6.RecursiveLeastSquares6.load(os.path.join(d, "rls.pkl"), device="cpu")
    assert torch.equal(back.W, rls.W) and torch.equal(back.cov, rls.cov) and not back.skip_update
    assert not back.assume_symmetric                                  # the loaded covariance is not symmetric bit for bit
    assert len(list(__import__("itertools").product(*H6.SWEEP_GRID.values()))) == 3 * 3 * 5 * 3 * 3   # main.py:107-112
    lim, which = H6.envelope_limits_reached(torch.tensor([[0.0] * 6 + [0.0, 1.6] + [0.0] * 7, [0.0] * 6 + [-1.56, 0.0] + [0.0] * 7,
                                                          [0.0] * 15]).T.contiguous())
    assert lim.tolist() == [True, True, False] and which.tolist() == [1, 2, 0]


def test_6dof_logger_collects_rows_like_util_logger():
    """heli6dof.Logger.log_states / finalize (util.py:30-101) on CPU tensors: one row per call for the chosen episode."""
    import importlib
    H6 = importlib.import_module("msc-thesis_b200.heli6dof")
    lg = H6.Logger(params={"x": 1}, episode=1)
    n = 3
    obs = torch.arange(15 * n, dtype=torch.float64).reshape(15, n)
    ref = torch.full((15, n), float("nan"), dtype=torch.float64)
    ref[7] = 0.1
    act = torch.tensor([[0.5] * n, [0.6] * n, [0.4] * n, [0.3] * n], dtype=torch.float64)
    rew = torch.tensor([[0.0] * n, [-1.0, -2.0, -3.0]], dtype=torch.float64)
    for k in range(2):
        lg.log_states(0.01 * (k + 1), obs + k, ref, act, rew, torch.full((n,), 5e5, dtype=torch.float64),
                      torch.tensor([6e5, 4e5, 3e5], dtype=torch.float64))
    lg.finalize()
    df = lg.state_history
    assert list(df.columns) == ["t", "u", "v", "w", "p", "q", "r", "phi", "theta", "psi", "x", "y", "z", "Pav", "Peng", "omega",
                                "ref", "col", "lon", "lat", "ped", "r1", "r2"]
    assert len(df) == 2 and df["u"].tolist() == [1.0, 2.0] and df["omega"].tolist() == [43.0, 44.0]
    assert df["r2"].tolist() == [-2.0, -2.0] and df["Peng"].tolist() == [4e5, 4e5] and df["t"].tolist() == [0.01, 0.02]
    assert np.isnan(df["ref"].iloc[0][0]) and df["ref"].iloc[0][7] == 0.1
    with pytest.raises(ValueError):
        lg.finalize()


def test_default_env_params_are_the_reference_constants():
    p = L.default_env_params(0.01)
    assert (p.g, p.cl_alpha, p.blade_solidity, p.lock_number, p.cds, p.mass) == (9.81, 5.7, 0.075, 6, 1.5, 2200)
    assert (p.rho, p.v_tip, p.rotor_radius, p.iy, p.mast, p.tau, p.dt) == (1.225, 200, 7.32, 10615, 1, 0.1, 0.01)
    assert p.omega == 200 / 7.32 and p.area == np.pi * 7.32 ** 2


def test_reference_table_is_the_references_own_signal():
    """Entry j of the host-made table is get_ref() after tick j; the golden `ref` column holds get_ref() before
    tick k (main_3dof.py:62, :142) -> table[j] == golden[j + 1], bit for bit."""
    g = H.load("seed2_all64.npz")
    tab, t_end = h3.sinusoid_ref_table(0, 0.01, 3000)
    assert np.array_equal(tab[:2999], g["ref"][1:3000, 1])
    assert t_end == g["t"][2999]
    # chained launches continue the same clock
    tab2, _ = h3.sinusoid_ref_table(t_end, 0.01, 100)
    assert np.array_equal(tab2, g["ref"][3001:3101, 1])


def test_excitation_closed_form():
    assert np.abs(h3.excitation_signal(1000) - H.excitation()).max() < 1e-17
    assert h3.excitation_signal(1000).shape == (1000, 2)


def test_shard_range_partitions_the_episode_index():
    for n, w in ((8388608, 8), (10, 3), (5, 8), (1, 1)):
        spans = [h3.shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans[:-1], spans[1:]))
        assert max(hi - lo for lo, hi in spans) - min(hi - lo for lo, hi in spans) <= 1
    with pytest.raises(ValueError):
        h3.shard_range(10, 3, 3)


def test_agent_constructor_reproduces_reference_weights():
    """torch.manual_seed(2) + the constructor order of main_3dof.py:69-70 gives the reference's initial weights
    (the golden ref32 snapshot), because init='reference' consumes the global RNG exactly like agents.py."""
    g = H.load("seed2_ref32.npz")
    torch.manual_seed(2)
    kw = dict(action_network_final_layer='tanh', device="cpu", dtype=torch.float32)
    col = h3.DHPAgent(control_channel='col', discount_factor=0.9, n_hidden_actor=10, nn_stdev_actor=0.1,
                      learning_rate_actor=0.1, action_scaling=5, n_hidden_critic=10, nn_stdev_critic=0.1,
                      learning_rate_critic=0.1, tau_target_critic=0.01, tracked_state=1, ac_states=[3],
                      reward_weight=0.01, **kw)
    lon = h3.DHPAgent(control_channel='lon', discount_factor=0.99, n_hidden_actor=6, nn_stdev_actor=0.2,
                      learning_rate_actor=0.2, action_scaling=15, n_hidden_critic=6, nn_stdev_critic=0.2,
                      learning_rate_critic=0.2, tau_target_critic=1, tracked_state=4, ac_states=[5], **kw)
    for name, ag in (("col", col), ("lon", lon)):
        nets = (ag.actor.fc1, ag.actor.fc2, ag.critic.fc1, ag.critic.fc2, ag.target_critic.fc1, ag.target_critic.fc2)
        for key, layer in zip(("a_w1", "a_w2", "c_w1", "c_w2", "t_w1", "t_w2"), nets):
            assert np.array_equal(layer.weight[..., 0].double().numpy(), g["snap_%s_%s" % (name, key)][0]), (name, key)
    assert lon.actor.scale == np.deg2rad(15) and [p.shape for p in lon.networks[1].parameters()] == [(6, 2, 1), (2, 6, 1)]
    assert lon.control_channel_str == 'lon' and lon.control_channel_num == 1
    with pytest.raises(NotImplementedError):
        h3.DHPAgent(control_channel='lon', discount_factor=0.9, n_hidden_actor=6, nn_stdev_actor=0.1,
                    learning_rate_actor=0.1, action_scaling=5, action_network_final_layer='relu', n_hidden_critic=6,
                    nn_stdev_critic=0.1, learning_rate_critic=0.1, tau_target_critic=1, tracked_state=4,
                    ac_states=[5], device="cpu")


def test_checkpoints_round_trip(tmp_path):
    ag = h3.DHPAgent(control_channel='lon', discount_factor=0.99, n_hidden_actor=6, nn_stdev_actor=0.2,
                     learning_rate_actor=0.2, action_scaling=15, action_network_final_layer='tanh',
                     n_hidden_critic=6, nn_stdev_critic=0.2, learning_rate_critic=0.2, tau_target_critic=1,
                     tracked_state=4, ac_states=[5], n_envs=3, device="cpu", init="batched")
    ag.save(str(tmp_path / "lon.pt"))
    other = h3.DHPAgent(control_channel='lon', discount_factor=0.99, n_hidden_actor=6, nn_stdev_actor=0.2,
                        learning_rate_actor=0.2, action_scaling=15, action_network_final_layer='tanh',
                        n_hidden_critic=6, nn_stdev_critic=0.2, learning_rate_critic=0.2, tau_target_critic=1,
                        tracked_state=4, ac_states=[5], n_envs=3, device="cpu", init="batched")
    assert not torch.equal(ag.net, other.net)
    other.load(str(tmp_path / "lon.pt"))
    assert torch.equal(ag.net, other.net)
    # n_envs == 1 writes the reference's checkpoint format (agents.py:198-203): plain (out, in) matrices
    one = h3.DHPAgent(control_channel='lon', discount_factor=0.99, n_hidden_actor=6, nn_stdev_actor=0.2,
                      learning_rate_actor=0.2, action_scaling=15, action_network_final_layer='tanh',
                      n_hidden_critic=6, nn_stdev_critic=0.2, learning_rate_critic=0.2, tau_target_critic=1,
                      tracked_state=4, ac_states=[5], device="cpu")
    one.save(str(tmp_path / "one.pt"))
    ck = torch.load(str(tmp_path / "one.pt"))
    assert ck["actor_state_dict"]["fc1.weight"].shape == (6, 2) and ck["critic_state_dict"]["fc2.weight"].shape == (2, 6)
    other.load(str(tmp_path / "one.pt"))          # a reference-format file broadcasts over the batch
    assert torch.equal(other.actor.fc1.weight[..., 2], one.actor.fc1.weight[..., 0])

    rls = h3.RecursiveLeastSquares(state_size=7, action_size=2, gamma=1, covariance=10 ** 4, constant=False,
                                   n_envs=2, device="cpu")
    assert rls.W.shape == (9, 7, 2) and rls.cov.shape == (9, 9, 2) and rls.skip_update
    assert torch.equal(rls.cov[..., 1], torch.eye(9, dtype=torch.float64) * 1e4)
    assert rls.gradient_state().shape == (7, 7, 2) and rls.gradient_action().shape == (7, 2, 2)
    rls.W = torch.arange(63, dtype=torch.float64).reshape(9, 7)
    assert float(rls.gradient_state()[2, 5, 1]) == 5 * 7 + 2        # F = W[:7].T  (model.py:72)
    assert float(rls.gradient_action()[3, 1, 0]) == 8 * 7 + 3       # G = W[7:9].T (model.py:76)
    rls.save(str(tmp_path / "rls.pkl"))
    back = h3.RecursiveLeastSquares.load(str(tmp_path / "rls.pkl"), device="cpu")
    assert torch.equal(back.W, rls.W) and torch.equal(back.cov, rls.cov) and back.assume_symmetric
    c = torch.eye(9, dtype=torch.float64)
    c[0, 1] = 0.5
    back.cov = c
    assert not back.assume_symmetric                                # non-symmetric input -> general kernel
    with pytest.raises(NotImplementedError):
        h3.RecursiveLeastSquares(state_size=15, action_size=2, constant=False, device="cpu")
    # model.py:94-96 / train.py:70-72: the estimator OBJECT is pickled and unpickled, and a reference estimator's
    # attribute dict (numpy arrays without an episode axis) is understood
    import pickle
    single = h3.RecursiveLeastSquares(state_size=7, action_size=2, gamma=1, covariance=10 ** 4, constant=False, device="cpu")
    single.W = torch.arange(63, dtype=torch.float64).reshape(9, 7)
    single.save(str(tmp_path / "one.pkl"))
    with open(str(tmp_path / "one.pkl"), "rb") as f:
        obj = pickle.load(f)                                        # what train.py:70-72 does
    assert isinstance(obj, h3.RecursiveLeastSquares) and torch.equal(obj.W, single.W) and obj.skip_update
    st = single.__getstate__()
    assert st["W"].shape == (9, 7) and st["cov"].shape == (9, 9) and st["x"].shape == (7, 1) and st["skip_update"] is True

    class RefLike:                                                  # the attributes a reference estimator pickles
        pass
    ref = RefLike()
    ref.state_size, ref.action_size, ref.gamma, ref.covariance, ref.constant = 7, 2, 1, 10 ** 4, False
    ref.nb_vars = ref.nb_coefficients = 9
    ref.W, ref.cov = np.arange(63.0).reshape(9, 7), np.eye(9) * 3.0
    ref.x, ref.u, ref.X, ref.skip_update = np.ones((7, 1)), np.zeros((2, 1)), np.ones((9, 1)), False
    with open(str(tmp_path / "ref.pkl"), "wb") as f:
        pickle.dump(vars(ref), f)
    got = h3.RecursiveLeastSquares.load(str(tmp_path / "ref.pkl"))
    assert got.n_envs == 1 and torch.equal(got.W[..., 0], torch.as_tensor(ref.W)) and not got.skip_update
    assert float(got.cov[4, 4, 0]) == 3.0 and got.assume_symmetric and torch.equal(got.x, torch.ones(7, 1, dtype=torch.float64))
    r6 = h3.heli6dof.RecursiveLeastSquares6(state_size=15, action_size=2, gamma=1, covariance=1e8, constant=False, n_envs=2, device="cpu")
    r6.save(str(tmp_path / "r6.pkl"))
    b6 = h3.heli6dof.RecursiveLeastSquares6.load(str(tmp_path / "r6.pkl"))
    assert torch.equal(b6.W, r6.W) and torch.equal(b6.cov, r6.cov) and b6.n_envs == 2


def test_config_parsing_and_loud_failure_without_gpu(tmp_path):
    cfg = {"dt": 0.02, "training": {"trim_speed": 10},
           "env": {"tasks": {"velocity": {"kp": -1, "ki": -0.05, "kd": 0, "ref": 30}}}}
    path = tmp_path / "c.json"
    path.write_text(json.dumps(cfg))
    env = h3.Helicopter3DOF(dt=0.01, t_max=120, n_envs=2, device="cpu")
    env.task = "velocity"
    env._set_pid_weights(cfg)
    assert env.pid_weights == (-1, -0.05, 0) and env.ref == 30
    assert env.state.shape == (7, 2) and env.episode_ticks == 12000
    env.task = "hover"
    with pytest.raises(NotImplementedError):
        env._set_pid_weights(cfg)
    # no CPU fallback: anything that computes needs the CUDA library AND a CUDA tensor
    env.task = "velocity"
    with pytest.raises(h3.H3Error):
        env.setup_from_config("velocity", str(path))
    assert env.dt == 0.02                                           # the json was read before the compute call
    pid = h3.CollectivePID3DOF(h_ref=5, dt=0.01)
    with pytest.raises(h3.H3Error):
        pid(torch.zeros(7, 2, dtype=torch.float64))


def test_launch_plan_splits_at_the_phase_switch():
    class _E:
        n_envs, device, dtype = 1, torch.device("cpu"), torch.float64
    loop = h3.ClosedLoop3DOF.__new__(h3.ClosedLoop3DOF)
    loop.k, loop.switch_step, loop.ticks_per_launch = 5500, 6000, 400
    assert loop.plan(1000) == [(0, 400), (400, 100), (500, 400), (900, 100)]
    loop.k = 6000
    assert loop.plan(500) == [(0, 400), (400, 100)]
    loop.k = None
    assert loop.plan(300) == [(0, 300)]


def test_sweep_grid_and_compat_modules():
    keys, settings, n_total = h3.sweep.expand_grid({"a": [1, 2, 3], "b": [10, 20]}, n_seeds=4)
    assert keys == ["a", "b"] and settings.shape == (6, 2) and n_total == 24
    assert list(settings[1]) == [1, 20]
    # the parameter block of a shard is a slice of the whole job's block (setting = global index // n_seeds)
    k2, s2, n2 = h3.sweep.expand_grid({"tau_target_critic": [0.01, 1.0], "nn_stdev": [0.1, 0.2, 0.4]}, n_seeds=5)
    whole, so = h3.sweep.sweep_params(k2, s2, 5, 0, n2, "cpu")
    part, so_p = h3.sweep.sweep_params(k2, s2, 5, 7, 19, "cpu")
    assert torch.equal(part, whole[:, 7:19]) and torch.equal(so_p, so[7:19]) and list(so[:6]) == [0, 0, 0, 0, 0, 1]
    R = h3.episodes.ROW
    assert set(whole[R["lon_tau"]].tolist()) == {0.01, 1.0} and set(whole[R["col_tau"]].tolist()) == {0.01}
    assert set(whole[R["sigma_lon_critic"]].tolist()) == {0.1, 0.2, 0.4} and set(whole[R["sigma_col_actor"]].tolist()) == {0.1}
    # interleaved order (setting = global index % S): every contiguous shard holds the same mix of settings
    wi, si = h3.sweep.sweep_params(k2, s2, 5, 0, n2, "cpu", interleave=True)
    pi, si_p = h3.sweep.sweep_params(k2, s2, 5, 12, 24, "cpu", interleave=True)
    assert torch.equal(pi, wi[:, 12:24]) and torch.equal(si_p, si[12:24]) and list(si[:7]) == [0, 1, 2, 3, 4, 5, 0]
    assert torch.bincount(si).tolist() == [5] * 6 and torch.bincount(si_p).tolist() == [2] * 6
    assert sorted(wi[R["lon_tau"]].tolist()) == sorted(whole[R["lon_tau"]].tolist())
    # the reference's module names resolve to the batched classes when compat/ is on sys.path
    import importlib
    import sys
    sys.path.insert(0, os.path.join(ROOT, "compat"))
    try:
        for mod, name in (("agents", "DHPAgent"), ("model", "RecursiveLeastSquares"), ("heli_models", "Helicopter3DOF"),
                          ("PID", "CollectivePID3DOF")):
            sys.modules.pop(mod, None)
            assert getattr(importlib.import_module(mod), name) is getattr(h3, name)
        H6 = h3.heli6dof
        for mod, names in (("heli_models", ("Helicopter6DOF",)), ("PID", ("CollectivePID6DOF", "LatPedPID")),
                           ("util", ("Logger", "RefGenerator", "envelope_limits_reached")),
                           ("params", ("ac_params_train", "env_params_test2", "pid_params", "rls_params")),
                           ("train", ("train", "mp_train"))):
            sys.modules.pop(mod, None)
            for name in names:
                assert getattr(importlib.import_module(mod), name) is getattr(H6, name)
    finally:
        sys.path.remove(os.path.join(ROOT, "compat"))
        for mod in ("agents", "model", "heli_models", "PID", "util", "params", "train"):
            sys.modules.pop(mod, None)
