"""Device-side episode initialisation (h3_init_episodes) and the runners built on it (EpisodeRunner, sweep).

CPU: the oracle's Philox4x32-10 against Random123's known-answer vectors, the distribution and shard independence
of the draws.  GPU (through the C ABI): every field the kernel writes against the oracle's init_state with the same
draws; a batch equals the union of its shards; >= 1e5 random sweep episodes (per-episode hyper-parameters, the
tensor-memory kernel's SWEEP variant) against the compiled C oracle."""
import os

import numpy as np
import pytest

from oracle import heli3dof_np as O
from oracle import philox_np as PH
from tests import helpers as H


# ---------------------------------------------------------------------------------------------------------
# CPU
# ---------------------------------------------------------------------------------------------------------
def test_philox_known_answer_vectors():
    """Random123 kat_vectors, philox4x32 with 10 rounds."""
    for ctr, key, want in PH.KNOWN_ANSWERS:
        out = PH.philox4x32_10(*[np.array([c], dtype=np.uint32) for c in ctr], *[np.array([k], dtype=np.uint32) for k in key])
        assert tuple(int(o[0]) for o in out) == want


def test_draws_are_standard_normal_and_depend_on_the_global_index_only():
    ep = np.arange(200000)
    z = PH.normal_pairs(seed=7, episode=ep, stream=2, n_pairs=3)
    assert z.shape == (6, 200000) and np.isfinite(z).all()
    assert abs(z.mean()) < 5e-3 and abs(z.std() - 1.0) < 5e-3
    assert abs(np.mean(z ** 3)) < 2e-2 and abs(np.mean(z ** 4) - 3.0) < 5e-2
    assert abs(np.corrcoef(z[0], z[1])[0, 1]) < 1e-2 and abs(np.corrcoef(z[0, :-1], z[0, 1:])[0, 1]) < 1e-2
    part = PH.normal_pairs(seed=7, episode=ep[1234:5678], stream=2, n_pairs=3)
    assert np.array_equal(part, z[:, 1234:5678])                        # a shard draws what the whole batch draws
    assert not np.array_equal(PH.normal_pairs(8, ep[:100], 2, 3), z[:, :100])          # seed matters
    assert not np.array_equal(PH.normal_pairs(7, ep[:100], 3, 3), z[:, :100])          # stream (matrix) matters
    big = PH.normal_pairs(7, np.array([2 ** 40 + 5, 5]), 0, 2)
    assert not np.array_equal(big[:, 0], big[:, 1])                     # the high word of the index is used
    w = PH.initial_weights(3, ep[:64], sigma=(0.1, 0.3, 0.2, 0.4))
    assert w["col_a_w1"].shape == (10, 2, 64) and w["lon_c_w2"].shape == (2, 6, 64) and w["col_a_w2"].shape == (1, 10, 64)
    big = PH.initial_weights(3, np.arange(20000), sigma=(0.1, 0.3, 0.2, 0.4))
    assert abs(big["col_c_w1"].std() / 0.3 - 1) < 0.02 and abs(big["lon_a_w2"].std() / 0.2 - 1) < 0.02


# ---------------------------------------------------------------------------------------------------------
# GPU
# ---------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def h3():
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import msc_thesis_b200 as m
    m.lib()
    return m


def oracle_init(n, seed, episode0, v, sigma, cfg):
    ids = np.arange(episode0, episode0 + n)
    w = PH.initial_weights(seed, ids, sigma=sigma)
    return O.init_state(n, v_initial=v, cfg=cfg, weights=w)


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["f64", "mixed"])
def test_init_episodes_matches_oracle(h3, precision):
    import torch
    n, seed, ep0 = 1000, 20201020, 2 ** 33 + 17
    rng = np.random.default_rng(5)
    v = rng.uniform(5.0, 25.0, n)
    sig = np.stack([rng.choice([0.05, 0.1, 0.2, 0.4], n) for _ in range(4)])
    run = h3.EpisodeRunner(n, device="cuda", precision=precision, per_episode_hyper=False, ticks_per_launch=50)
    P = h3.episodes.default_params(n, v_trim=v, sigma_col_actor=sig[0], sigma_col_critic=sig[1], sigma_lon_actor=sig[2],
                                   sigma_lon_critic=sig[3])
    run.params.copy_(P)
    R = h3.episodes.ROW
    run.loop.init_episodes(seed, episode0=ep0, v_trim=run.params[R["v_trim"]],
                           sigma=run.params[R["sigma_col_actor"]:R["sigma_lon_critic"] + 1])
    torch.cuda.synchronize()
    got = H.loop_state(run.loop)
    cfg = O.default_config()
    S = oracle_init(n, seed, ep0, v, sig, cfg)
    tol = 1e-12 if precision == "f64" else 2e-7         # mixed: fp32 networks
    for k in ("state", "trim", "ref", "rls_W", "rls_cov", "rls_x", "rls_u", "corr_u", "pid_corr", "t"):
        assert H.relerr(got[k], S[k], floor=1e-9) < 1e-12, k
    for a in ("col", "lon"):
        for k in O.NET_KEYS:
            d = np.abs(got[a + "_" + k] - S[a + "_" + k])
            assert d.max() <= tol * (np.abs(S[a + "_" + k]) + 1e-3).max(), (a, k, d.max())
        assert np.array_equal(got[a + "_t_w1"], got[a + "_c_w1"]) and np.array_equal(got[a + "_t_w2"], got[a + "_c_w2"])
    assert np.array_equal(got["step"], np.zeros(n)) and np.array_equal(got["flags"], np.zeros(n))
    assert np.array_equal(got["term_step"], -np.ones(n)) and np.array_equal(got["rls_skip"], np.ones(n))
    assert np.array_equal(got["ret"], np.zeros((2, n)))
    th = run.loop.env.trimmed_state[4].cpu().numpy()
    assert H.relerr(th, S["theta_trim"]) < 1e-12
    # and the episodes then run like the oracle's
    if precision == "f64":
        cfg["switch_step"], cfg["excitation"] = 20, H.excitation_closed_form()
        run.loop.switch_step = 20
        run.loop.run(50)
        torch.cuda.synchronize()
        O.run(S, cfg, 50, "all64")
        H.compare_states(H.loop_state(run.loop), S, 1e-8, what="after init")


@pytest.mark.gpu
def test_init_and_run_are_shard_independent(h3):
    """A batch is the union of its shards bit for bit: rank r initialises and runs episodes [lo, hi) with episode0 = lo."""
    import torch
    n, seed = 3000, 99
    kw = dict(device="cuda", per_episode_hyper=True, ticks_per_launch=40, switch_step=30, score_window=(10, 60))
    rng = np.random.default_rng(1)
    P = h3.episodes.default_params(n, v_trim=rng.uniform(5, 25, n), lon_lr_actor=rng.choice([0.1, 0.2, 0.4], n),
                                   lon_tau=rng.choice([0.01, 0.1, 1.0], n), col_gamma=rng.uniform(0.6, 0.99, n),
                                   ref_amplitude=rng.uniform(5, 15, n), ref_period=rng.uniform(5, 15, n))
    whole = h3.EpisodeRunner(n, **kw)
    out = whole.result_buffers(pin=False)
    whole.run(P, out, seed, n_ticks=80)
    torch.cuda.synchronize()
    st_whole = whole.loop.stats()
    acc_parts = np.zeros(8)
    for r in range(3):
        lo, hi = h3.shard_range(n, r, 3)
        part = h3.EpisodeRunner(hi - lo, **kw)
        o = part.result_buffers(pin=False)
        part.run(P[:, lo:hi].contiguous(), o, seed, n_ticks=80, episode0=lo)
        torch.cuda.synchronize()
        assert torch.equal(o[0], out[0][:, lo:hi]) and torch.equal(o[1], out[1][:, lo:hi])
        assert torch.equal(part.loop.env.state, whole.loop.env.state[:, lo:hi])
        assert torch.equal(part.loop.agent_col.net, whole.loop.agent_col.net[:, lo:hi])
        acc_parts += o[2].numpy()
    assert acc_parts[0] == n == st_whole["n"] and acc_parts[7] == st_whole["sum_ticks"]
    assert abs(acc_parts[5] - st_whole["sum_r_lon"]) <= 1e-9 * abs(st_whole["sum_r_lon"])
    # compaction between launches changes nothing but the work done
    comp = h3.EpisodeRunner(n, **kw)
    o2 = comp.result_buffers(pin=False)
    comp.run(P, o2, seed, n_ticks=80, compact_every=40)
    torch.cuda.synchronize()
    assert torch.equal(o2[0], out[0]) and torch.equal(o2[1], out[1])


@pytest.mark.gpu
def test_sweep_variant_1e5_random_episodes_vs_c_oracle(h3):
    """102,400 episodes of a main.py-style sweep (8 settings of per-episode tau / nn_stdev / gamma / learning rates on
    the cyclic agent x 12,800 seeds, random trim speeds): device-side init + 50 ticks through the phase switch (the
    per-episode-hyper variants of the tensor-memory kernel: the launch that ends before the switch keeps the cyclic
    agent's target critic in TMEM, the launch that contains it the collective agent) against the compiled C oracle run
    setting by setting on the same Philox draws."""
    import torch
    from oracle import c_oracle as CO
    grid = {"tau_target_critic": [0.1, 1.0], "nn_stdev": [0.1, 0.3], "discount_factor": [0.8, 0.95]}
    n_seeds, n_ticks, switch = 12800, 50, 30      # launches of 25 ticks: one ends before the switch, one contains it
    keys, settings, n = h3.sweep.expand_grid(grid, n_seeds)
    assert n == 102400
    seed = 4242
    runner = h3.EpisodeRunner(n, device="cuda", per_episode_hyper=True, switch_step=switch, score_window=(10, 50),
                              ticks_per_launch=25)
    params, setting_of = h3.sweep.sweep_params(keys, settings, n_seeds, 0, n, runner.device)
    rng = np.random.default_rng(11)
    v = rng.uniform(5.0, 25.0, n)
    params[h3.episodes.ROW["v_trim"]] = torch.as_tensor(v, device=runner.device)
    out = runner.result_buffers(pin=False)
    runner.run(params, out, seed, n_ticks=n_ticks)
    torch.cuda.synchronize()
    info = runner.loop.info(phase2=True)
    assert info["block"] == 128 and info["smem_bytes"] > 100000        # the tensor-memory kernel, per-episode-hyper variant
    assert info["blocks_per_sm"] == 2 and info["lanes_per_warp"] == 32  # 8 warps x 32 episodes per SM, one lane per episode
    info1 = runner.loop.info(phase2=False)      # launches that end before the switch: the same kernel, target critic in TMEM
    assert info1["block"] == 128 and info1["blocks_per_sm"] == 2 and info1["lanes_per_warp"] == 32
    got = H.loop_state(runner.loop)
    worst = 0.0
    for s_i, (tau, std, gamma) in enumerate(settings):
        lo, hi = s_i * n_seeds, (s_i + 1) * n_seeds
        cfg = O.default_config()
        cfg["switch_step"], cfg["excitation"] = switch, H.excitation_closed_form()
        hp = cfg["agents"]["lon"]
        hp["tau"], hp["gamma"] = float(tau), float(gamma)
        sig = (0.1, 0.1, float(std), float(std))
        S = oracle_init(hi - lo, seed, lo, v[lo:hi], sig, cfg)
        CO.run(S, cfg, n_ticks, threads=os.cpu_count() or 1)
        part = {k: (a[..., lo:hi] if isinstance(a, np.ndarray) else a) for k, a in got.items()}
        worst = max(worst, H.compare_states(part, S, 1e-8, what="setting %d" % s_i))
    res = h3.sweep.sweep_results(runner.loop, setting_of, len(settings), 40)
    assert res["n"].sum() == n and np.all(res["n"] == n_seeds)
    assert np.array_equal(res["success"], np.bincount(setting_of.cpu().numpy(), weights=(got["flags"] == 0), minlength=8))
    print("worst relative difference vs the C oracle over 102,400 episodes x 50 ticks: %.2e" % worst)


@pytest.mark.gpu
def test_sweep_scores_ignore_diverged_episodes(h3):
    """A diverged episode (NaN / -inf accumulator) must not poison its setting's mean score (NaN * 0 is NaN)."""
    import torch
    grid = {"learning_rate_actor": [0.1, 0.4]}
    keys, settings, res, loop = h3.sweep.run_sweep(grid=grid, n_seeds=16, n_ticks=60, device="cuda", switch_step=30,
                                                   score_window=(10, 60), ticks_per_launch=30)
    setting_of = torch.arange(32, device=loop.device) // 16
    loop.env.flags[3] = 2                       # forced: NaN in the actions inside the window
    loop.acc[2, 3] = float("nan")
    loop.env.flags[20] = 1
    loop.acc[2, 20] = float("-inf")
    r = h3.sweep.sweep_results(loop, setting_of, 2, 50)
    assert np.isfinite(r["mean_score_of_successes"]).all()
    assert r["success"][0] == res["success"][0] - 1 and r["success"][1] == res["success"][1] - 1
