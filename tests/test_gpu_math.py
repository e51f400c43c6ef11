"""GPU: accuracy of the kernel's own fp64 tanh / sin / cos (msc-thesis_b200/csrc/h3_math.cuh) against NumPy
(glibc, < 1 ulp) over the ranges the closed loop visits, the table's grid points and the hand-over points of the
sin / cos branches."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

pytestmark = pytest.mark.gpu


def _probe(fn, x, x2=None):
    import msc_thesis_b200 as h3
    L = h3._lib
    xd = torch.as_tensor(x, device="cuda")
    y = torch.empty_like(xd)
    y2 = torch.empty_like(xd) if x2 is None else torch.as_tensor(x2, device="cuda")     # fn 3 / 4: the second input
    L.check(h3.lib().h3_math_probe(fn, L.ptr(xd), L.ptr(y), L.ptr(y2), xd.numel(), L.stream_ptr(xd.device)))
    return y.cpu().numpy(), y2.cpu().numpy()


def _ulps(a, b):
    with np.errstate(invalid="ignore", divide="ignore"):
        return np.abs(a - b) / np.spacing(np.abs(b))


@pytest.mark.parametrize("fn,bound", [(0, 2.5), (2, 3.0)])
def test_tanh_accuracy(fn, bound):
    """fn 0: the table-driven tanh of the 3-DOF kernels (h3_math.cuh: tanh_tab; measured 2.0 ulp of NumPy, 1.4 ulp of the exact value), fn 2: the table-free
    formula of the 6-DOF per-tick kernels (tanh_exp; measured 2.4 ulp).  Random arguments over every range the loop
    visits, every grid point k / 16 of the table with both neighbours, the saturation point, signed zeros, inf, NaN."""
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    rng = np.random.default_rng(0)
    grid = np.arange(0, 321) / 16.0
    x = np.concatenate([rng.uniform(-0.55, 0.55, 200000), rng.uniform(-6, 6, 200000), rng.uniform(-40, 40, 50000),
                        rng.uniform(-1e-6, 1e-6, 10000), 10.0 ** rng.uniform(-300, 0, 10000),
                        grid, -grid, np.nextafter(grid, 100.0), np.nextafter(grid, -100.0), -np.nextafter(grid, 100.0),
                        [0.0, -0.0, 1e-300, 5e-324, 19.0, 19.1, 19.5, 25.0, 25.1, 700.0, -700.0, np.inf, -np.inf]])
    y, _ = _probe(fn, x)
    want = np.tanh(x)
    nz = want != 0          # the sign of a zero is not part of the contract
    assert np.array_equal(np.signbit(y[nz]), np.signbit(want[nz]))
    worst = float(np.nanmax(_ulps(y, want)))
    print("tanh fn %d: worst error %.2f ulp of NumPy's" % (fn, worst))
    assert worst <= bound
    assert np.all(np.abs(y) <= 1.0) and np.array_equal(y[np.abs(x) >= 19.5], np.sign(x[np.abs(x) >= 19.5]))
    yn, _ = _probe(fn, np.array([np.nan]))
    assert np.isnan(yn[0])


def test_sincos_accuracy():
    """One branch-free formula (Cody-Waite by pi/2 in three pieces + a polynomial pair) for every argument: flight angles,
    the pi/4 hand-over of the polynomial pair, many periods, and far beyond anything a live episode produces."""
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    rng = np.random.default_rng(1)
    x = np.concatenate([rng.uniform(-np.pi / 4, np.pi / 4, 200000), rng.uniform(-1e-6, 1e-6, 10000),
                        rng.uniform(-10, 10, 100000), rng.uniform(-1e6, 1e6, 10000), rng.uniform(-1e9, 1e9, 10000),
                        np.arange(-64, 65) * (np.pi / 4), np.nextafter(np.arange(-64, 65) * (np.pi / 4), 1e9),
                        [0.0, -0.0, 5e-324, 1e-300, np.pi / 4, -np.pi / 4, np.nextafter(np.pi / 4, 1), 1e12, -3e14]])
    s, c = _probe(1, x)
    assert float(np.nanmax(_ulps(s, np.sin(x)))) <= 3.0
    assert float(np.nanmax(_ulps(c, np.cos(x)))) <= 3.0
    assert np.all(np.abs(s) <= 1.0) and np.all(np.abs(c) <= 1.0)
    s, c = _probe(1, np.array([np.nan, np.inf, -np.inf]))
    assert np.isnan(s).all() and np.isnan(c).all()


def test_atan2_division_sqrt_accuracy():
    """The call-free replacements of atan2, / and sqrt (h3_math.cuh: atan2_nr, div_nr, sqrt_nr) over the magnitudes the
    models produce, every octant, the hand-over ratio tan(pi/8), exact zeros and NaN."""
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    rng = np.random.default_rng(2)
    n = 200000
    r = 10.0 ** rng.uniform(-6, 3, n)
    ang = rng.uniform(-np.pi, np.pi, n)
    yy = np.concatenate([r * np.sin(ang), [0.0, 1.0, -1.0, 0.0, 3.0, -3.0, np.tan(np.pi / 8), 1e-300, 0.0]])
    xx = np.concatenate([r * np.cos(ang), [1.0, 0.0, 0.0, -2.0, 3.0, 3.0, 1.0, 1e-300, 0.0]])
    got, _ = _probe(3, yy, xx)
    want = np.arctan2(yy, xx)
    nz = want != 0
    assert float(np.nanmax(_ulps(got[nz], want[nz]))) <= 3.0 and np.all(got[~nz] == 0.0)
    assert np.isnan(_probe(3, np.array([np.nan, 1.0]), np.array([1.0, np.nan]))[0]).all()
    a = np.concatenate([rng.normal(size=n) * 10.0 ** rng.uniform(-8, 8, n), [0.0, 1.0, -7.0]])
    b = np.concatenate([rng.normal(size=n) * 10.0 ** rng.uniform(-8, 8, n), [3.0, 3.0, 6.0]])
    b[b == 0] = 1.0
    q, _ = _probe(4, a, b)
    w = a / b
    assert float(np.nanmax(_ulps(q[w != 0], w[w != 0]))) <= 1.0 and np.all(q[w == 0] == 0.0)
    v = np.concatenate([10.0 ** rng.uniform(-200, 200, n), rng.uniform(0, 4, n), [0.0, 1.0, 4.0, 2.0, 1e-300, 5e-324]])
    sq, _ = _probe(5, v)
    w = np.sqrt(v)
    big = v >= 1e-290
    assert float(np.nanmax(_ulps(sq[big], w[big]))) <= 1.0 and np.all(sq[~big] == 0.0)
    assert np.isnan(_probe(5, np.array([-1.0, np.nan]))[0]).all()


@pytest.mark.parametrize("warps,blocks", [(4, 1), (8, 1), (4, 2 * 148), (8, 148), (4, 8 * 148)])
def test_tensor_memory_as_thread_private_storage(warps, blocks):
    """h3_tmem_probe: every thread of 4- and 8-warp CTAs (one CTA, one wave, several waves per SM) fills its 128 private
    fp64 words of tensor memory (tcgen05.st 32x32b), reads them back in the piece sizes the fused kernel uses
    (tcgen05.ld x16 + x8 at runtime offsets, x16 tail), runs read-modify-write passes and gets exact values back."""
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import msc_thesis_b200 as h3
    L = h3._lib
    reps = 8
    nt = blocks * warps * 32
    out = torch.zeros(137, nt, dtype=torch.float64, device="cuda")
    cyc = torch.zeros(2 * blocks, dtype=torch.int64, device="cuda")
    L.check(h3.lib().h3_tmem_probe(warps, blocks, reps, L.ptr(out), L.ptr(cyc), L.stream_ptr(out.device)))
    torch.cuda.synchronize()
    gt = torch.arange(nt, dtype=torch.float64, device="cuda")
    want = gt[None, :] * 1024 + torch.arange(128, dtype=torch.float64, device="cuda")[:, None]
    assert torch.equal(out[:128], want)
    assert torch.equal(out[128:136], want[:8] + reps)
    assert int(cyc.min()) > 0
