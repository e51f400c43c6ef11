"""TEST INFRASTRUCTURE -- not product code.  Only tests/, __graft_entry__.smoke() and the cpu_baseline / --impl
reference legs of bench.py may import this module.

NumPy restatement of the counter-based initial-weight draws of h3_init_episodes (include/heli3dof.h):

    agents.py:19-22, :42-45   every weight matrix ~ N(0, nn_stdev), no bias
    agents.py:107             target critic = deep copy of the critic

The reference draws from torch's global Mersenne-Twister stream, which ties an episode's weights to how many
draws happened before it; a batched implementation needs draws that depend on (seed, GLOBAL episode index, matrix,
element) only, so that a batch gives the same episodes however it is sharded.  The generator is Philox4x32-10
(Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3", SC'11; Random123 1.14), pinned below
against Random123's published known-answer vectors (tests/test_init.py), followed by a Box-Muller transform.

    counter = (episode_lo, episode_hi, 4 * agent + matrix, pair index j),  key = (seed_lo, seed_hi)
    agent: 0 col, 1 lon;  matrix: 0 actor fc1 (nh,2), 1 actor fc2 (1,nh), 2 critic fc1 (nh,2), 3 critic fc2 (2,nh)
    (r0, r1, r2, r3) -> u1 = (((r0 >> 5) << 26 | r1 >> 6) + 0.5) 2^-53,  u2 likewise from (r2, r3)
    z0 = sqrt(-2 ln u1) cos(2 pi u2) -> element 2 j,   z1 = sqrt(-2 ln u1) sin(2 pi u2) -> element 2 j + 1
    of the matrix flattened row-major in torch's (out, in) layout; weight = sigma * z.
"""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)

# Random123 kat_vectors: philox4x32, 10 rounds: (counter, key, output)
KNOWN_ANSWERS = (
    ((0x00000000, 0x00000000, 0x00000000, 0x00000000), (0x00000000, 0x00000000),
     (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff), (0xffffffff, 0xffffffff),
     (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
     (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
)

MATRICES = (("a_w1", 0), ("a_w2", 1), ("c_w1", 2), ("c_w2", 3))


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10: uint32 arrays in, four uint32 arrays out."""
    c0, c1, c2, c3 = (np.asarray(x, dtype=np.uint32).copy() for x in np.broadcast_arrays(c0, c1, c2, c3))
    k0 = np.asarray(k0, dtype=np.uint32).copy()
    k1 = np.asarray(k1, dtype=np.uint32).copy()
    mask = np.uint64(0xFFFFFFFF)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = M0 * c0.astype(np.uint64)
            p1 = M1 * c2.astype(np.uint64)
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), (p0 & mask).astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), (p1 & mask).astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0 = k0 + W0
            k1 = k1 + W1
    return c0, c1, c2, c3


def _uniform53(hi, lo):
    m = ((hi.astype(np.uint64) >> np.uint64(5)) << np.uint64(26)) | (lo.astype(np.uint64) >> np.uint64(6))
    return (m.astype(np.float64) + 0.5) * 2.0 ** -53


def normal_pairs(seed, episode, stream, n_pairs):
    """(2 * n_pairs, N) standard normals of `stream` for the global episode indices `episode` (N,)."""
    episode = np.asarray(episode, dtype=np.uint64)
    j = np.arange(n_pairs, dtype=np.uint32)[:, None]
    ep_lo = (episode & np.uint64(0xFFFFFFFF)).astype(np.uint32)[None, :]
    ep_hi = (episode >> np.uint64(32)).astype(np.uint32)[None, :]
    seed = int(seed)
    r0, r1, r2, r3 = philox4x32_10(ep_lo, ep_hi, np.uint32(stream), j, np.uint32(seed & 0xFFFFFFFF),
                                   np.uint32((seed >> 32) & 0xFFFFFFFF))
    u1, u2 = _uniform53(r0, r1), _uniform53(r2, r3)
    rad = np.sqrt(-2.0 * np.log(u1))
    ang = 2.0 * np.pi * u2
    out = np.empty((2 * n_pairs, episode.shape[0]))
    out[0::2] = rad * np.cos(ang)
    out[1::2] = rad * np.sin(ang)
    return out


def initial_weights(seed, episode, nh_col=10, nh_lon=6, sigma=(0.1, 0.1, 0.2, 0.2)):
    """`weights` dict for heli3dof_np.init_state: the draws h3_init_episodes makes for the global episode indices
    `episode` (N,).  sigma = (col actor, col critic, lon actor, lon critic), scalars or (N,) arrays."""
    out = {}
    for a, (name, nh) in enumerate((("col", nh_col), ("lon", nh_lon))):
        shapes = {"a_w1": (nh, 2), "a_w2": (1, nh), "c_w1": (nh, 2), "c_w2": (2, nh)}
        for key, m in MATRICES:
            shp = shapes[key]
            cnt = shp[0] * shp[1]
            z = normal_pairs(seed, episode, 4 * a + m, (cnt + 1) // 2)[:cnt]
            sg = np.asarray(sigma[2 * a + (0 if key[0] == "a" else 1)], dtype=np.float64)
            out[name + "_" + key] = (z * sg).reshape(shp + (z.shape[1],))
    return out
