"""TEST INFRASTRUCTURE (oracle) -- numpy restatement of the site-keyed draw contract.

Philox4x32-10 is a third-party published algorithm (Salmon, Moraes, Dror, Shaw: "Parallel random
numbers: as easy as 1, 2, 3", SC'11; Random123 v1.09).  It does not exist in the reference; it
replaces the reference's four NumPy generators (SURVEY 8a RNG site table) as described in
include/hlynr_rng.h.  Pinned by the Random123 known-answer vectors in tests/test_rng.py.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this package.
"""
import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)

BLK_EVADE, BLK_WIND, BLK_UNI, BLK_GPOS, BLK_GVEL, BLK_GUST_DIR, BLK_GUST_MAG = 0, 1, 2, 3, 4, 5, 6
BLK_SPAWN0, BLK_SPAWN1, BLK_SPAWN2 = 8, 9, 10
BLK_DR0 = 12
BLK_ACT0, BLK_ACT1 = 16, 17


def blk_vspawn(m):
    """HLYNR_BLK_VSPAWN(m): spawn uniforms of volley missile m >= 1."""
    return 32 + m


def blk_vevade(m):
    """HLYNR_BLK_VEVADE(m): evasion normals of volley missile m >= 1."""
    return 48 + m


def philox4x32_10(ctr, key):
    """ctr: (..., 4) uint32, key: (..., 2) uint32 -> (..., 4) uint32."""
    ctr = np.asarray(ctr, dtype=np.uint32)
    key = np.asarray(key, dtype=np.uint32)
    c0, c1, c2, c3 = (ctr[..., i].astype(np.uint64) for i in range(4))
    k0 = key[..., 0].astype(np.uint64)
    k1 = key[..., 1].astype(np.uint64)
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        c0, c1, c2, c3 = (hi1 ^ c1 ^ k0) & MASK, lo1, (hi0 ^ c3 ^ k1) & MASK, lo0
        k0 = (k0 + np.uint64(W0)) & MASK
        k1 = (k1 + np.uint64(W1)) & MASK
    return np.stack([c0, c1, c2, c3], axis=-1).astype(np.uint32)


def block(seed, env_id, episode, step, blk):
    """Raw 4xuint32 block for scalar or broadcastable array arguments."""
    seed = int(seed)
    env_id = np.asarray(env_id, dtype=np.uint64)
    episode = np.asarray(episode, dtype=np.uint64)
    step = np.asarray(step, dtype=np.uint64)
    blk = np.asarray(blk, dtype=np.uint64)
    env_id, episode, step, blk = np.broadcast_arrays(env_id, episode, step, blk)
    ctr = np.stack([env_id & MASK, episode & MASK, step & MASK,
                    (blk | ((env_id >> np.uint64(32)) << np.uint64(16))) & MASK], axis=-1).astype(np.uint32)
    key = np.empty(ctr.shape[:-1] + (2,), dtype=np.uint32)
    key[..., 0] = seed & 0xFFFFFFFF
    key[..., 1] = (seed >> 32) & 0xFFFFFFFF
    return philox4x32_10(ctr, key)


def u01(x):
    """[0,1) float32, exact."""
    return ((np.asarray(x, dtype=np.uint32) >> np.uint32(8)).astype(np.float32)) * np.float32(2.0 ** -24)


def u01_open(x):
    """(0,1] float32, exact."""
    return (((np.asarray(x, dtype=np.uint32) >> np.uint32(8)) + np.uint32(1)).astype(np.float32)) * np.float32(2.0 ** -24)


def normals(raw):
    """raw (..., 4) uint32 -> (..., 4) float32 standard normals (float32 Box-Muller)."""
    raw = np.asarray(raw, dtype=np.uint32)
    out = np.empty(raw.shape, dtype=np.float32)
    for a in (0, 2):
        r = np.sqrt(np.float32(-2.0) * np.log(u01_open(raw[..., a])))
        t = np.float64(2.0) * u01(raw[..., a + 1]).astype(np.float64)  # exact
        out[..., a] = r * np.cos(np.pi * t).astype(np.float32)
        out[..., a + 1] = r * np.sin(np.pi * t).astype(np.float32)
    return out


def exponential(raw):
    return -np.log(u01_open(np.asarray(raw, dtype=np.uint32)[..., 0]))


def uniforms(raw):
    return u01(raw)


def random_actions(seed, env_ids, episode, step):
    """The synthetic random policy of hlynr_rollout(actions=NULL): a = 2u-1, float32, (N,6)."""
    b0 = u01(block(seed, env_ids, episode, step, BLK_ACT0))
    b1 = u01(block(seed, env_ids, episode, step, BLK_ACT1))
    a = np.concatenate([b0, b1[..., :2]], axis=-1)
    return (np.float32(2.0) * a - np.float32(1.0)).astype(np.float32)
