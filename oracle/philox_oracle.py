"""NumPy restatement of the kernel's counter-based draws (test infrastructure; see oracle/__init__.py).

The reference draws from NumPy's global MT19937 stream (her.py:108-116), which a counter-based
generator cannot reproduce, so bit-exact parity with the reference is established on the INJECTED
stream.  This file pins the other mode: it restates the Philox4x32-10 mapping documented in DESIGN.md
("Philox draws") so tests can check the kernel's own draws and feed them to the HER oracle.

    x = philox4x32_10(counter = (c_lo, c_hi, off_lo, off_hi), key = (seed_lo, seed_hi))   c = concat row
    ep = mulhi32(x0, E)   t = mulhi32(x1, T)   u_her = (x2 + .5) / 2^32   u_off = (x3 + .5) / 2^32
    module choice (RANDOM/CP modes): y = philox(counter with off_hi ^ 0x80000000);
        random: mulhi32(y0, N)     cp: searchsorted(cdf, (y0 + .5) / 2^32, side='right')
"""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    c0, c1, c2, c3 = [np.asarray(c, np.uint64) & MASK for c in (c0, c1, c2, c3)]
    k0, k1 = int(k0) & 0xFFFFFFFF, int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        n0 = hi1 ^ c1 ^ np.uint64(k0)
        n2 = hi0 ^ c3 ^ np.uint64(k1)
        c0, c1, c2, c3 = n0, lo1, n2, lo0
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return c0, c1, c2, c3


def mulhi32(x, n):
    return ((np.asarray(x, np.uint64) * np.uint64(n)) >> np.uint64(32)).astype(np.int64)


def u01(x):
    return (np.asarray(x, np.uint64).astype(np.float64) + 0.5) * (1.0 / 4294967296.0)


def draws(concat_rows, E_per_row, T, seed, call_offset, mode=None, n_tasks=0, cdf=None):
    """concat_rows: int array of concat indices; E_per_row: episodes in the buffer each row samples from."""
    c = np.asarray(concat_rows, np.uint64)
    c_lo, c_hi = c & MASK, c >> np.uint64(32)
    off = np.uint64(call_offset)
    off_lo, off_hi = int(off & MASK), int(off >> np.uint64(32))
    k0, k1 = int(seed) & 0xFFFFFFFF, (int(seed) >> 32) & 0xFFFFFFFF
    n = c.shape[0]
    x0, x1, x2, x3 = philox4x32_10(c_lo, c_hi, np.full(n, off_lo), np.full(n, off_hi), k0, k1)
    out = dict(ep=((x0 * np.asarray(E_per_row, np.uint64)) >> np.uint64(32)).astype(np.int64),
               t=mulhi32(x1, T), u_her=u01(x2), u_off=u01(x3), choice=np.full(n, -1, np.int64))
    if mode in ('random', 'cp'):
        y0, _, _, _ = philox4x32_10(c_lo, c_hi, np.full(n, off_lo), np.full(n, off_hi ^ 0x80000000), k0, k1)
        if mode == 'random':
            out['choice'] = mulhi32(y0, n_tasks)
        else:
            out['choice'] = np.minimum(np.searchsorted(np.asarray(cdf), u01(y0), side='right'), n_tasks - 1)
    return out


ACTION_NOISE_TAG = 0x40000000


def action_noise(u, max_u, noise_eps, random_eps, seed, call):
    """NumPy restatement of cur_action_noise (csrc/norm_adam.cu): the reference's exploration noise, ddpg.py:147-152,
    with the counter-based draws documented in include/curious_b200.h instead of the host MT19937 stream.

        x = philox(counter = (row, pair, call_lo, call_hi ^ TAG), key = seed)
        z(2p), z(2p+1) = Box-Muller((x0 + .5) / 2^32, (x1 + .5) / 2^32);  random action components from x2, x3
        eps-greedy draw of the row: philox(counter = (row, 0xFFFFFFFF, ...)).x0

    float64 arithmetic on a float32 action array updated in place, like the reference."""
    u = np.array(u, np.float32, copy=True).reshape(len(u), -1)
    n, dimu = u.shape
    pairs = (dimu + 1) // 2
    rows = np.repeat(np.arange(n, dtype=np.uint64), pairs)
    pidx = np.tile(np.arange(pairs, dtype=np.uint64), n)
    c2, c3 = int(call) & 0xFFFFFFFF, ((int(call) >> 32) & 0xFFFFFFFF) ^ ACTION_NOISE_TAG
    k0, k1 = int(seed) & 0xFFFFFFFF, (int(seed) >> 32) & 0xFFFFFFFF
    m = rows.shape[0]
    x0, x1, x2, x3 = philox4x32_10(rows, pidx, np.full(m, c2), np.full(m, c3), k0, k1)
    b0, _, _, _ = philox4x32_10(np.arange(n, dtype=np.uint64), np.full(n, 0xFFFFFFFF), np.full(n, c2), np.full(n, c3),
                                k0, k1)
    rad = np.sqrt(-2.0 * np.log(u01(x0)))
    ang = 6.283185307179586 * u01(x1)
    z = np.stack([rad * np.cos(ang), rad * np.sin(ang)], axis=1).reshape(n, 2 * pairs)[:, :dimu]
    ra = (-max_u + (max_u - (-max_u)) * np.stack([u01(x2), u01(x3)], axis=1)).reshape(n, 2 * pairs)[:, :dimu]
    u += noise_eps * max_u * z                                           # ddpg.py:148-149
    u = np.clip(u, -max_u, max_u)                                        # ddpg.py:150
    explore = (u01(b0) < random_eps).astype(np.int64).reshape(-1, 1)
    u += explore * (ra - u)                                              # ddpg.py:151
    return u, explore.reshape(-1).astype(bool)
