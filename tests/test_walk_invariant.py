"""The pruning rule of dense mode (sweep_walk_kernel, phaserotate/lv2_b200/csrc/kernels.cuh) restated in numpy
float32 and checked against brute force on the CPU: every grid angle at which a point would raise the running
peak must be among the angles the walk evaluates (or the point must go to the wide list, where every angle is
evaluated).  The GPU tests check the kernel itself bit for bit against brute force on audio; this test checks
the RULE - first window, stop at the global threshold, hop at a sector threshold, margins, the fast atan2 -
on adversarial tables and points that audio rarely produces (thresholds far apart between neighbouring
sectors, points 1e-7 above a threshold, directions exactly between two grid angles, the grid's wrap-around).
"""
import numpy as np
import pytest

F = np.float32
K_SECTORS = 60


def lut(MS):
    """(ca, sa) of grid angle j: cos / sin of -j pi / MS in fp32 (cli/phase-rotate.cc:41-72 at the grid of the test)."""
    a = -np.arange(MS, dtype=np.float64) * np.pi / MS
    return np.cos(a).astype(F), np.sin(a).astype(F)


def eval_y(ca, sa, qx, qy):
    """sweep_kernel's expression fabsf (fmaf (w.x, q.x, w.y * q.y)) for one point at every angle."""
    t = (sa * F(qy)).astype(F)
    return np.abs((ca.astype(np.float64) * np.float64(qx) + t.astype(np.float64)).astype(F))


def atan2_fast(y, x):
    ax, ay = abs(F(x)), abs(F(y))
    mx, mn = max(ax, ay), min(ax, ay)
    z = F(mn * (F(1.0) / max(mx, F(1e-30))))
    z2 = F(z * z)
    p = F(F(-0.0117212) * z2 + F(0.05265332))
    for c in (-0.11643287, 0.19354346, -0.33262347, 0.99997726):
        p = F(p * z2 + F(c))
    a = F(z * p)
    if ay > ax:
        a = F(F(1.57079632679) - a)
    if x < 0:
        a = F(F(3.14159265359) - a)
    return F(-a) if y < 0 else a


def walk(qx, qy, ca, sa, sec, MS, WH, slot_base=0, A=None):
    """Angles (slots) the kernel evaluates for the point, or None when the point goes to the wide list."""
    A = MS if A is None else A
    G = MS // K_SECTORS
    inv_step = F(MS / 3.14159265358979)
    sec = (sec * F(1.0 - 4e-6)).astype(F)
    tg = sec.min()
    tgs = F(F(-2e-6) * F(abs(F(qx)) + abs(F(qy))) + tg)
    j0 = int(np.rint(F(-atan2_fast(qy, qx)) * inv_step))
    j0 %= MS
    k0 = j0 - slot_base
    kc = min(max(k0, WH), A - 1 - WH)
    y = eval_y(ca, sa, qx, qy)                       # y[slot]
    seen = set(range(kc - WH, kc + WH + 1))
    yl, yr = y[kc - WH], y[kc + WH]
    done = 2 * WH + 1
    limit = 2 * WH + 1 + 12
    if not (max(abs(F(qx)), abs(F(qy))) > 0 and (kc != k0 or yl >= tgs or yr >= tgs)):
        return seen
    for d in (0, 1):
        room = (MS - 1) // 2 if d else MS // 2
        t = 1 if kc != k0 else ((WH + 1) if (yl if d else yr) >= tgs else MS)
        while t <= room:
            j = (j0 - t if d else j0 + t) % MS
            k = j - slot_base
            s = int(np.floor((F(j) + F(0.5)) * F(1.0 / G)))
            yy = np.inf
            if 0 <= k < A:
                yy = y[k]
                seen.add(k)
            done += 1
            if not yy >= tgs:
                break
            t += 1 if yy >= F(sec[s] - F(tg - tgs)) else ((j - s * G + 1) if d else ((s + 1) * G - j))
            if done > limit:
                break
        if done > limit:
            return None
    return seen


def tables(rng, MS, kind):
    j = np.arange(MS)
    if kind == "flat":
        pk = np.full(MS, 0.5, F)
    elif kind == "smooth":
        pk = (0.5 + 0.2 * np.cos(2 * np.pi * j / MS + rng.uniform(0, 6))).astype(F)
    elif kind == "steps":                               # neighbouring sectors far apart
        pk = np.repeat(rng.choice([0.2, 0.5, 0.9], K_SECTORS), MS // K_SECTORS).astype(F)
    else:                                                # one deep notch
        pk = np.full(MS, 0.8, F)
        lo = rng.integers(0, MS)
        pk[(lo + np.arange(5)) % MS] = F(0.3)
    return (pk * (1 + 1e-6 * rng.standard_normal(MS))).astype(F)


@pytest.mark.parametrize("MS,WH", [(180, 1), (720, 2), (1800, 3), (3600, 6)])
@pytest.mark.parametrize("kind", ["flat", "smooth", "steps", "notch"])
def test_every_angle_a_point_can_raise_is_evaluated(MS, WH, kind):
    rng = np.random.default_rng(1234 + MS + len(kind))
    ca, sa = lut(MS)
    pk = tables(rng, MS, kind)
    sec = pk.reshape(K_SECTORS, MS // K_SECTORS).min(1)
    step = np.pi / MS
    n_wide = n_walked = 0
    for trial in range(400):
        mode = trial % 4
        if mode == 0:      # direction exactly between two grid angles, radius just above the local peak
            jj = rng.integers(0, MS)
            phi = -(jj + 0.5) * step
            r = float(pk[jj]) * (1 + rng.choice([1e-7, 1e-6, 1e-5, 1e-4]))
        elif mode == 1:    # random direction, radius around the smallest threshold
            phi = rng.uniform(-np.pi, np.pi)
            r = float(sec.min()) * (1 + rng.uniform(-1e-5, 3e-5))
        elif mode == 2:    # well above everything nearby
            phi = rng.uniform(-np.pi, np.pi)
            r = float(pk.max()) * rng.uniform(0.9, 1.5)
        else:              # near the wrap-around of the grid
            phi = rng.choice([0.0, np.pi, -np.pi]) + rng.uniform(-3, 3) * step
            r = float(pk[0]) * (1 + rng.uniform(-1e-5, 1e-4))
        qx, qy = F(r * np.cos(phi)), F(r * np.sin(phi))
        for slot_base, A in ((0, MS), (1, MS - 1)):      # the whole grid / the CLI's sweep (angle 0 left out)
            seen = walk(qx, qy, ca[slot_base:slot_base + A], sa[slot_base:slot_base + A], sec, MS, WH, slot_base, A)
            y = eval_y(ca[slot_base:slot_base + A], sa[slot_base:slot_base + A], qx, qy)
            raises = np.nonzero(y > pk[slot_base:slot_base + A])[0]
            if seen is None:
                n_wide += 1
                continue
            n_walked += 1
            missed = [int(k) for k in raises if int(k) not in seen]
            assert not missed, (MS, kind, trial, slot_base, float(qx), float(qy), missed[:5])
    assert n_walked > 100          # the rule was exercised, not bypassed through the wide list
