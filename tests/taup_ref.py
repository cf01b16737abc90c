"""Independent ray-theoretical first-arrival times T(Delta) for a surface source and surface receivers in the same
radial velocity table the solvers use (AK135 Vp on 1-km knots) -- the classical tau-p integrals

    Delta(p) = 2 * int_{r_t}^{R} p / (r sqrt(eta^2 - p^2)) dr,     T(p) = 2 * int_{r_t}^{R} eta^2 / (r sqrt(eta^2 - p^2)) dr,
    eta(r) = r / v(r),   p = ray parameter [s/rad],   eta(r_t) = p at the turning radius,

evaluated analytically per 1-km shell with Bullen's law v = a r^b fitted through the two knots of the shell (for which
both integrands have closed-form antiderivatives: acos(p/eta)/(1-b) and sqrt(eta^2-p^2)/(1-b)).  No graph, no mesh, no
code shared with the shortest-path solvers: an external anchor for the WHOLE chain (mesh + velocity + solver), the role
TauP plays for the reference's error.png.  Beyond the core shadow the first arrival of a shortest-path method is the
wave diffracted along the core-mantle boundary: T(p_cmb) + p_cmb * (Delta - Delta(p_cmb)).  Test infrastructure only."""
import numpy as np


def shells(r, v):
    """Per shell [r_k, r_{k+1}]: eta at both ends and 1/(1-b)."""
    r = np.asarray(r, np.float64)
    v = np.asarray(v, np.float64)
    keep = r > 0
    r, v = r[keep], v[keep]
    eta = r / v
    b = np.log(v[1:] / v[:-1]) / np.log(r[1:] / r[:-1])
    return r, v, eta, 1.0 / (1.0 - b)


def ray(p, r, eta, inv1mb):
    """(Delta [rad], T [s], turning radius) of the ray with parameter p leaving the surface downwards; descends shell by
    shell (through the core-mantle boundary too) until eta drops to p."""
    delta = 0.0
    t = 0.0
    k = len(r) - 2
    while k >= 0:
        top, bot = eta[k + 1], eta[k]
        if top <= p:  # cannot enter this shell: p is a grazing value at its top
            return 2 * delta, 2 * t, r[k + 1]
        c = inv1mb[k]
        if bot <= p:  # turns inside the shell
            delta += c * np.arccos(p / top)
            t += c * np.sqrt(top * top - p * p)
            # eta = r^(1-b)/a  ->  r_t = r_top * (p/top)^(1/(1-b))
            return 2 * delta, 2 * t, r[k + 1] * (p / top) ** c
        delta += c * (np.arccos(p / top) - np.arccos(p / bot))
        t += c * (np.sqrt(top * top - p * p) - np.sqrt(bot * bot - p * p))
        k -= 1
    return 2 * delta, 2 * t, r[0]


def rays(ps, r, eta, inv1mb):
    """Vectorised `ray` for an array of ray parameters: (Delta, T) arrays."""
    top, bot = eta[1:][::-1], eta[:-1][::-1]  # shells ordered from the surface downwards
    c = inv1mb[::-1]
    D = np.zeros(len(ps))
    T = np.zeros(len(ps))
    for a in range(0, len(ps), 256):
        p = ps[a:a + 256, None]
        stop = (top[None, :] <= p) | (bot[None, :] <= p)
        first = np.where(stop.any(axis=1), stop.argmax(axis=1), len(top))
        k = np.arange(len(top))[None, :]
        above = k < first[:, None]  # shells crossed completely
        with np.errstate(invalid="ignore"):
            at, ab = np.arccos(np.minimum(p / top, 1.0)), np.arccos(np.minimum(p / bot, 1.0))
            st, sb = np.sqrt(np.maximum(top * top - p * p, 0.0)), np.sqrt(np.maximum(bot * bot - p * p, 0.0))
        D[a:a + 256] = np.sum(np.where(above, c * (at - ab), 0.0), axis=1)
        T[a:a + 256] = np.sum(np.where(above, c * (st - sb), 0.0), axis=1)
        # the shell in which the ray turns (top > p >= bot); a ray stopped by top <= p grazes that interface instead
        rows = np.flatnonzero(first < len(top))
        kk = first[rows]
        turn = top[kk] > ps[a:a + 256][rows]
        rr, kt = rows[turn], kk[turn]
        D[a + rr] += c[kt] * at[rr, kt]
        T[a + rr] += c[kt] * st[rr, kt]
    return 2 * D, 2 * T


def first_arrivals(r, v, deltas_rad, n_p=6000):
    """Lower envelope over all turning rays (mantle and core) plus the CMB-diffracted line.  Returns T [s] per Delta."""
    r, v, eta, c = shells(r, v)
    ps = np.linspace(eta[-1] * (1 - 1e-9), 1e-3, n_p)
    D, T = rays(ps, r, eta, c)
    out = np.full(len(deltas_rad), np.inf)
    # every consecutive pair of rays spans a piece of a travel-time branch: linear interpolation inside it
    d0, d1, t0, t1 = D[:-1], D[1:], T[:-1], T[1:]
    close = np.abs(d1 - d0) < 0.02  # do not bridge jumps between branches (shadow zones, triplication ends)
    for k, dl in enumerate(deltas_rad):
        dl = min(dl, 2 * np.pi - dl)
        inside = close & (np.minimum(d0, d1) <= dl) & (np.maximum(d0, d1) >= dl) & (d0 != d1)
        if inside.any():
            w = (dl - d0[inside]) / (d1[inside] - d0[inside])
            out[k] = np.min(t0[inside] + w * (t1[inside] - t0[inside]))
        # diffraction along the core-mantle boundary: the grazing ray of the mantle side, i.e. of the knot just above the
        # largest velocity drop of the table (AK135 on 1-km knots: 13.69 -> 8.01 km/s between r = 3482 and 3481 km)
        kc = int(np.argmax(v[1:] / v[:-1]))
        p_c = eta[kc + 1]
        dc, tc, _ = ray(p_c * (1 + 1e-12), r, eta, c)
        if dl >= dc:
            out[k] = min(out[k], tc + p_c * (dl - dc))
    return out
