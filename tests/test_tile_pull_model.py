"""CPU model of the 3-D near-far tile-pull rounds (raytracer.jl_b200/csrc/grid3d.cu: tp_release_kernel, tp_pull_kernel,
tp_round_control), written from the schedule's rules in numpy: pending / released node sets per 8 x 4 x 4 tile, tiles
activated through the released bounding box grown by the window half width, and the three prunings --
  (1) a tile is not visited for a source tile when 1 + the high word of its largest travel time (as of its last visit) is
      not above the high word of the smallest released travel time of that source tile,
  (2) a thread skips when the high word of its own travel time lies below the smallest released high word of the staged
      block,
  (3) sources that improve during a round are pending again (their released value may be read old or new).
The model must reach the oracle's travel times BIT FOR BIT on random velocities, every bucket width and either read
order: the prunings only ever drop candidates that cannot improve anything (weights >= 0, fp + monotone).  This checks
the rules, not the CUDA code (the GPU suite does that against the same oracle)."""
import numpy as np
import pytest

from helpers import weight3d

TX, TY, TZ = 8, 4, 4
INF = np.inf


def hi(v):
    """high 32 bits of non-negative doubles (monotone in the value)"""
    return (np.atleast_1d(np.asarray(v, np.float64)).view(np.uint64) >> np.uint64(32)).astype(np.int64)


def tile_pull_solve(nn, w, X, Y, Z, U, source, delta, live_reads, early=0):
    nx, ny, nz = nn
    n = nx * ny * nz
    tn = (-(-nx // TX), -(-ny // TY), -(-nz // TZ))
    ntile = tn[0] * tn[1] * tn[2]
    idx = np.arange(n)
    gx, gy, gz = idx % nx, (idx // nx) % ny, idx // (nx * ny)
    tile_of = (gx // TX) + tn[0] * ((gy // TY) + tn[1] * (gz // TZ))
    nodes_of = [np.flatnonzero(tile_of == t) for t in range(ntile)]
    dist = np.full(n, INF)
    dist[source] = 0.0
    pend = np.zeros(n, bool)
    pend[source] = True
    tmaxhi = np.full(ntile, 2 ** 32 - 1, np.int64)  # never visited
    tau, rounds, visits, pruned_tiles, gated = delta, 0, 0, 0, 0
    while True:
        rounds += 1
        assert rounds < 100000
        rel = pend & (dist < tau)
        if not rel.any():
            if not pend.any():
                break
            tau = dist[pend].min() + delta
            continue
        pend &= ~rel
        rel_val = np.where(rel, dist, INF)  # value at release time
        active = set()
        for t in np.unique(tile_of[rel]):
            r = nodes_of[t][rel[nodes_of[t]]]
            tminhi = hi(dist[r]).min()
            x0, x1 = gx[r].min() - w, gx[r].max() + w
            y0, y1 = gy[r].min() - w, gy[r].max() + w
            z0, z1 = gz[r].min() - w, gz[r].max() + w
            tx, ty, tz = t % tn[0], (t // tn[0]) % tn[1], t // (tn[0] * tn[1])
            for dz in (-1, 0, 1):
                for dy in (-1, 0, 1):
                    for dx in (-1, 0, 1):
                        ux, uy, uz = tx + dx, ty + dy, tz + dz
                        if not (0 <= ux < tn[0] and 0 <= uy < tn[1] and 0 <= uz < tn[2]):
                            continue
                        if not (x0 <= ux * TX + TX - 1 and x1 >= ux * TX and y0 <= uy * TY + TY - 1 and y1 >= uy * TY
                                and z0 <= uz * TZ + TZ - 1 and z1 >= uz * TZ):
                            continue
                        nb = ux + tn[0] * (uy + tn[1] * uz)
                        if tmaxhi[nb] > tminhi:  # pruning (1)
                            active.add(nb)
                        else:
                            pruned_tiles += 1
        n_rel = int(rel.sum())
        read = dist if live_reads else rel_val  # (3): a source may be read with a value it took later in the round
        for t in sorted(active):
            visits += 1
            mine = nodes_of[t]
            tx, ty, tz = t % tn[0], (t // tn[0]) % tn[1], t // (tn[0] * tn[1])
            blk = np.flatnonzero(rel & (gx >= tx * TX - w) & (gx <= tx * TX + TX - 1 + w) & (gy >= ty * TY - w) &
                                 (gy <= ty * TY + TY - 1 + w) & (gz >= tz * TZ - w) & (gz <= tz * TZ + TZ - 1 + w))
            bminhi = hi(read[blk]).min() if len(blk) else 2 ** 32 - 1
            for i in mine:
                if hi(dist[i])[0] < bminhi:  # pruning (2)
                    gated += 1
                    continue
                s = blk[(np.abs(gx[blk] - gx[i]) <= w) & (np.abs(gy[blk] - gy[i]) <= w) & (np.abs(gz[blk] - gz[i]) <= w)]
                s = s[s != i]
                if len(s) == 0:
                    continue
                cand = read[s] + weight3d(X, Y, Z, U, np.full(len(s), i), s)
                b = cand.min()
                if b < dist[i]:
                    dist[i] = b
                    pend[i] = True
            tmaxhi[t] = hi(dist[mine]).max() + 1
        if early and n_rel < early and pend.any():  # early threshold advance (option early_advance)
            tau = max(tau, dist[pend].min() + delta)
    return dist, dict(rounds=rounds, visits=visits, pruned_tiles=pruned_tiles, gated=gated)


@pytest.mark.parametrize("nn,lv,cs", [((13, 9, 10), 1, 1), ((17, 6, 5), 1, 0), ((10, 9, 9), 0, 1)])
def test_tile_pull_rules_reach_the_fixed_point(O, nn, lv, cs):
    R = 6371.0
    if cs:
        c0, c1 = (np.deg2rad(70.0), np.deg2rad(70.0), R - 2000.0), (np.deg2rad(110.0), np.deg2rad(110.0), R)
    else:
        c0, c1 = (0.0, 0.0, 0.0), (1.0, 1.0, 1.0)
    X, Y, Z = O.grid3d_coords(c0, c1, nn, cs)
    n = int(np.prod(nn))
    w = 1 << lv
    rng = np.random.default_rng(5 + n)
    seen = dict(pruned_tiles=0, gated=0)
    for U in (4.0 + 6.0 * rng.random(n), np.ones(n)):
        for src in (1, n // 2 + 3):
            want = O.bfm3d(nn, lv, X, Y, Z, U, src)[0]
            scale = float(np.median(want[np.isfinite(want)])) / 6.0
            for delta, live, early in ((scale, False, 0), (scale, True, 0), (scale / 7.0, True, 0), (1e9, False, 0),
                                       (scale / 3.0, False, 40)):
                got, info = tile_pull_solve(nn, w, X, Y, Z, U, src - 1, delta, live, early)
                assert np.array_equal(got, want), (nn, src, delta, live, info)
                seen["pruned_tiles"] += info["pruned_tiles"]
                seen["gated"] += info["gated"]
    assert seen["pruned_tiles"] > 0 and seen["gated"] > 0, "the prunings were never exercised: %s" % seen
