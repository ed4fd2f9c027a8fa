"""CPU ORACLE (numpy / plain Python) for SURVEY.md section 8(f) row N4 -- DSMC -- TEST INFRASTRUCTURE ONLY.

Direct Simulation Monte Carlo between two kinetic species, restated line by line from
    Chemistry/src/dsmc.jl:1-142   (DSMC.ElasticCollision, cache!, perform!, PIC.perform!(dsmc, ...))
    Chemistry/src/cross_section.jl:15-16  (maximum / argmax of a CrossSection)
Only tests/, __graft_entry__.smoke() and bench.py's cpu legs may import this module.  The reference holds no test
or stored output for DSMC, and every result depends on Julia's MersenneTwister stream ("parity unpinned"): the
restatement is pinned by what does NOT depend on the stream -- the number of candidate pairs per cell and its
fractional carry (dsmc.jl:109-119), exact conservation of momentum and energy in equal-weight collisions (:29-65) --
and device results are compared statistically.

Reference quirks kept:
  D1  cache! (:25-30) is called inside the loop over collisions (:98-99) on lists created once per perform! (:94-95):
      with more than one collision the lists hold duplicates and indices of other species.  Only the first collision
      of a DSMC object is well defined; oracle and device take exactly one collision.
  D2  unequal macro-weights (:66-78): the target update writes `target.v[s,:]` (source index) with `mr2`.
  D3  `while source != target && sR == tR` (:124-126) compares row indices of two different species.
  D4  sigma_g_max = maximum(sigma) * argmax(sigma) (:107): the largest sigma times the abscissa where it occurs.
  D5  IonizationCollision (:8-13) has no perform! method: such a reaction raises MethodError when it fires.
"""
import math

import numpy as np


class ElasticCollision:
    """DSMC.ElasticCollision  dsmc.jl:2-6"""

    def __init__(self, rate, source, target):
        self.rate, self.source, self.target = rate, source, target


class DirectSimulationMonteCarlo:
    """dsmc.jl:16-23"""

    def __init__(self, collision):
        self.collisions = [collision]
        self.collisions_remaining = None


def sigma_g_max(rate):
    """maximum(rate) * argmax(rate)  dsmc.jl:107 with cross_section.jl:15-16"""
    k = int(np.argmax(rate.nodes[:, 1]))
    return float(rate.nodes[k, 1] * rate.nodes[k, 0])


def cell_lists(species, nx, ny, dh):
    """cache!  dsmc.jl:25-30 -> dict (i,j) -> list of 1-based rows, plus counts (nx, ny)"""
    lists = {}
    cnt = np.zeros((nx, ny), dtype=np.int64)
    for p in range(1, species.np + 1):
        i = int(math.floor(1.0 + species.x[p - 1, 0] / dh[0]))
        j = int(math.floor(1.0 + species.x[p - 1, 1] / dh[1]))
        lists.setdefault((i, j), []).append(p)
        cnt[i - 1, j - 1] += 1
    return lists, cnt


def candidate_pairs(Na, Nb, Wa, Wb, dx, dy, dt, sgmax, same_species, remaining):
    """dsmc.jl:109-119 for one cell -> (floor(Nc), new remainder)"""
    if Wa > Wb:
        Pab, Pba = Wb / Wa, 1.0
    else:
        Pab, Pba = 1.0, Wa / Wb
    na = Na * Wa / (dx * dy)
    Nc = na * Nb * dt * sgmax
    Nc /= Pab + (Wb / Wa) * Pba
    if not same_species:
        Nc *= 2
    Nc += remaining
    k = int(math.floor(Nc))
    return k, Nc - k


def perform_collision_(c, s, t, rng):
    """perform!(collision::DSMC.ElasticCollision, s, t)  dsmc.jl:32-79 ; s, t are 1-based rows"""
    source, target = c.source, c.target
    mr1 = source.m / (source.m + target.m)
    mr2 = target.m / (source.m + target.m)
    g = source.v[s - 1, :] - target.v[t - 1, :]
    vc_cm = mr1 * source.v[s - 1, :] + mr2 * target.v[t - 1, :]
    B = 2 * rng.random() - 1.0                             # vss_inv == 1 branch  :43-47
    A = math.sqrt(1 - B ** 2)
    C = 2 * math.pi * rng.random()
    ng = math.sqrt(float(g[0] * g[0] + g[1] * g[1] + g[2] * g[2]))
    vr_cp = ng * np.array([B, A * math.cos(C), A * math.sin(C)])
    if source.wg[s - 1] == target.wg[t - 1]:
        source.v[s - 1, :] = vc_cm + mr2 * vr_cp
        target.v[t - 1, :] = vc_cm - mr1 * vr_cp
    else:
        Pab = target.wg[t - 1] / source.wg[s - 1]
        Pba = source.wg[s - 1] / target.wg[t - 1]
        R = rng.random()
        if Pab > R:
            source.v[s - 1, :] = vc_cm + mr2 * vr_cp
        if Pba > R:
            target.v[s - 1, :] = vc_cm - mr2 * vr_cp       # D2: as written in the reference


def perform_(dsmc, dt, grid, rng):
    """PIC.perform!(dsmc, E, dt, config)  dsmc.jl:87-142 -> (nu, n_candidate_pairs)"""
    nx, ny = grid.n
    dx, dy = grid.dh
    nu = np.zeros((nx, ny))
    if dsmc.collisions_remaining is None:
        dsmc.collisions_remaining = np.zeros((nx, ny))
    c = dsmc.collisions[0]
    source, target = c.source, c.target
    ls, cs = cell_lists(source, nx, ny, grid.dh)
    lt, ct = (ls, cs) if target is source else cell_lists(target, nx, ny, grid.dh)   # :98-99 two separate list arrays
    sgmax = sigma_g_max(c.rate)
    ncand = 0
    for i in range(1, nx + 1):
        for j in range(1, ny + 1):
            Na, Nb = int(cs[i - 1, j - 1]), int(ct[i - 1, j - 1])
            if Na < 2 or Nb < 2:
                continue
            k, rem = candidate_pairs(Na, Nb, source.w0, target.w0, dx, dy, dt, sgmax, target is source,
                                     dsmc.collisions_remaining[i - 1, j - 1])
            dsmc.collisions_remaining[i - 1, j - 1] = rem
            ncand += k
            a, b = ls[(i, j)], lt[(i, j)]
            for _ in range(k):
                sR = a[int(rng.integers(0, Na))]
                tR = b[int(rng.integers(0, Nb))]
                while (source is not target) and sR == tR:     # D3
                    tR = b[int(rng.integers(0, Nb))]
                d = source.v[sR - 1, :] - target.v[tR - 1, :]
                g = math.sqrt(float(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]))
                sg = float(c.rate(g)) * g
                if sg / sgmax < rng.random():
                    continue
                perform_collision_(c, sR, tR, rng)
                nu[i - 1, j - 1] += 1
    return nu, ncand
