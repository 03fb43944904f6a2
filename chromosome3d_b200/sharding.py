"""Host-side partitioning of the two natural shards of the DGSA path (no numerics here).

* Ensembles: a (matrix, alpha) job with its m models is the unit; jobs are independent, so ranks take
  disjoint job subsets and there is no data-path collective (the reference's only parallelism is
  test.sh launching one OS process per run, TS:4-12).  Static LPT split on the cost model N^2 * m.
* One large structure: contiguous ranges of the upper-triangle 128x128 tile list
  (chromosome3d_b200/csrc/large_n.cu: ln_alloc), mirrored here for planning and tests.
"""

TILE = 128


def job_cost(n_beads, n_models):
    """Relative cost of one job: pair work per force evaluation times models."""
    return float(n_beads) * float(n_beads) * float(n_models)


def shard_jobs(costs, world):
    """Longest-processing-time-first assignment.  costs: per-job cost.  Returns `world` lists of job
    indices (each sorted ascending); deterministic: ties go to the lower job index / lower rank."""
    order = sorted(range(len(costs)), key=lambda k: (-costs[k], k))
    load = [0.0] * world
    out = [[] for _ in range(world)]
    for k in order:
        r = min(range(world), key=lambda q: (load[q], q))
        out[r].append(k)
        load[r] += costs[k]
    return [sorted(s) for s in out]


def n_tiles(n_beads):
    nt = (n_beads + TILE - 1) // TILE
    return nt * (nt + 1) // 2


def tile_split(n_beads, world, ctas=148):
    """The engine's own split (chr3d_single_tile_split, host-only): rank cuts [world + 1], per-rank CTA cuts [world][ctas + 1]
    and the modelled cost of every rank's range."""
    import ctypes as C
    import numpy as np
    from . import api
    cut = np.zeros(world + 1, dtype=np.int64)
    cta = np.zeros((world, ctas + 1), dtype=np.int64)
    cost = np.zeros(world)
    api._check(api.lib().chr3d_single_tile_split(n_beads, world, ctas, cut.ctypes.data_as(C.POINTER(C.c_int64)), cta.ctypes.data_as(C.POINTER(C.c_int64)),
                                                 cost.ctypes.data_as(C.POINTER(C.c_double))))
    return cut, cta, cost


def tile_range(n_beads, rank, world):
    """[t0, t1) of the row-major upper-triangle tile list owned by `rank`: contiguous ranges of equal modelled cost, the split
    chr3d_single_create makes (every rank derives the same cuts)."""
    cut, _, _ = tile_split(n_beads, world, 1)
    return int(cut[rank]), int(cut[rank + 1])


def tile_coords(t, n_beads):
    """(I, J) of tile t in the row-major upper-triangle order."""
    nt = (n_beads + TILE - 1) // TILE
    i = 0
    while (i + 1) * nt - (i + 1) * i // 2 <= t:
        i += 1
    start = i * nt - i * (i - 1) // 2
    return i, i + (t - start)


def run_sharded(jobs, costs, runner, rank=0, world=1, gather=None):
    """Runs this rank's share of `jobs` with `runner(list_of_jobs) -> list_of_results` and returns the
    results of ALL jobs in job order on every rank.  `gather(obj) -> list over ranks` is the only
    communication (results, after the timed compute); default: single process."""
    shards = shard_jobs(costs, world)
    mine = shards[rank]
    res = runner([jobs[k] for k in mine]) if mine else []
    packed = list(zip(mine, res))
    parts = gather(packed) if gather is not None else [packed]
    out = [None] * len(jobs)
    for part in parts:
        for k, r in part:
            out[k] = r
    return out
