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


def tile_range(n_beads, rank, world):
    """[t0, t1) of the row-major upper-triangle tile list owned by `rank` (balanced by pair count)."""
    n = n_tiles(n_beads)
    return n * rank // world, n * (rank + 1) // world


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
