// ln_sparse.cuh — the sparse force provider of the large-N engine (included by large_n.cu inside namespace chr3d).
//
// Real high-resolution Hi-C is sparse: most bin pairs have IF = 0 and carry no restraint (C3D:148-152 writes -1, dist2rr drops
// it, C3D:196).  The dense tile engine would stream 4 B for every one of the N^2/2 pairs regardless; this provider makes the
// traffic proportional to the number of restraints:
//   * NOE term: the restraint table as a symmetric CSR (both (i, j) and (j, i), columns ascending): one warp per bead row, lanes
//     stride over the row, each entry is {column, target}; the force goes to the row's bead only (the mirrored entry serves
//     the other bead), so there are no atomics and every sum has a fixed order.  8 B per entry per evaluation.
//   * repel term: a cell list.  CNS keeps a neighbour list with cutnb = 4.5 / 6.0 A and rebuilds it when an atom has moved more
//     than `tolerance` (C3D:1395-1404, 1417-1426, 1729-1738); here beads are binned into cells of edge R_max + skin on a hashed
//     grid (no bounding box), the bins are rebuilt on the device when a bead has moved more than skin / 2 since the last
//     build, and every bead scans the 27 cells around it.  Beads inside a bin are kept
//     in ascending index order, so the summation order is fixed here too.
//   * bond (+ angle stand-in) term and the three energies in the same launch.
// The result is the total force per bead and the three energies — exactly what the multi-launch integrator / minimiser kernels of
// large_n.cu (ln_md_step_kernel, ln_min_step_kernel: the `fsum` path) consume.

#define SP_THREADS 256
#define SP_SKIN 2.0f                    // CNS: nbonds tolerance=0.5 (C3D:1402, 1736); a performance parameter only — a wider skin trades a few more
                                        // candidates per bead for fewer rebuilds (hot MD at 2000 K moves a bead 0.25 A in two or three steps)

struct SpGrid {                         // hashed cell grid
    float inv_cs; int hbits;            // 1 / cell edge, log2(buckets)
    int* cell_start;                    // [H + 1]
    int* cell_count;                    // [H]
    int* cell_fill;                     // [H] placement cursors
    int* order;                         // [N] beads sorted by (bucket, index)
    int* bsum;                          // [CTAs] bucket-segment totals of the scan
    int4* ent;                          // [N] the sorted beads with the cell they were binned in: {bead, cx, cy, cz}
    float4* xref;                       // [N] coordinates at the last build
    int* rebuild;                       // device flag: a bead has moved more than skin / 2
    long long* n_rebuilds;
};
__device__ __forceinline__ int sp_cell_of(float v, float inv_cs) { return (int)floorf(v * inv_cs); }
__device__ __forceinline__ unsigned int sp_hash(int cx, int cy, int cz, int hbits) {
    const unsigned int h = (unsigned)cx * 73856093u ^ (unsigned)cy * 19349663u ^ (unsigned)cz * 83492791u;
    return (h ^ (h >> 15)) & ((1u << hbits) - 1u);
}
struct SpCsr { const long long* row_ptr; const int* col; const float* dist; };

// One force evaluation = ONE cooperative launch (grid barriers between its phases):
//   A  x = a + lam * b for every bead; a bead that has moved more than skin / 2 since the last build raises the rebuild flag
//   B  (only when the flag is up — the decision is made on the device, identically in every CTA) rebuild the bins: count (a
//      histogram: independent of the order of the atomics), two-level exclusive scan over the grid, place, sort every bucket by
//      bead index
//   C  a warp per bead: lanes stride over the bead's CSR row (NOE), lanes 0..26 scan one neighbour cell each (repel), the
//      butterfly adds the lane partials in a fixed order, lane 0 adds the bond term and writes the total force
//   D  energies: per-CTA partials, summed in CTA order by the CTA that finishes last (a ticket; no barrier)
// The result is the total force per bead and the three energies — what the integrator / minimiser kernels of large_n.cu
// (ln_md_step_kernel, ln_min_step_kernel: the `fsum` path) consume.
struct SpEvalArgs {
    const float4* a; const float4* b; const float* lam; int N; int force_rebuild;
    float4* xs; SpGrid g; SpCsr m; LnTerms t; LnBonded bd; int vdw_min_sep;
    float4* fsum; double* epart; double* escal; unsigned int* ticket;
};
template <bool LOW_SOFT>
__global__ void __launch_bounds__(SP_THREADS) sp_eval_kernel(const SpEvalArgs A) {
    cooperative_groups::grid_group grid = cooperative_groups::this_grid();
    const int N = A.N, tid = threadIdx.x, lane = tid & 31, gtid = blockIdx.x * SP_THREADS + tid, gsize = gridDim.x * SP_THREADS;
    const SpGrid g = A.g;
    const float4* xs = A.xs;             // written in phase A of this launch: plain loads, never the read-only path
    __shared__ double sh[3][SP_THREADS / 32];
    __shared__ int sh_scan[SP_THREADS];
    __shared__ int sh_last;
    {   // ---- A
        const float lam = A.b ? *A.lam : 0.f;
        const float lim2 = 0.25f * SP_SKIN * SP_SKIN;
        int moved = 0;
        for (int i = gtid; i < N; i += gsize) {
            const float4 p = A.a[i];
            float4 x = p;
            if (A.b) { const float4 q = A.b[i]; x = make_float4(fmaf(lam, q.x, p.x), fmaf(lam, q.y, p.y), fmaf(lam, q.z, p.z), 0.f); }
            A.xs[i] = x;
            const float4 r = g.xref[i];
            const float dx = x.x - r.x, dy = x.y - r.y, dz = x.z - r.z;
            if (dx * dx + dy * dy + dz * dz > lim2) moved = 1;
        }
        if (__syncthreads_or(moved | A.force_rebuild) && tid == 0) *(volatile int*)g.rebuild = 1;       // (the flag is order-independent)
    }
    grid.sync();
    if (*(volatile int*)g.rebuild) {      // ---- B (uniform over the grid: the flag is not touched again before the next barrier)
        const int H = 1 << g.hbits;
        const int seg = (H + (int)gridDim.x - 1) / (int)gridDim.x, k0 = min(H, (int)blockIdx.x * seg), k1 = min(H, k0 + seg);
        for (int i = gtid; i < N; i += gsize) {          // beads per bucket (cell_count / cell_fill are zero on entry: cleared below and at creation)
            const float4 x = xs[i];
            atomicAdd(&g.cell_count[sp_hash(sp_cell_of(x.x, g.inv_cs), sp_cell_of(x.y, g.inv_cs), sp_cell_of(x.z, g.inv_cs), g.hbits)], 1);
            g.xref[i] = x;
        }
        grid.sync();
        if (gtid == 0) { *(volatile int*)g.rebuild = 0; *g.n_rebuilds += 1; }
        {   // exclusive scan of the counts, level 1: the total of this CTA's segment of buckets
            int s = 0;
            for (int k = k0 + tid; k < k1; k += SP_THREADS) s += g.cell_count[k];
            s = __reduce_add_sync(0xffffffffu, s);
            if (lane == 0) sh_scan[tid >> 5] = s;
            __syncthreads();
            if (tid == 0) { int v = 0; for (int w = 0; w < SP_THREADS / 32; w++) v += sh_scan[w]; g.bsum[blockIdx.x] = v; }
        }
        grid.sync();
        {   // level 2: offset of the segment = total of the segments before it, then the segment itself in chunks of a CTA
            int off = 0;
            for (int c = tid; c < (int)blockIdx.x; c += SP_THREADS) off += __ldcg(g.bsum + c);
            off = __reduce_add_sync(0xffffffffu, off);
            __syncthreads();
            if (lane == 0) sh_scan[tid >> 5] = off;
            __syncthreads();
            int run = 0;
            for (int w = 0; w < SP_THREADS / 32; w++) run += sh_scan[w];
            for (int c0 = k0; c0 < k1; c0 += SP_THREADS) {
                const int k = c0 + tid;
                const int v = k < k1 ? g.cell_count[k] : 0;
                int inc = v;                                                  // inclusive scan inside the warp, then across the warps
                for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += u; }
                __syncthreads();
                if (lane == 31) sh_scan[tid >> 5] = inc;
                __syncthreads();
                int wo = 0, tot = 0;
                for (int w = 0; w < SP_THREADS / 32; w++) { const int u = sh_scan[w]; if (w < (tid >> 5)) wo += u; tot += u; }
                if (k < k1) g.cell_start[k] = run + wo + inc - v;
                run += tot;
            }
            if (blockIdx.x == gridDim.x - 1 && tid == 0) g.cell_start[H] = run;
        }
        grid.sync();
        for (int i = gtid; i < N; i += gsize) {          // place (any order inside a bucket) ...
            const float4 x = xs[i];
            const unsigned int h = sp_hash(sp_cell_of(x.x, g.inv_cs), sp_cell_of(x.y, g.inv_cs), sp_cell_of(x.z, g.inv_cs), g.hbits);
            g.order[g.cell_start[h] + atomicAdd(&g.cell_fill[h], 1)] = i;
        }
        for (int k = gtid; k < H; k += gsize) g.cell_count[k] = 0;           // (the counts are not read any more)
        grid.sync();
        for (int k = gtid; k < H; k += gsize) {          // ... and sort every bucket by bead index (short lists): the scan order of the repel term is fixed
            const int s = g.cell_start[k], e = g.cell_start[k + 1];
            for (int a = s + 1; a < e; a++) {
                const int v = g.order[a]; int b = a - 1;
                while (b >= s && g.order[b] > v) { g.order[b + 1] = g.order[b]; b--; }
                g.order[b + 1] = v;
            }
            for (int a = s; a < e; a++) {                                    // the scan of the repel term reads one 16-byte entry per candidate
                const int j = g.order[a]; const float4 x = xs[j];
                g.ent[a] = make_int4(j, sp_cell_of(x.x, g.inv_cs), sp_cell_of(x.y, g.inv_cs), sp_cell_of(x.z, g.inv_cs));
            }
            g.cell_fill[k] = 0;
        }
        grid.sync();
    }
    // ---- C
    const int warps = gsize >> 5;
    const LnTerms t = A.t;
    double e_noe = 0.0, e_vdw = 0.0, e_bond = 0.0;          // lane 0: this warp's beads, in bead order
    for (int i = gtid >> 5; i < N; i += warps) {
        const float4 xi = xs[i];
        float fx = 0.f, fy = 0.f, fz = 0.f, en = 0.f, ev = 0.f;
        const long long r0 = A.m.row_ptr[i], r1 = A.m.row_ptr[i + 1];
#pragma unroll 4
        for (long long k = r0 + lane; k < r1; k += 32) {      // (unrolled: four column -> coordinate load chains in flight per lane)
            const float4 xj = xs[__ldg(A.m.col + k)];
            const float d = __ldg(A.m.dist + k);
            const float dx = xi.x - xj.x, dy = xi.y - xj.y, dz = xi.z - xj.z;
            const float r2 = fmaf(dx, dx, fmaf(dy, dy, fmaf(dz, dz, 1e-12f))), rinv = ln_rsqrt(r2);
            const float delta = fmaf(r2, rinv, -d);
            float h = fminf(delta, t.rsw);
            if (LOW_SOFT) h = fmaxf(h, t.lo);
            const float w2h = t.wnC2m * h;                                   // -2 w C h
            en = fmaf(-0.5f * w2h, 2.f * delta - h, en);                     // w C h (2 delta - h)
            const float coef = w2h * rinv;                                   // -(dE/dr)/r
            fx = fmaf(coef, dx, fx); fy = fmaf(coef, dy, fy); fz = fmaf(coef, dz, fz);
        }
        if (t.wv != 0.f && lane < 27) {
            // centre = the cell bead i was binned in: two beads within R now were within R + skin = one cell edge at the last build
            const float4 ri = g.xref[i];
            const int nx = sp_cell_of(ri.x, g.inv_cs) + lane % 3 - 1, ny = sp_cell_of(ri.y, g.inv_cs) + (lane / 3) % 3 - 1, nz = sp_cell_of(ri.z, g.inv_cs) + lane / 9 - 1;
            const unsigned int h = sp_hash(nx, ny, nz, g.hbits);
            const int k1 = g.cell_start[h + 1];
            // the bucket may hold beads of other cells (hash collisions), and two of the 27 cells may share a bucket: a bead counts
            // only in the pass of the cell it was binned in.  Two candidates in flight (their loads do not depend on each other);
            // the sums still run in bucket order.
            for (int k = g.cell_start[h]; k < k1; k += 2) {
                const int4 c0 = g.ent[k];
                const bool two = k + 1 < k1;
                const int4 c1 = two ? g.ent[k + 1] : c0;
                const float4 x0 = xs[c0.x], x1 = xs[c1.x];
                const int s0 = c0.x > i ? c0.x - i : i - c0.x, s1 = c1.x > i ? c1.x - i : i - c1.x;
                const bool ok0 = s0 >= A.vdw_min_sep && c0.y == nx && c0.z == ny && c0.w == nz;
                const bool ok1 = two && s1 >= A.vdw_min_sep && c1.y == nx && c1.z == ny && c1.w == nz;
                if (ok0) {
                    const float ddx = xi.x - x0.x, ddy = xi.y - x0.y, ddz = xi.z - x0.z;
                    const float tr = fmaxf(t.R2 - fmaf(ddx, ddx, fmaf(ddy, ddy, ddz * ddz)), 0.f);
                    const float coef = t.wv4 * tr;                           // 4 w_vdw t = -(dE/dr)/r
                    fx = fmaf(coef, ddx, fx); fy = fmaf(coef, ddy, fy); fz = fmaf(coef, ddz, fz);
                    ev = fmaf(t.wv * tr, tr, ev);
                }
                if (ok1) {
                    const float ddx = xi.x - x1.x, ddy = xi.y - x1.y, ddz = xi.z - x1.z;
                    const float tr = fmaxf(t.R2 - fmaf(ddx, ddx, fmaf(ddy, ddy, ddz * ddz)), 0.f);
                    const float coef = t.wv4 * tr;
                    fx = fmaf(coef, ddx, fx); fy = fmaf(coef, ddy, fy); fz = fmaf(coef, ddz, fz);
                    ev = fmaf(t.wv * tr, tr, ev);
                }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            fx += __shfl_xor_sync(0xffffffffu, fx, o); fy += __shfl_xor_sync(0xffffffffu, fy, o); fz += __shfl_xor_sync(0xffffffffu, fz, o);
            en += __shfl_xor_sync(0xffffffffu, en, o); ev += __shfl_xor_sync(0xffffffffu, ev, o);
        }
        if (lane == 0) {
            float eb = 0.f;
            ln_bond(xs, xs, 0.f, i, N, A.bd, fx, fy, fz, eb);
            A.fsum[i] = make_float4(fx, fy, fz, 0.f);
            e_noe += 0.5 * (double)en; e_vdw += 0.5 * (double)ev; e_bond += (double)eb;
        }
    }
    // ---- D: CTA partials in warp order, grid total in CTA order by the last CTA to finish
    if (lane == 0) { sh[0][tid >> 5] = e_noe; sh[1][tid >> 5] = e_vdw; sh[2][tid >> 5] = e_bond; }
    __syncthreads();
    if (tid < 3) { double s = 0; for (int w = 0; w < SP_THREADS / 32; w++) s += sh[tid][w]; A.epart[(size_t)blockIdx.x * 3 + tid] = s; }
    __threadfence();
    __syncthreads();
    if (tid == 0) sh_last = atomicAdd(A.ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (sh_last) {
        __threadfence();
        if (tid < 3) {
            double s = 0;
            for (int b = 0; b < (int)gridDim.x; b++) s += __ldcg(A.epart + (size_t)b * 3 + tid);
            A.escal[tid] = s;          // [0] e_noe, [1] e_vdw, [2] e_bond: the layout the step kernels read
        }
        if (tid == 0) *A.ticket = 0u;
    }
}
