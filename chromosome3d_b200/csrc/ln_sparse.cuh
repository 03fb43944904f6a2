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
//     build (skin = the reference's tolerance, 0.5 A), and every bead scans the 27 cells around it.  Beads inside a bin are kept
//     in ascending index order, so the summation order is fixed here too.
//   * bond (+ angle stand-in) term and the three energies: one combine kernel.
// The result is the total force per bead and the three energies — exactly what the multi-launch integrator / minimiser kernels of
// large_n.cu (ln_md_step_kernel, ln_min_step_kernel: the `fsum` path) consume.

#define SP_THREADS 256
#define SP_SKIN 0.5f                    // nbonds tolerance=0.5 (C3D:1402, 1736)

struct SpGrid {                         // hashed cell grid
    float inv_cs; int hbits;            // 1 / cell edge, log2(buckets)
    int* cell_start;                    // [H + 1]
    int* cell_count;                    // [H]
    int* cell_fill;                     // [H] placement cursors
    int* order;                         // [N] beads sorted by (bucket, index)
    float4* xref;                       // [N] coordinates at the last build
    int* rebuild;                       // device flag: a bead has moved more than skin / 2
    long long* n_rebuilds;
};
__device__ __forceinline__ int sp_cell_of(float v, float inv_cs) { return (int)floorf(v * inv_cs); }
__device__ __forceinline__ unsigned int sp_hash(int cx, int cy, int cz, int hbits) {
    const unsigned int h = (unsigned)cx * 73856093u ^ (unsigned)cy * 19349663u ^ (unsigned)cz * 83492791u;
    return (h ^ (h >> 15)) & ((1u << hbits) - 1u);
}

// x = a + lam * b for every bead; the largest displacement since the last build decides whether the bins are rebuilt
__global__ void __launch_bounds__(SP_THREADS) sp_coords_kernel(const float4* __restrict__ a, const float4* __restrict__ b, const float* lam_p, int N, float4* __restrict__ xs,
                                                                SpGrid g, int force_rebuild) {
    const float lam = b ? *lam_p : 0.f;
    const float lim2 = 0.25f * SP_SKIN * SP_SKIN;
    int moved = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
        const float4 p = a[i];
        float4 x = p;
        if (b) { const float4 q = b[i]; x = make_float4(fmaf(lam, q.x, p.x), fmaf(lam, q.y, p.y), fmaf(lam, q.z, p.z), 0.f); }
        xs[i] = x;
        const float4 r = g.xref[i];
        const float dx = x.x - r.x, dy = x.y - r.y, dz = x.z - r.z;
        if (dx * dx + dy * dy + dz * dz > lim2) moved = 1;
    }
    if (__syncthreads_or(moved | force_rebuild) && threadIdx.x == 0) *g.rebuild = 1;       // (the flag is order-independent)
}
// rebuild, step 1: clear the counters
__global__ void sp_grid_clear_kernel(SpGrid g) {
    if (!*g.rebuild) return;
    const int H = 1 << g.hbits;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < H; k += gridDim.x * blockDim.x) { g.cell_count[k] = 0; g.cell_fill[k] = 0; }
}
// step 2: beads per bucket (a histogram: the result does not depend on the order of the atomics)
__global__ void sp_grid_count_kernel(const float4* __restrict__ xs, int N, SpGrid g) {
    if (!*g.rebuild) return;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
        const float4 x = xs[i];
        atomicAdd(&g.cell_count[sp_hash(sp_cell_of(x.x, g.inv_cs), sp_cell_of(x.y, g.inv_cs), sp_cell_of(x.z, g.inv_cs), g.hbits)], 1);
        g.xref[i] = x;
    }
}
// step 3: exclusive scan of the counts (one CTA, H <= 2^17)
__global__ void __launch_bounds__(1024) sp_grid_scan_kernel(SpGrid g) {
    if (!*g.rebuild) return;
    __shared__ int part[1024];
    const int H = 1 << g.hbits, per = (H + 1023) / 1024, t = threadIdx.x;
    int s = 0;
    for (int k = t * per; k < min(H, (t + 1) * per); k++) s += g.cell_count[k];
    part[t] = s;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) { int v = t >= o ? part[t - o] : 0; __syncthreads(); part[t] += v; __syncthreads(); }
    int run = part[t] - s;
    for (int k = t * per; k < min(H, (t + 1) * per); k++) { g.cell_start[k] = run; run += g.cell_count[k]; }
    if (t == 1023) g.cell_start[H] = part[1023];
    if (t == 0) *g.n_rebuilds += 1;
}
// step 4: place the beads (any order inside a bucket) ...
__global__ void sp_grid_place_kernel(const float4* __restrict__ xs, int N, SpGrid g) {
    if (!*g.rebuild) return;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
        const float4 x = xs[i];
        const unsigned int h = sp_hash(sp_cell_of(x.x, g.inv_cs), sp_cell_of(x.y, g.inv_cs), sp_cell_of(x.z, g.inv_cs), g.hbits);
        g.order[g.cell_start[h] + atomicAdd(&g.cell_fill[h], 1)] = i;
    }
}
// ... step 5: and sort every bucket by bead index (short lists: insertion sort), so that the scan order of the repel term is fixed;
// the last block clears the flag
__global__ void sp_grid_sort_kernel(SpGrid g, unsigned int* done_counter) {
    if (!*g.rebuild) return;
    const int H = 1 << g.hbits;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < H; k += gridDim.x * blockDim.x) {
        const int s = g.cell_start[k], e = g.cell_start[k + 1];
        for (int a = s + 1; a < e; a++) {
            const int v = g.order[a]; int b = a - 1;
            while (b >= s && g.order[b] > v) { g.order[b + 1] = g.order[b]; b--; }
            g.order[b + 1] = v;
        }
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0 && atomicAdd(done_counter, 1u) == gridDim.x - 1) { *done_counter = 0; *g.rebuild = 0; }
}

struct SpCsr { const long long* row_ptr; const int* col; const float* dist; };
// NOE term of the restraint rows: one warp per bead, 8 B per entry.  f_noe[i] = force on bead i, e_row[i] = half the energy of its entries.
template <bool LOW_SOFT>
__global__ void __launch_bounds__(SP_THREADS) sp_noe_kernel(SpCsr m, const float4* __restrict__ xs, int N, LnTerms t, float4* __restrict__ f_noe, double* __restrict__ e_row) {
    const int lane = threadIdx.x & 31, warps = (gridDim.x * blockDim.x) >> 5;
    for (int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < N; i += warps) {
        const float4 xi = xs[i];
        float fx = 0.f, fy = 0.f, fz = 0.f, en = 0.f;
        const long long r0 = m.row_ptr[i], r1 = m.row_ptr[i + 1];
        for (long long k = r0 + lane; k < r1; k += 32) {
            const float4 xj = __ldg(xs + __ldg(m.col + k));
            const float d = __ldg(m.dist + k);
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
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            fx += __shfl_xor_sync(0xffffffffu, fx, o); fy += __shfl_xor_sync(0xffffffffu, fy, o); fz += __shfl_xor_sync(0xffffffffu, fz, o);
            en += __shfl_xor_sync(0xffffffffu, en, o);
        }
        if (lane == 0) { f_noe[i] = make_float4(fx, fy, fz, 0.f); e_row[i] = 0.5 * (double)en; }
    }
}
// repel term from the cell list (a thread per bead, the 27 cells around it in a fixed order) + bond term + combine:
// fsum[i] = total force, e3[i] = {NOE half-row energy (copied), half the repel energy of the bead's pairs, bond energy (i, i+1)}
__global__ void __launch_bounds__(SP_THREADS) sp_repel_combine_kernel(const float4* __restrict__ xs, int N, SpGrid g, LnTerms t, LnBonded bd, int vdw_min_sep,
                                                                       const float4* __restrict__ f_noe, const double* __restrict__ e_row,
                                                                       float4* __restrict__ fsum, double* __restrict__ epart /* [blocks][3] */) {
    __shared__ double sh[3][SP_THREADS / 32];
    double e_noe = 0.0, e_vdw = 0.0, e_bond = 0.0;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < N) {
        const float4 xi = xs[i];
        float4 f = f_noe[i];
        float ev = 0.f;
        if (t.wv != 0.f) {
            // centre = the cell bead i was binned in: two beads within R now were within R + skin = one cell edge at the last build
            const float4 ri = g.xref[i];
            const int cx = sp_cell_of(ri.x, g.inv_cs), cy = sp_cell_of(ri.y, g.inv_cs), cz = sp_cell_of(ri.z, g.inv_cs);
            for (int dz = -1; dz <= 1; dz++) for (int dy = -1; dy <= 1; dy++) for (int dx = -1; dx <= 1; dx++) {
                const int nx = cx + dx, ny = cy + dy, nz = cz + dz;
                const unsigned int h = sp_hash(nx, ny, nz, g.hbits);
                for (int k = g.cell_start[h]; k < g.cell_start[h + 1]; k++) {
                    const int j = g.order[k];
                    const int sep = j > i ? j - i : i - j;
                    if (sep < vdw_min_sep) continue;
                    const float4 xj = xs[j];
                    // the bucket may hold beads of other cells (hash collisions), and two of the 27 cells may share a bucket:
                    // a bead counts only in the pass of the cell it was binned in (g.xref: its coordinates at the last build)
                    const float4 rj = g.xref[j];
                    if (sp_cell_of(rj.x, g.inv_cs) != nx || sp_cell_of(rj.y, g.inv_cs) != ny || sp_cell_of(rj.z, g.inv_cs) != nz) continue;
                    const float ddx = xi.x - xj.x, ddy = xi.y - xj.y, ddz = xi.z - xj.z;
                    const float r2 = fmaf(ddx, ddx, fmaf(ddy, ddy, ddz * ddz));
                    const float tr = fmaxf(t.R2 - r2, 0.f);
                    const float coef = t.wv4 * tr;                           // 4 w_vdw t = -(dE/dr)/r
                    f.x = fmaf(coef, ddx, f.x); f.y = fmaf(coef, ddy, f.y); f.z = fmaf(coef, ddz, f.z);
                    ev = fmaf(t.wv * tr, tr, ev);
                }
            }
        }
        float fx = f.x, fy = f.y, fz = f.z, eb = 0.f;
        ln_bond(xs, xs, 0.f, i, N, bd, fx, fy, fz, eb);
        fsum[i] = make_float4(fx, fy, fz, 0.f);
        e_noe = e_row[i]; e_vdw = 0.5 * (double)ev; e_bond = (double)eb;
    }
    // block sums in a fixed order
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { e_noe += __shfl_xor_sync(0xffffffffu, e_noe, o); e_vdw += __shfl_xor_sync(0xffffffffu, e_vdw, o); e_bond += __shfl_xor_sync(0xffffffffu, e_bond, o); }
    if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = e_noe; sh[1][threadIdx.x >> 5] = e_vdw; sh[2][threadIdx.x >> 5] = e_bond; }
    __syncthreads();
    if (threadIdx.x < 3) { double s = 0; for (int w = 0; w < SP_THREADS / 32; w++) s += sh[threadIdx.x][w]; epart[(size_t)blockIdx.x * 3 + threadIdx.x] = s; }
}
__global__ void sp_energy_kernel(const double* __restrict__ epart, int nblocks, double* __restrict__ escal) {
    if (threadIdx.x >= 3 || blockIdx.x != 0) return;
    double s = 0; for (int b = 0; b < nblocks; b++) s += epart[(size_t)b * 3 + threadIdx.x];
    escal[threadIdx.x] = s;          // [0] e_noe, [1] e_vdw, [2] e_bond: the layout the step kernels read
}
