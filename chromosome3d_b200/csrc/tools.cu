// tools.cu — the callers and data formats either side of the hot path (SURVEY 8(f) "next" rows 3 and 4):
//   * the cross-resolution tools whose outputs ship under output_models/ (their programs were never published):
//     pair-averaged "_reduced.pdb" models and the similarity.txt scorer (formats: SURVEY App. B.6);
//   * matrix input at scale: a threaded parser for the reference's text matrices (C3D:116-129), a raw binary
//     container, and sparse (bin_i, bin_j, count) triplets scattered into the dense symmetric matrix on the device.
#include <charconv>
#include <stdlib.h>
#include <string.h>
#include <atomic>
#include <thread>
#include "common.h"

namespace chr3d {

__global__ void reduce_model_kernel(const double* __restrict__ x, int N, int m, double* __restrict__ out) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 3 * m) return;
    const int k = t / 3, q = t % 3;
    const double a = x[3 * (2 * k) + q];
    out[t] = (2 * k + 1 < N) ? __ddiv_rn(__dadd_rn(a, x[3 * (2 * k + 1) + q]), 2.0) : a;
}
__global__ void coo_scatter_kernel(const int32_t* __restrict__ bi, const int32_t* __restrict__ bj, const double* __restrict__ v, long long nnz,
                                   int L, int base, double* __restrict__ out, int* __restrict__ bad) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nnz) return;
    const int i = bi[k] - base, j = bj[k] - base;
    if (i < 0 || j < 0 || i >= L || j >= L) { *bad = 1; return; }
    atomicAdd(out + (size_t)i * L + j, v[k]);
    if (i != j) atomicAdd(out + (size_t)j * L + i, v[k]);
}

// "%.3f" as the tool that wrote the shipped _reduced.pdb files printed it (java.util.Formatter): the shortest decimal
// digits that round-trip the double, then HALF_UP at the third decimal.  Coordinates that are means of two 3-decimal
// values end in ...5 exactly in that representation, so they round away from zero, unlike C's printf.
static std::string fmt3_half_up(double v) {
    char buf[40];
    const double a = v < 0 ? -v : v;
    auto r = std::to_chars(buf, buf + sizeof buf, a, std::chars_format::scientific);      // shortest round-trip, d[.ddd]e[+-]XX
    std::string t(buf, r.ptr);
    const size_t e = t.find('e');
    std::string dig;
    for (size_t i = 0; i < e; i++) if (t[i] != '.') dig.push_back(t[i]);
    const int ex = atoi(t.c_str() + e + 1);
    // milli-units: digits * 10^(ex - (nd-1) + 3)
    const int shift = ex - ((int)dig.size() - 1) + 3;
    unsigned long long milli = 0;
    if (shift >= 0) { for (char c : dig) milli = milli * 10 + (unsigned)(c - '0'); for (int i = 0; i < shift; i++) milli *= 10; }
    else {
        const int keep = (int)dig.size() + shift;
        for (int i = 0; i < keep; i++) milli = milli * 10 + (unsigned)(dig[i] - '0');
        if (keep >= 0 && dig[keep] >= '5') milli++;
    }
    char out[48];
    snprintf(out, sizeof out, "%s%llu.%03llu", std::signbit(v) ? "-" : "", milli / 1000, milli % 1000);
    return out;
}

static int slurp(const char* path, std::string& s) {
    FILE* f = fopen(path, "rb");
    if (!f) { set_error("cannot open %s", path); return CHR3D_ERR_IO; }
    fseek(f, 0, SEEK_END); long long n = ftell(f); fseek(f, 0, SEEK_SET);
    s.resize((size_t)n);
    if (n > 0 && fread(&s[0], 1, (size_t)n, f) != (size_t)n) { fclose(f); set_error("short read on %s", path); return CHR3D_ERR_IO; }
    fclose(f);
    return 0;
}
static inline bool is_ws(char c) { return c == ' ' || c == '\t' || c == '\r' || c == '\n' || c == '\f' || c == '\v'; }

}  // namespace chr3d

using namespace chr3d;

// ------------------------------ 8(f) rank 3: cross-resolution tools ------------------------------
CHR3D_API int chr3d_reduce_model(const double* xyz, int32_t N, double* out_xyz) {
    if (!xyz || !out_xyz || N < 1) { set_error("chr3d_reduce_model: bad arguments"); return CHR3D_ERR_ARG; }
    if (int rc = require_device()) return rc;
    const int m = (N + 1) / 2;
    DevBuf<double> dx, dout;
    CHR3D_CUDA(dx.alloc((size_t)N * 3)); CHR3D_CUDA(dout.alloc((size_t)m * 3));
    CHR3D_CUDA(cudaMemcpy(dx.p, xyz, sizeof(double) * (size_t)N * 3, cudaMemcpyHostToDevice));
    reduce_model_kernel<<<(3 * m + 255) / 256, 256>>>(dx.p, N, m, dout.p);
    CHR3D_CUDA(cudaGetLastError());
    CHR3D_CUDA(cudaMemcpy(out_xyz, dout.p, sizeof(double) * (size_t)m * 3, cudaMemcpyDeviceToHost));
    return 0;
}
CHR3D_API int chr3d_write_reduced_pdb(const char* path, const double* xyz, int32_t n) {
    if (!path || !xyz || n < 1) { set_error("chr3d_write_reduced_pdb: bad arguments"); return CHR3D_ERR_ARG; }
    std::string s = "\r\n"; char b[160];
    for (int k = 0; k < n; k++) {
        snprintf(b, sizeof b, "ATOM  %5d   CA MET B%-8d%8s%8s%8s  0.20 10.00\r\n", k + 1, k + 1, fmt3_half_up(xyz[3 * k]).c_str(),
                 fmt3_half_up(xyz[3 * k + 1]).c_str(), fmt3_half_up(xyz[3 * k + 2]).c_str());
        s += b;
    }
    for (int k = 1; k < n; k++) { snprintf(b, sizeof b, "CONECT%5d%5d\r\n", k, k + 1); s += b; }
    s += "END\r\n";
    FILE* f = fopen(path, "wb");
    if (!f) { set_error("cannot open %s for writing", path); return CHR3D_ERR_IO; }
    const bool ok = fwrite(s.data(), 1, s.size(), f) == s.size();
    fclose(f);
    if (!ok) { set_error("short write on %s", path); return CHR3D_ERR_IO; }
    return 0;
}
CHR3D_API int chr3d_model_similarity(const double* xyz_a, int32_t n_a, const double* xyz_b, int32_t n_b, double* spearman, double* rmsd) {
    if (!xyz_a || !xyz_b || !spearman || !rmsd || n_a < 1 || n_b < 1) { set_error("chr3d_model_similarity: bad arguments"); return CHR3D_ERR_ARG; }
    if (int rc = require_device()) return rc;
    const int n = n_a < n_b ? n_a : n_b;
    DevBuf<double> da, db;
    CHR3D_CUDA(da.alloc((size_t)n * 3)); CHR3D_CUDA(db.alloc((size_t)n * 3));
    CHR3D_CUDA(cudaMemcpy(da.p, xyz_a, sizeof(double) * (size_t)n * 3, cudaMemcpyHostToDevice));
    CHR3D_CUDA(cudaMemcpy(db.p, xyz_b, sizeof(double) * (size_t)n * 3, cudaMemcpyHostToDevice));
    return dev_model_similarity(da.p, db.p, n, spearman, rmsd, 0);
}

// ------------------------------ 8(f) rank 4: matrix input at scale -------------------------------
// Text matrix as IF2dist_new reads it (C3D:116-129): one row per line, tokens split on whitespace (which eats the CR of
// the CRLF line ends); L = token count of the first non-blank row (calc_len_IF, C3D:164-179); short rows are padded
// with 0, extra tokens and extra rows are ignored.  Two-call protocol: cap == 0 returns only L.  Rows are parsed by all
// host threads (strtod per token: correctly rounded, what Perl's numeric conversion gives for these inputs).
CHR3D_API int chr3d_read_if_text(const char* path, double* out, int64_t cap, int32_t* L_out) {
    if (!path || !L_out) { set_error("chr3d_read_if_text: bad arguments"); return CHR3D_ERR_ARG; }
    std::string s;
    if (int rc = slurp(path, s)) return rc;
    std::vector<size_t> rows;                     // start offsets of the non-blank lines
    const size_t n = s.size();
    for (size_t p = 0; p < n;) {
        size_t e = p; while (e < n && s[e] != '\n') e++;
        size_t q = p; while (q < e && is_ws(s[q])) q++;
        if (q < e) rows.push_back(p);
        p = e + 1;
    }
    if (rows.empty()) { set_error("%s: empty matrix", path); return CHR3D_ERR_IO; }
    int L = 0;
    {
        size_t p = rows[0]; bool in = false;
        for (; p < n && s[p] != '\n'; p++) { const bool w = is_ws(s[p]); if (!w && !in) L++; in = !w; }
    }
    *L_out = L;
    if (cap == 0 || !out) return 0;
    if (cap < (int64_t)L * L) { set_error("chr3d_read_if_text: buffer of %lld doubles for L=%d", (long long)cap, L); return CHR3D_ERR_CAPACITY; }
    if ((int)rows.size() < L) { set_error("%s: %zu rows for %d columns", path, rows.size(), L); return CHR3D_ERR_IO; }
    s.push_back('\n');                            // sentinel: strtod never runs off the buffer
    std::atomic<int> next(0), bad(0);
    auto work = [&]() {
        for (int r; (r = next.fetch_add(1)) < L;) {
            const char* p = s.data() + rows[r];
            double* o = out + (size_t)r * L;
            int c = 0;
            while (c < L) {
                while (*p != '\n' && is_ws(*p)) p++;
                if (*p == '\n') break;
                char* e = nullptr;
                o[c++] = strtod(p, &e);
                if (e == p) { bad = 1; while (*p != '\n' && !is_ws(*p)) p++; } else p = e;
            }
            for (; c < L; c++) o[c] = 0.0;
        }
    };
    unsigned nt = std::thread::hardware_concurrency(); if (nt == 0) nt = 1; if (nt > 64) nt = 64; if ((int)nt > L) nt = (unsigned)L;
    std::vector<std::thread> th;
    for (unsigned t = 1; t < nt; t++) th.emplace_back(work);
    work();
    for (auto& t : th) t.join();
    if (bad) { set_error("%s: non-numeric token", path); return CHR3D_ERR_IO; }
    return 0;
}
// raw container for matrices too large for decimal text (25k^2 tokens = 10+ GB): "CHR3DIF1", int32 L, int32 element
// bytes (8: fp64, 4: fp32), then L*L little-endian values in row-major order
CHR3D_API int chr3d_write_if_binary(const char* path, const double* IF, int32_t L, int32_t elem_bytes) {
    if (!path || !IF || L < 1 || (elem_bytes != 4 && elem_bytes != 8)) { set_error("chr3d_write_if_binary: bad arguments"); return CHR3D_ERR_ARG; }
    FILE* f = fopen(path, "wb");
    if (!f) { set_error("cannot open %s for writing", path); return CHR3D_ERR_IO; }
    bool ok = fwrite("CHR3DIF1", 1, 8, f) == 8 && fwrite(&L, 4, 1, f) == 1 && fwrite(&elem_bytes, 4, 1, f) == 1;
    const size_t n = (size_t)L * L;
    if (ok && elem_bytes == 8) ok = fwrite(IF, 8, n, f) == n;
    else if (ok) {
        std::vector<float> row((size_t)L);
        for (int r = 0; r < L && ok; r++) { for (int c = 0; c < L; c++) row[c] = (float)IF[(size_t)r * L + c]; ok = fwrite(row.data(), 4, (size_t)L, f) == (size_t)L; }
    }
    fclose(f);
    if (!ok) { set_error("short write on %s", path); return CHR3D_ERR_IO; }
    return 0;
}
CHR3D_API int chr3d_read_if_binary(const char* path, double* out, int64_t cap, int32_t* L_out) {
    if (!path || !L_out) { set_error("chr3d_read_if_binary: bad arguments"); return CHR3D_ERR_ARG; }
    FILE* f = fopen(path, "rb");
    if (!f) { set_error("cannot open %s", path); return CHR3D_ERR_IO; }
    char magic[8]; int32_t L = 0, eb = 0;
    if (fread(magic, 1, 8, f) != 8 || memcmp(magic, "CHR3DIF1", 8) != 0 || fread(&L, 4, 1, f) != 1 || fread(&eb, 4, 1, f) != 1 || L < 1 || (eb != 4 && eb != 8)) {
        fclose(f); set_error("%s: not a CHR3DIF1 matrix", path); return CHR3D_ERR_IO;
    }
    *L_out = L;
    if (cap == 0 || !out) { fclose(f); return 0; }
    if (cap < (int64_t)L * L) { fclose(f); set_error("chr3d_read_if_binary: buffer of %lld doubles for L=%d", (long long)cap, L); return CHR3D_ERR_CAPACITY; }
    const size_t n = (size_t)L * L;
    bool ok = true;
    if (eb == 8) ok = fread(out, 8, n, f) == n;
    else {
        std::vector<float> row((size_t)L);
        for (int r = 0; r < L && ok; r++) { ok = fread(row.data(), 4, (size_t)L, f) == (size_t)L; for (int c = 0; c < L; c++) out[(size_t)r * L + c] = (double)row[c]; }
    }
    fclose(f);
    if (!ok) { set_error("%s: truncated", path); return CHR3D_ERR_IO; }
    return 0;
}
// sparse upper-triangle triplets (bin_i, bin_j, value), `base` = 0 or 1, -> dense symmetric L x L matrix (scatter on the
// device; cells that do not appear are 0, values of repeated triplets add up)
CHR3D_API int chr3d_if_from_coo(const int32_t* bi, const int32_t* bj, const double* v, int64_t nnz, int32_t L, int32_t base, double* out) {
    if (!out || L < 1 || nnz < 0 || (nnz > 0 && (!bi || !bj || !v)) || (base != 0 && base != 1)) { set_error("chr3d_if_from_coo: bad arguments"); return CHR3D_ERR_ARG; }
    if (int rc = require_device()) return rc;
    DevBuf<int32_t> di, dj; DevBuf<double> dv, dout; DevBuf<int> dbad;
    CHR3D_CUDA(dout.alloc((size_t)L * L)); CHR3D_CUDA(dbad.alloc(1));
    CHR3D_CUDA(cudaMemset(dout.p, 0, sizeof(double) * (size_t)L * L)); CHR3D_CUDA(cudaMemset(dbad.p, 0, sizeof(int)));
    if (nnz > 0) {
        CHR3D_CUDA(di.alloc((size_t)nnz)); CHR3D_CUDA(dj.alloc((size_t)nnz)); CHR3D_CUDA(dv.alloc((size_t)nnz));
        CHR3D_CUDA(cudaMemcpy(di.p, bi, sizeof(int32_t) * (size_t)nnz, cudaMemcpyHostToDevice));
        CHR3D_CUDA(cudaMemcpy(dj.p, bj, sizeof(int32_t) * (size_t)nnz, cudaMemcpyHostToDevice));
        CHR3D_CUDA(cudaMemcpy(dv.p, v, sizeof(double) * (size_t)nnz, cudaMemcpyHostToDevice));
        coo_scatter_kernel<<<(unsigned)((nnz + 255) / 256), 256>>>(di.p, dj.p, dv.p, nnz, L, base, dout.p, dbad.p);
        CHR3D_CUDA(cudaGetLastError());
    }
    int bad = 0;
    CHR3D_CUDA(cudaMemcpy(&bad, dbad.p, sizeof(int), cudaMemcpyDeviceToHost));
    if (bad) { set_error("chr3d_if_from_coo: bin index outside [%d, %d]", base, L - 1 + base); return CHR3D_ERR_ARG; }
    CHR3D_CUDA(cudaMemcpy(out, dout.p, sizeof(double) * (size_t)L * L, cudaMemcpyDeviceToHost));
    return 0;
}
