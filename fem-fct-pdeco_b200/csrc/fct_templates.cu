// Row templates for the static mass matrix.
//
// ChebSI applies the SAME matrix M 19 times per FCT step (and 20 times per gradient slice): 12 B/nnz of CSR traffic
// per application is by far the largest item of the step.  On meshes with repeated element shapes most rows of M are
// copies of a few "templates": the same column offsets relative to the row and the same values.  At set-up the rows
// are hashed on the device, the distinct (offsets, values) tuples are collected (<= 65535, rows of <= 8 entries) and
// every row gets a 16-bit template code.  A matrix application then reads 2 B per ROW instead of ~84 B; the template
// table (a few MB at most) is served by L1/L2, and the neighbour gathers of consecutive rows are consecutive
// addresses.  The arithmetic is unchanged: same values, same column order, same summation order => bit-identical
// results to the CSR kernels.  If the matrix does not compress (too many templates, long rows, or a hash collision
// detected by the exact verification pass) the CSR/TMA kernels are used -- there is nothing mesh-specific here.
#include "fct_common.cuh"
#include "../../include/fctpdeco.h"

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_select.cuh>

__device__ __forceinline__ unsigned long long mix64(unsigned long long h, unsigned long long v) {
    h ^= v + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
    h *= 0xFF51AFD7ED558CCDull;
    h ^= h >> 33;
    return h;
}

__global__ void k_row_hash(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                           const double* __restrict__ vals, int n, unsigned long long* __restrict__ hash) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const int k0 = rowptr[r], k1 = rowptr[r + 1];
    unsigned long long h = mix64(0x1234567ull, (unsigned long long)(k1 - k0));
    for (int k = k0; k < k1; ++k) {
        h = mix64(h, (unsigned long long)(unsigned int)(colidx[k] - r));
        h = mix64(h, (unsigned long long)__double_as_longlong(vals[k]));
    }
    hash[r] = h;
}

__global__ void k_assign_codes(const unsigned long long* __restrict__ hash, const unsigned long long* __restrict__ uniq,
                               int T, int n, uint16_t* __restrict__ code, int* __restrict__ rep) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const unsigned long long h = hash[r];
    int lo = 0, hi = T - 1;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (uniq[mid] < h) lo = mid + 1; else hi = mid;
    }
    code[r] = (uint16_t)lo;
    atomicMin(rep + lo, r);
}

__global__ void k_fill_templates(const int* __restrict__ rep, const int32_t* __restrict__ rowptr,
                                 const int32_t* __restrict__ colidx, const double* __restrict__ vals, int T,
                                 int32_t* __restrict__ toff, double* __restrict__ tval, double* __restrict__ tdiag) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    const int r = rep[t];
    const int k0 = rowptr[r], len = rowptr[r + 1] - k0;
    double dg = 0.0;
    for (int j = 0; j < FCT_TPL_W; ++j) {
        toff[FCT_TPL_W * t + j] = (j < len) ? colidx[k0 + j] - r : 0;     // padding: offset 0, value 0
        tval[FCT_TPL_W * t + j] = (j < len) ? vals[k0 + j] : 0.0;
        if (j < len && colidx[k0 + j] == r) dg = vals[k0 + j];
    }
    tdiag[t] = dg;
}

// exact check: every row must reproduce its template bit for bit
__global__ void k_verify_templates(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                                   const double* __restrict__ vals, int n, const uint16_t* __restrict__ code,
                                   const int32_t* __restrict__ toff, const double* __restrict__ tval, int* __restrict__ bad) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const int k0 = rowptr[r], len = rowptr[r + 1] - k0;
    const int t = code[r];
    bool ok = len <= FCT_TPL_W;
    for (int j = 0; j < FCT_TPL_W && ok; ++j) {
        const int off = (j < len) ? colidx[k0 + j] - r : 0;
        const long long vb = (j < len) ? __double_as_longlong(vals[k0 + j]) : 0ll;
        ok = (toff[FCT_TPL_W * t + j] == off) && (__double_as_longlong(tval[FCT_TPL_W * t + j]) == vb);
    }
    if (!ok) atomicAdd(bad, 1);
}

int fct_tiles_prepare(fct_ctx* ctx);      // fct_tile.cu

void fct_templates_free(fct_ctx* ctx) {
    ctx->tiles_ok = false;           // the tile kernels' neighbour-delta table is indexed by template code
    cudaFree(ctx->tpl_code); cudaFree(ctx->tpl_off); cudaFree(ctx->tpl_val); cudaFree(ctx->tpl_diag);
    ctx->tpl_code = nullptr; ctx->tpl_off = nullptr; ctx->tpl_val = nullptr; ctx->tpl_diag = nullptr;
    ctx->tpl_count = 0;
    ctx->jac_mode = 0;
    ctx->jgraph.mode = -1;           // a captured Jacobi loop points into the freed tables: force a rebuild
}

// (Re)build the templates of ctx->M.  Never fails the caller: on any problem the context simply has no templates.
int fct_templates_build(fct_ctx* ctx) {
    fct_templates_free(ctx);
    const char* e = getenv("FCT_NO_TEMPLATES");
    if (e && atoi(e) == 1) return 0;
    if (ctx->max_row > FCT_TPL_W) return 0;
    const int n = ctx->n;
    unsigned long long *hash = nullptr, *sorted = nullptr, *uniq = nullptr;
    int *dT = nullptr, *rep = nullptr, *bad = nullptr;
    void* tmp = nullptr;
    size_t tmp_bytes = 0, tb2 = 0;
    int T = 0, hbad = 0;
    bool ok = false;
    cudaStream_t st = ctx->stream;
    do {
        if (cudaMalloc((void**)&hash, 8 * (size_t)n) != cudaSuccess) break;
        if (cudaMalloc((void**)&sorted, 8 * (size_t)n) != cudaSuccess) break;
        if (cudaMalloc((void**)&uniq, 8 * (size_t)n) != cudaSuccess) break;
        if (cudaMalloc((void**)&dT, sizeof(int) * 2) != cudaSuccess) break;
        k_row_hash<<<(n + 255) / 256, 256, 0, st>>>(ctx->rowptr, ctx->colidx, ctx->M, n, hash);
        cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, hash, sorted, n, 0, 64, st);
        cub::DeviceSelect::Unique(nullptr, tb2, sorted, uniq, dT, n, st);
        if (tb2 > tmp_bytes) tmp_bytes = tb2;
        if (cudaMalloc(&tmp, tmp_bytes) != cudaSuccess) break;
        cub::DeviceRadixSort::SortKeys(tmp, tmp_bytes, hash, sorted, n, 0, 64, st);
        cub::DeviceSelect::Unique(tmp, tmp_bytes, sorted, uniq, dT, n, st);
        if (cudaMemcpyAsync(&T, dT, sizeof(int), cudaMemcpyDeviceToHost, st) != cudaSuccess) break;
        if (cudaStreamSynchronize(st) != cudaSuccess) break;
        if (T < 1 || T > FCT_TPL_MAX) break;
        if (cudaMalloc((void**)&rep, sizeof(int) * (size_t)T) != cudaSuccess) break;
        if (cudaMalloc((void**)&bad, sizeof(int)) != cudaSuccess) break;
        if (cudaMalloc((void**)&ctx->tpl_code, sizeof(uint16_t) * ((size_t)n + 8)) != cudaSuccess) break;
        if (cudaMalloc((void**)&ctx->tpl_off, sizeof(int32_t) * FCT_TPL_W * (size_t)T) != cudaSuccess) break;
        if (cudaMalloc((void**)&ctx->tpl_val, sizeof(double) * FCT_TPL_W * (size_t)T) != cudaSuccess) break;
        if (cudaMalloc((void**)&ctx->tpl_diag, sizeof(double) * (size_t)T) != cudaSuccess) break;
        cudaMemsetAsync(rep, 0x7f, sizeof(int) * (size_t)T, st);
        cudaMemsetAsync(bad, 0, sizeof(int), st);
        k_assign_codes<<<(n + 255) / 256, 256, 0, st>>>(hash, uniq, T, n, ctx->tpl_code, rep);
        k_fill_templates<<<(T + 255) / 256, 256, 0, st>>>(rep, ctx->rowptr, ctx->colidx, ctx->M, T, ctx->tpl_off, ctx->tpl_val,
                                                        ctx->tpl_diag);
        k_verify_templates<<<(n + 255) / 256, 256, 0, st>>>(ctx->rowptr, ctx->colidx, ctx->M, n, ctx->tpl_code, ctx->tpl_off,
                                                            ctx->tpl_val, bad);
        ctx->launches += 4;
        if (cudaMemcpyAsync(&hbad, bad, sizeof(int), cudaMemcpyDeviceToHost, st) != cudaSuccess) break;
        if (cudaStreamSynchronize(st) != cudaSuccess) break;
        ok = (hbad == 0);
    } while (0);
    cudaFree(hash); cudaFree(sorted); cudaFree(uniq); cudaFree(dT); cudaFree(rep); cudaFree(bad); cudaFree(tmp);
    cudaGetLastError();
    if (ok) {
        ctx->tpl_count = T;
        const char* jm = getenv("FCT_JAC_TPL");      // 0: CSR sweeps, 1: template columns, 2 (default): + pre-scaled rows
        ctx->jac_mode = (jm && atoi(jm) >= 0 && atoi(jm) <= 2) ? atoi(jm) : 2;
        const char* mt = getenv("FCT_CHEB_MDTAB");   // 0: ChebSI always reads Md from memory
        ctx->cheb_mdtab = (mt && atoi(mt) == 0) ? 0 : 1;
        fct_tiles_prepare(ctx);
    } else {
        fct_templates_free(ctx);
    }
    return 0;
}

// ---- matrix applications through the templates -------------------------------------------------------------
__device__ __forceinline__ double tpl_row_dot_t(int t, const int32_t* __restrict__ toff, const double* __restrict__ tval,
                                                const double* __restrict__ x, int r);
__device__ __forceinline__ double tpl_row_dot(const uint16_t* __restrict__ code, const int32_t* __restrict__ toff,
                                              const double* __restrict__ tval, const double* __restrict__ x, int r) {
    return tpl_row_dot_t((int)code[r], toff, tval, x, r);
}
__device__ __forceinline__ double tpl_row_dot_t(int t, const int32_t* __restrict__ toff, const double* __restrict__ tval,
                                                const double* __restrict__ x, int r) {
    const int4 o0 = __ldg(reinterpret_cast<const int4*>(toff + FCT_TPL_W * t));
    const int4 o1 = __ldg(reinterpret_cast<const int4*>(toff + FCT_TPL_W * t) + 1);
    const double2 v0 = __ldg(reinterpret_cast<const double2*>(tval + FCT_TPL_W * t));
    const double2 v1 = __ldg(reinterpret_cast<const double2*>(tval + FCT_TPL_W * t) + 1);
    const double2 v2 = __ldg(reinterpret_cast<const double2*>(tval + FCT_TPL_W * t) + 2);
    const double2 v3 = __ldg(reinterpret_cast<const double2*>(tval + FCT_TPL_W * t) + 3);
    const double x0 = x[r + o0.x], x1 = x[r + o0.y], x2 = x[r + o0.z], x3 = x[r + o0.w];
    const double x4 = x[r + o1.x], x5 = x[r + o1.y], x6 = x[r + o1.z], x7 = x[r + o1.w];
    double acc = 0.0;          // same order as the CSR row loop; padded slots add +0.0 * x[r]
    acc += v0.x * x0; acc += v0.y * x1; acc += v1.x * x2; acc += v1.y * x3;
    acc += v2.x * x4; acc += v2.y * x5; acc += v3.x * x6; acc += v3.y * x7;
    return acc;
}

// Measured on B200 (4097^2): this plain grid-stride form runs at 0.148 ms (4.7 TB/s of actual traffic); fetching the
// next block's inputs ahead (0.175 ms) or several rows per thread (0.195 ms) is slower -- both widen the window of
// rows in flight, and the neighbour gathers then miss in L2 more often.
template <bool PF, bool PF2>
__global__ void __launch_bounds__(FCT_RB, 8)
k_cheb_iter_tpl(const uint16_t* __restrict__ code, const int32_t* __restrict__ toff, const double* __restrict__ tval,
                const double* __restrict__ tdiag, const double* __restrict__ Md, const double* __restrict__ g,
                const double* __restrict__ ymid, const double* __restrict__ yold, double* __restrict__ ynew, double omega,
                double dscale, int has_old, int row_begin, int row_end) {
    const int stride = (int)gridDim.x * FCT_RB;
    int r = row_begin + (int)blockIdx.x * FCT_RB + (int)threadIdx.x;
    // PF: the template code of the thread's next row is fetched one iteration ahead (one register), so the
    // code -> table -> gather chain of an iteration starts from the (L1-resident) table instead of DRAM
    int t = (PF && r < row_end) ? (int)code[r] : 0;
    for (; r < row_end; r += stride) {
        if (!PF) t = code[r];
        const int tn = (PF && r + stride < row_end) ? (int)code[r + stride] : 0;
        if (PF2 && (threadIdx.x & 15) == 0 && r + stride < row_end) {
            // pull the next iteration's own-row operands into L2 (one request per 128-byte line, no registers); the
            // neighbour rows it gathers are own rows of other CTAs, prefetched by them
            asm volatile("prefetch.global.L2 [%0];" ::"l"(g + r + stride));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(ymid + r + stride));
            if (has_old) asm volatile("prefetch.global.L2 [%0];" ::"l"(yold + r + stride));
        }
        // Md == nullptr: the caller's Md is diag(M) of the templated matrix itself, so it comes out of the table
        // (same bits) and the 8 B/row read is saved
        const double gr = g[r], mdr = Md ? Md[r] : __ldg(tdiag + t), ym = ymid[r];
        const double yo = has_old ? yold[r] : 0.0;
        const double acc = tpl_row_dot_t(t, toff, tval, ymid, r);
        const double z = (gr - acc) / (dscale * mdr);
        ynew[r] = omega * (z + ym - yo) + yo;
        if (PF) t = tn;
    }
}

__global__ void __launch_bounds__(FCT_RB)
k_spmv_tpl(const uint16_t* __restrict__ code, const int32_t* __restrict__ toff, const double* __restrict__ tval,
           const double* __restrict__ x, double alpha, double beta, const double* __restrict__ z, double* __restrict__ y,
           int row_begin, int row_end) {
    const int stride = (int)gridDim.x * FCT_RB;
    for (int r = row_begin + (int)blockIdx.x * FCT_RB + (int)threadIdx.x; r < row_end; r += stride) {
        const double acc = tpl_row_dot(code, toff, tval, x, r);
        double out = alpha * acc;
        if (beta != 0.0) out += beta * z[r];
        y[r] = out;
    }
}

static inline int tpl_grid(const fct_ctx* ctx) {
    const int nb = fct_nblocks(ctx);
    int cap = ctx->grid_cap;             // SMs x 8 resident 256-thread CTAs
    static int per_sm = -1;              // tuning knob: FCT_TPL_CTAS = CTAs per SM (1..8)
    if (per_sm < 0) { const char* e = getenv("FCT_TPL_CTAS"); per_sm = (e && atoi(e) >= 1 && atoi(e) <= 8) ? atoi(e) : 0; }
    if (per_sm > 0) cap = ctx->grid_cap / 8 * per_sm;
    return nb < cap ? nb : cap;
}

int fct_cheb_iter_tpl(fct_ctx* ctx, const double* Md, const double* g, const double* ymid, const double* yold,
                      double* ynew, double omega, double dscale) {
    const int grid = tpl_grid(ctx);
    if (grid > 0) {
        const double* Mdk = (ctx->cheb_mdtab && Md == ctx->Mdiag) ? nullptr : Md;
        static int pf = -1;                  // FCT_CHEB_PF: 0 plain, 1 (default) next row's template code one iteration ahead,
        if (pf < 0) {                        // 2: + L2 prefetch of the next iteration's own-row operands
            const char* e = getenv("FCT_CHEB_PF");
            pf = (e && atoi(e) >= 0 && atoi(e) <= 2) ? atoi(e) : 1;
        }
#define CHEB_TPL_ARGS ctx->tpl_code, ctx->tpl_off, ctx->tpl_val, ctx->tpl_diag, Mdk, g, ymid, yold, ynew, omega, dscale, \
                      yold != nullptr, ctx->cur_rb, ctx->cur_re
        if (pf == 2) k_cheb_iter_tpl<true, true><<<grid, FCT_RB, 0, ctx->stream>>>(CHEB_TPL_ARGS);
        else if (pf == 1) k_cheb_iter_tpl<true, false><<<grid, FCT_RB, 0, ctx->stream>>>(CHEB_TPL_ARGS);
        else k_cheb_iter_tpl<false, false><<<grid, FCT_RB, 0, ctx->stream>>>(CHEB_TPL_ARGS);
#undef CHEB_TPL_ARGS
        ctx->launches++;
    }
    return 0;
}

int fct_spmv_tpl(fct_ctx* ctx, const double* x, double alpha, double beta, const double* z, double* y) {
    const int grid = tpl_grid(ctx);
    if (grid > 0) {
        k_spmv_tpl<<<grid, FCT_RB, 0, ctx->stream>>>(ctx->tpl_code, ctx->tpl_off, ctx->tpl_val, x, alpha, beta, z, y,
                                                    ctx->cur_rb, ctx->cur_re);
        ctx->launches++;
    }
    return 0;
}
