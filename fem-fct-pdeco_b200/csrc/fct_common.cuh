// Shared device/host plumbing for libfctpdeco (sm_100a).
//
// Kernel skeleton used by every CSR pass ("row-block staging"):
//   * one CTA owns FCT_RB consecutive rows and therefore one contiguous range [k0,k1) of the CSR
//     value / index arrays;
//   * phase 1 streams that range into shared memory with 16-byte, L1-bypassing loads, perfectly
//     coalesced whatever the row lengths are;
//   * phase 2 is thread-per-row on shared memory (P1 rows have <= 7..9 entries; the 8-byte stride
//     between consecutive rows is odd on P1 meshes, so the row walk is bank-conflict free);
//     neighbour values x[col] are gathered through L1/L2 -- consecutive rows gather consecutive
//     addresses, so these gathers coalesce too;
//   * results that are value arrays go back through shared memory and a coalesced phase 3.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#define FCT_RB 256            // rows (= threads) per CTA
#define FCT_ALIGN 4           // staging alignment in elements (16 B for int32, 32 B for fp64)
#define FCT_SMEM_OPTIN (200 * 1024)   // dynamic shared memory opt-in ceiling for every row-block kernel
#define FCT_TPL_W 8            // row-template width (entries per row), fct_templates.cu
#define FCT_TPL_MAX 65535

void fct_set_error(const char* fmt, ...);

// Checked allocator (fct_guard.cu).  With FCT_GUARD=1 in the environment every device buffer of the library gets a 4 KB
// canary band on both sides and a 0xFF fill (NaN doubles / -1 indices); fct_guard_check counts the buffers whose canaries were
// overwritten.  compute-sanitizer is not available on the target pool, so this plus the oracle comparisons is how
// out-of-bounds writes and reads of uninitialised memory are caught.  Off (default): plain cudaMalloc / cudaFree.
cudaError_t fct_guard_malloc(void** p, size_t bytes);
cudaError_t fct_guard_free(void* p);
#ifndef FCT_GUARD_IMPL
#define cudaMalloc(p, b) fct_guard_malloc((void**)(p), (size_t)(b))
#define cudaFree(p) fct_guard_free((void*)(p))
#endif

#define FCT_CUDA(call)                                                                          \
    do {                                                                                        \
        cudaError_t e__ = (call);                                                               \
        if (e__ != cudaSuccess) {                                                               \
            fct_set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
            return 1;                                                                           \
        }                                                                                       \
    } while (0)

#define FCT_CHECK(cond, ...)                                                                    \
    do {                                                                                        \
        if (!(cond)) {                                                                          \
            fct_set_error(__VA_ARGS__);                                                         \
            return 2;                                                                           \
        }                                                                                       \
    } while (0)

// staging state of the host-trajectory entry point (fct_advdrift_state_host), allocated on first use
struct fct_hoststage {
    bool ready = false;
    cudaStream_t d2h_stream = 0;
    double* cbuf[2] = {nullptr, nullptr};
    double* ubuf[3] = {nullptr, nullptr, nullptr};
    cudaEvent_t c_ready[2] = {}, c_free[2] = {}, u_ready[3] = {}, u_free[3] = {};
};

struct fct_comm;   // NCCL state (fct_comm.cu)
struct fct_p2p;    // NVLink peer-memory mailboxes (fct_p2p.cu)
struct fct_tiles;  // structured-numbering data of the overlapped-tile kernels (fct_tile.cu)
struct fct_cheb_tiles;

// CUDA-graph WHILE loop of the low-order Jacobi solve (fct_kernels.cu), cached per operand set
struct fct_jgraph {
    void* graph = nullptr;
    void* exec = nullptr;
    const double *Lv = nullptr, *b = nullptr, *dinv = nullptr;
    double *x = nullptr, *tmp = nullptr;
    double rtol = 0.0;
    int max_sweeps = 0;
    int depth = 0;
    int mode = -1;
    const void* tpl = nullptr;   // template table the captured sweeps point into (rebuilt by fct_ctx_set_mass)
};

struct fct_ctx {
    int device = 0;
    cudaStream_t stream = 0;
    cudaStream_t copy_stream = 0;
    int32_t n = 0;              // local rows (= local vector length)
    int64_t nnz = 0;
    int32_t row_begin = 0, row_end = 0;
    // Deep halos (multi-GPU): ring j = local rows within j mesh rings of the owned rows, ring 0 = owned, ring `depth`
    // = all local rows.  A pass computed on ring j reads its inputs on ring j+1, so after one exchange (valid on ring
    // `depth`) up to `depth` dependent passes run without communication, each on a ring one smaller.
    int32_t depth = 1;
    int32_t ring_lo[9] = {0}, ring_hi[9] = {0};
    int32_t cur_rb = 0, cur_re = 0;          // row range of the next row-block launches (default: the owned rows)
    int32_t cap = 0;            // max staged entries of any row block (incl. alignment slack)
    int32_t max_row = 0;
    int32_t grid_cap = 148 * 8; // persistent grid: SMs x resident 256-thread CTAs
    int32_t nst1 = 2, grid_nst1 = 148 * 4;       // ring stages / persistent grid of k_cheb_iter, k_jacobi_sweep
    int32_t nst_jtpl = 3;                        // ring stages of k_jacobi_sweep_tpl (FCT_NST_JTPL)
    int32_t grid_jtpl = 148 * 4;                 // persistent grid of k_jacobi_sweep_tpl
    int32_t grid_flux_tpl = 148;                 // persistent grid of the template variants of the flux kernels
    int32_t grid_low[2] = {148, 148};            // persistent grids of k_low_build<0/1>
    int32_t grid_pipe1 = 148, grid_pipe2 = 148;   // persistent grids of the TMA-ring kernels (1 / 2 fp64 arrays)
    int32_t* rowptr = nullptr;  // device
    int32_t* colidx = nullptr;
    int32_t* tpos = nullptr;
    // mesh
    int64_t ncells = 0;
    int32_t* cells = nullptr;       // [ncells*3]
    double* xy = nullptr;           // [n*2]
    int32_t* v2c_ptr = nullptr;     // vertex -> incident cells (CSR), built on set_mesh
    int32_t* v2c_idx = nullptr;     // [2 * incidences]: the two other vertices of each incident cell, cyclic order
    // geometry templates of the mesh (fct_assembly.cu): 16-bit code per row + table; gt_count == 0: generic kernels
    uint16_t* gt_code = nullptr;
    void* gt_tab = nullptr;
    int32_t gt_count = 0;
    // static matrices
    double* M = nullptr;
    double* ML = nullptr;
    double* Mdiag = nullptr;
    double* K = nullptr;
    bool mass_set = false;
    // row templates of M (fct_templates.cu): 16-bit code per row, T x 8 column offsets / values
    uint16_t* tpl_code = nullptr;
    int32_t* tpl_off = nullptr;
    double* tpl_val = nullptr;
    double* tpl_diag = nullptr;      // [T] diagonal value of each template (= diag(M) of its rows)
    int32_t tpl_count = 0;
    // Low-order Jacobi on the row templates: 0 = CSR kernels, 1 = column offsets from the templates (bit-identical to 0),
    // 2 = additionally rows pre-scaled by 1/l_ii in k_low_build (no dinv read in the sweep).  FCT_JAC_TPL selects.
    int32_t jac_mode = 0;
    fct_tiles* tiles = nullptr;      // set by fct_ctx_set_rect: K Jacobi sweeps / Chebyshev iterations per launch (fct_tile.cu)
    bool tiles_ok = false;           // tile kernels usable (structured numbering verified, row templates present)
    int32_t tile_grid_cap = 0;       // FCT_TILE_GRID: cap on the CTAs of a tile launch (tests: several tiles per CTA on small meshes)
    fct_cheb_tiles* cheb_tiles = nullptr;   // tile lists + exception descriptors of the ChebSI tile kernel
    bool cheb_tiles_ok = false;
    int32_t tile_kj = 4;             // sweeps per fused Jacobi launch (FCT_TILE_KJ, 2..4)
    int32_t tile_kc = 5;             // iterations per fused ChebSI launch (FCT_TILE_KC = 2..5; 0 = per-iteration kernels)
    int32_t cheb_mdtab = 1;          // ChebSI takes diag(M) from the template table when the caller passes ctx->Mdiag
    // workspace
    double* Lvals = nullptr;    // low-order operator
    double* Dvals = nullptr;    // artificial diffusion (off-diagonals)
    double* Avals = nullptr;    // assembled operator (time loops) / host-call staging
    double* Svals = nullptr;    // host-call staging
    double* w[12] = {nullptr};  // vector workspace [n] each
    double* sys_m[6] = {nullptr};   // nnz-sized operators of the PDE-system time loops (fct_drivers.cu, allocated on first use)
    double* sys_v[4] = {nullptr};   // n-sized right-hand sides of those loops
    double* sys_wind = nullptr;     // 20 doubles: polynomial wind coefficients
    double* fb_w[12] = {nullptr};   // private workspace of the BiCGStab fallback of the low-order solve (allocated on first use)
    bool checked_steps = false;     // fct_step reads the Jacobi outcome back after every low-order solve and falls back to BiCGStab
    double* red = nullptr;      // device scalars for reductions (64 doubles)
    unsigned long long* jstate = nullptr;  // Jacobi state words
    double* pinned = nullptr;   // small pinned host buffer (64 doubles)
    double rtol = 1e-14;
    int32_t max_sweeps = 100;
    int64_t launches = 0;
    int64_t exchanges = 0;      // halo exchanges enqueued by the host (NCCL path; the peer-memory path counts on the device)
    fct_jgraph jgraph;
    fct_hoststage hs;
    bool tile_adapt = true;     // FCT_TILE_ADAPT=0: every fused Jacobi launch runs tile_kj sweeps (no device-side sweep schedule)
    bool tile_sched = false;    // the solve being enqueued uses the sweep schedule (set by fct_step)
    bool in_time_loop = false;  // a device-resident time loop is being enqueued (fct_drivers.cu): the only place the schedule is used
    bool use_pdl = false;       // FCT_PDL=1 enables programmatic dependent launch (measured: no gain, persistent grids)
    bool capturing = false;
    bool use_graph = true;      // FCT_NO_GRAPH=1 falls back to the static launch sequence with device-side early exit
    int32_t last_pairs = 0;     // Jacobi sweep pairs the previous multi-GPU solve needed
    fct_comm* comm = nullptr;
    fct_p2p* p2p = nullptr;
    // halo description (multi-GPU)
    int32_t send_lo[2] = {0, 0}, send_hi[2] = {0, 0};
};

static inline int fct_nblocks(const fct_ctx* c) { return (c->cur_re - c->cur_rb + FCT_RB - 1) / FCT_RB; }
static inline void fct_set_ring(fct_ctx* c, int j) {
    if (j < 0) j = 0;
    if (j > c->depth) j = c->depth;
    c->cur_rb = c->ring_lo[j];
    c->cur_re = c->ring_hi[j];
}
// launch grid of a persistent row-block kernel over `nblocks` row blocks
static inline int fct_grid(const fct_ctx* c, int nblocks) { return nblocks < c->grid_cap ? nblocks : c->grid_cap; }

#ifdef __CUDACC__
// ---- sweep schedule of the fused Jacobi tile launches (fct_tile.cu; used by the stopping-test kernels of fct_kernels.cu and
// fct_p2p.cu, which must agree to the bit) ---------------------------------------------------------------------------------
// A launch of K sweeps is tested once, at its end, so a solve that needs 14 sweeps costs 16 with launches of 4.  The sweep count
// of the low-order solve is nearly constant from one time step to the next: jstate[14] keeps the count the previous solve
// needed, the launches of the next solve are sized to end exactly there (4+4+4+2 for 14; the kernel reads its depth from
// jstate[13]).  The count is learnt from the last two tests of a solve (geometric decay of ||x_k - x_{k-1}||), and every
// `period` solves one sweep fewer is tried (a failed probe costs one 2-sweep launch and doubles the period).  Everything is
// derived from all-reduced test results, so N ranks and one GPU follow the same schedule; a time loop starts without history.
//   jstate[13] depth of the next launch   [14] learnt count (0: none)   [16] target of this solve   [17] solves since the last
//   probe   [18] probe period   [19] this solve is a probe   [20] delta bits / [21] sweep count of the last failed test
__device__ __forceinline__ unsigned long long tile_next_k(unsigned long long done, unsigned long long target,
                                                          unsigned long long kmax) {
    if (target == 0ull) return kmax;                 // nothing learnt yet
    if (done >= target) return 2ull;                 // past the expected count: short launches until the test passes
    const unsigned long long r = target - done;
    if (r > kmax + 1ull) return kmax;
    if (r == kmax + 1ull) return kmax >= 3ull ? kmax - 1ull : 2ull;      // never leave a single sweep for the last launch
    return r < 2ull ? 2ull : r;
}
__device__ __forceinline__ void tile_schedule_begin(unsigned long long* jstate, unsigned long long kmax) {
    const unsigned long long learnt = jstate[14];
    unsigned long long target = learnt, probing = 0ull;
    if (learnt > 6ull) {
        const unsigned long long period = jstate[18] ? jstate[18] : 8ull;
        if (++jstate[17] >= period) { target = learnt - 1ull; probing = 1ull; jstate[17] = 0ull; }
    }
    jstate[16] = target;
    jstate[19] = probing;
    jstate[20] = 0ull;
    jstate[21] = 0ull;
    jstate[13] = tile_next_k(0ull, target, kmax);
}
__device__ __forceinline__ void tile_schedule_failed(unsigned long long* jstate, double delta, unsigned long long kmax) {
    jstate[20] = (unsigned long long)__double_as_longlong(delta);
    jstate[21] = jstate[4];
    jstate[13] = tile_next_k(jstate[4], jstate[16], kmax);
}
// the test passed after jstate[4] sweeps with ||x_k - x_{k-1}|| = delta <= tol
__device__ __forceinline__ void tile_schedule_converged(unsigned long long* jstate, double tol, double delta) {
    const unsigned long long done = jstate[4];
    unsigned long long need = done;
    const unsigned long long sprev = jstate[21];
    const double dprev = __longlong_as_double((long long)jstate[20]);
    if (sprev > 0ull && sprev < done && dprev > tol && delta > 0.0 && delta < dprev) {
        // decay per sweep between the last failed test and this one; the first sweep count at which the test would pass
        const double lr = log(delta / dprev) / (double)(done - sprev);          // < 0
        const double extra = ceil(log(tol / dprev) / lr);
        if (extra >= 1.0 && extra < (double)(done - sprev)) need = sprev + (unsigned long long)extra;
    }
    if (jstate[19]) {
        const unsigned long long period = jstate[18] ? jstate[18] : 8ull;
        if (done <= jstate[16]) jstate[14] = done;                                   // the probe passed: one sweep fewer from now on
        else { jstate[18] = period < 64ull ? 2ull * period : 64ull; if (done > jstate[14] + 1ull) jstate[14] = need; }
    } else {
        jstate[14] = need;
    }
}

// ---- streaming loads (read-once data: bypass L1 so it stays free for the x gathers) -------------
__device__ __forceinline__ double2 ld_stream_f64x2(const double* p) {
    double2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ int4 ld_stream_s32x4(const int32_t* p) {
    int4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ double ld_stream_f64(const double* p) {
    double r;
    asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(r) : "l"(p));
    return r;
}

struct RowBlock {
    int r0;      // first row of this CTA
    int nr;      // rows in this CTA
    int k0, k1;  // CSR range
    int ka;      // k0 rounded down to FCT_ALIGN: shared index = k - ka
};

// Row-block kernels are persistent: a fixed grid (SM count x resident CTAs) strides over the row blocks, so a
// launch that has nothing to do (converged Jacobi sweep) costs microseconds instead of a 65k-CTA empty grid.
#define FCT_FOR_ROW_BLOCKS(blk, row_begin, row_end)                                                    \
    for (int blk = blockIdx.x, nblk__ = ((row_end) - (row_begin) + FCT_RB - 1) / FCT_RB; blk < nblk__; \
         blk += gridDim.x)

__device__ __forceinline__ RowBlock row_block(const int32_t* __restrict__ rowptr, int row_begin, int row_end, int blk) {
    RowBlock b;
    b.r0 = row_begin + blk * FCT_RB;
    b.nr = min(FCT_RB, row_end - b.r0);
    b.k0 = rowptr[b.r0];
    b.k1 = rowptr[b.r0 + b.nr];
    b.ka = b.k0 & ~(FCT_ALIGN - 1);
    return b;
}

// stage g[ka .. k1) into s[0 .. k1-ka); vector loads where the whole vector is inside [0,nnz)
__device__ __forceinline__ void stage_f64(double* __restrict__ s, const double* __restrict__ g, const RowBlock& b,
                                          int64_t nnz) {
    const int cnt = b.k1 - b.ka;
    for (int i = 2 * threadIdx.x; i < cnt; i += 2 * FCT_RB) {
        const int64_t k = (int64_t)b.ka + i;
        if (k + 1 < nnz) {
            double2 v = ld_stream_f64x2(g + k);
            s[i] = v.x;
            s[i + 1] = v.y;
        } else if (k < nnz) {
            s[i] = ld_stream_f64(g + k);
        }
    }
}
__device__ __forceinline__ void stage_s32(int32_t* __restrict__ s, const int32_t* __restrict__ g, const RowBlock& b,
                                          int64_t nnz) {
    const int cnt = b.k1 - b.ka;
    for (int i = 4 * threadIdx.x; i < cnt; i += 4 * FCT_RB) {
        const int64_t k = (int64_t)b.ka + i;
        if (k + 3 < nnz) {
            int4 v = ld_stream_s32x4(g + k);
            s[i] = v.x; s[i + 1] = v.y; s[i + 2] = v.z; s[i + 3] = v.w;
        } else {
            for (int q = 0; q < 4; ++q)
                if (k + q < nnz) s[i + q] = g[k + q];
        }
    }
}
// write s[k0-ka .. k1-ka) back to g[k0 .. k1): scalar 8-byte stores, fully coalesced
__device__ __forceinline__ void unstage_f64(double* __restrict__ g, const double* __restrict__ s, const RowBlock& b) {
    const int off = b.k0 - b.ka;
    const int cnt = b.k1 - b.k0;
    for (int i = threadIdx.x; i < cnt; i += FCT_RB) g[(int64_t)b.k0 + i] = s[off + i];
}

// ---- block reductions (FCT_RB threads) ----------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
// deterministic block sum: warp shuffles, then warp 0 adds the FCT_RB/32 partials in a fixed order
__device__ __forceinline__ double block_sum(double v, double* sred /* >= 8 doubles */) {
    v = warp_sum(v);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) sred[w] = v;
    __syncthreads();
    double t = 0.0;
    if (w == 0) {
        t = (l < FCT_RB / 32) ? sred[l] : 0.0;
        t = warp_sum(t);
    }
    return t;   // valid in warp 0
}
__device__ __forceinline__ double block_max(double v, double* sred) {
    v = warp_max(v);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) sred[w] = v;
    __syncthreads();
    double t = 0.0;
    if (w == 0) {
        t = (l < FCT_RB / 32) ? sred[l] : sred[0];
        t = warp_max(t);
    }
    return t;
}
__device__ __forceinline__ double block_min(double v, double* sred) {
    v = warp_min(v);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) sred[w] = v;
    __syncthreads();
    double t = 0.0;
    if (w == 0) {
        t = (l < FCT_RB / 32) ? sred[l] : sred[0];
        t = warp_min(t);
    }
    return t;
}
#endif  // __CUDACC__

// internal cross-file entry points
int fct_launch_error(fct_ctx* ctx, const char* what);
