// Windowed, multi-sweep ("wavefront") versions of the two iterations that dominate an FCT step:
//   * the Jacobi sweeps of the low-order solve        (helpers.py:1782, spsolve replaced by Jacobi)
//   * the Chebyshev iterations of ChebSI              (helpers.py:143-185)
//
// Both are stencil-like passes x_{s+1} = f(A x_s, ...) over the SAME matrix, repeated 14-20 times.  Launched one sweep
// at a time every sweep streams all its operands from HBM and its neighbour gathers are latency-bound LSU loads.
// Here:
//
//  1. Every operand of a block of 256 rows is staged by the copy engine (1-D TMA bulk copies on mbarriers) by a
//     dedicated producer warp, up to NST-1 blocks ahead: the 16-bit template codes, the right-hand side, the matrix values
//     (Jacobi), and the gathered iterate itself as up to three contiguous *windows* (the columns of 256 consecutive rows
//     of a banded P1 ordering fall into "far below / around / far above the diagonal" intervals, found per block at
//     set-up from the row templates).  The eight consumer warps only touch shared memory and write the new iterate.
//     Columns outside the staged windows (unstructured orderings, truncated windows) fall back to an L2 load, so the
//     windows are an optimisation and never a correctness condition.
//  2. All sweeps of a solve run in ONE persistent launch as a wavefront over (sweep, block) work items:  ticket
//     w -> (t = w / S, s = w % S), block j = t - s*lag.  Item (s, j) needs sweep s-1 on the blocks its columns live in;
//     those have smaller tickets, are handed out cyclically to co-resident CTAs, and publish a per-item flag (a
//     signaller warp: consumers' mbarrier -> fence.acq_rel.gpu -> st.relaxed.gpu; loader warp: ld.relaxed.gpu, prefetched
//     one item ahead -> fence.acq_rel.gpu -> fence.proxy.async -> TMA reads).  Sweep s+1
//     therefore trails sweep s by `lag` blocks and finds the matrix block, the right-hand side and the iterate in the
//     126 MB L2 instead of HBM: per solve the static operands cross the HBM interface once instead of S times.
//
// The arithmetic per row is exactly that of k_jacobi_sweep_tpl<.,true> / k_cheb_iter_tpl (same operands, same order),
// so results are bit-identical to the one-sweep-per-launch kernels; tests/test_gpu_parity.py checks that.
#include "fct_common.cuh"
#include "fct_pipe.cuh"
#include "../../include/fctpdeco.h"

#include <limits.h>
#include <stdlib.h>

#define WIN_WCAP 272                 // doubles per staged window
#define WIN_SMAX 32                  // sweeps per launch
#define WIN_THREADS (FCT_RB + 64)    // 8 consumer warps + loader warp + signaller warp
#define WIN_NCONS (FCT_RB / 32)

#define WIN_LT 8                     // distinct row templates a block may have on the fast path

struct __align__(16) WinMeta {       // head of the per-block record (64 B)
    int lo[3];                       // first column of the below / around / above window (even)
    int len[3];                      // staged length (even, <= WIN_WCAP; 0 = not staged)
    int dep_lo, dep_hi;              // aligned blocks that hold this block's columns
    int k0, k1;                      // CSR range of the block
    int fast;                        // 1: <= WIN_LT templates and every column inside a staged window
    int own_ib;                      // fast path: window index of the row's own entry = tid + own_ib
    int pad[4];
};

// Per-block record: everything static a block of FCT_RB rows needs, contiguous, fetched by one or two bulk copies.
//   [0,64)      WinMeta
//   [64,320)    local template code of every row (u8)
//   [320,576)   per local template: window index base of its 8 slots (gather index = tid + ib)
//   [576,1152)  per local template: 8 mass-matrix values + the diagonal (ChebSI)
//   [1152,1680) start of every row inside the staged range of matrix values, u16 (Jacobi)
#define REC_LCODE 64
#define REC_IB 320
#define REC_CHEB 576
#define REC_RS 1152
#define REC_BYTES 1680
// ring stage
#define ST_CODE16 REC_BYTES                  // slow path: 16-bit global template codes
#define ST_RP (ST_CODE16 + 512)              // slow path (Jacobi): rowptr
#define ST_WIN (ST_RP + 1040)
#define ST_OWN (ST_WIN + 3 * WIN_WCAP * 8)
#define ST_X (ST_OWN + 2048)                 // ChebSI: y_{k-2} (own rows); Jacobi: matrix values

enum { WIN_CHEB = 0, WIN_JAC = 1 };

struct WinSweeps {
    const double* in[WIN_SMAX];      // iterate gathered by sweep s
    const double* old[WIN_SMAX];     // ChebSI: y_{k-2} (own rows) or nullptr
    double* out[WIN_SMAX];
    double omega[WIN_SMAX];          // ChebSI
    int rb[WIN_SMAX], re[WIN_SMAX];  // rows written by sweep s
    int first[WIN_SMAX];             // ChebSI: iteration k == 1 (no matrix application)
    int nsweeps;
    int lag;                         // blocks by which sweep s+1 trails sweep s
    int dbg;                         // timing experiments only (FCT_WIN_DBG): 1 no dependency wait, 2 no proxy fence,
                                     // 4 proxy fence .global, 8 no signaller fence, 16 consumers skip the arithmetic
};

struct WinArgs {
    WinSweeps P;
    const uint16_t* code;            // row templates (fct_templates.cu)
    const int32_t* toff;
    const double* tval;
    const double* tdiag;
    const unsigned char* rec;        // per-block records (REC_BYTES each)
    const double* own;               // ChebSI: right-hand side g; Jacobi: b' (own rows)
    const double* Lv;                // Jacobi: row-scaled off-diagonals of the low-order operator
    const int32_t* rowptr;
    unsigned long long* jstate;      // Jacobi: norms [0,1,16,17], sweeps [4,12,18]; both: error word [13]
    int* flags;                      // [sweep][chunk of 32 aligned blocks]: blocks of the chunk that finished the sweep
    int nchunks;                     // chunks spanned by [blk_lo, blk_hi)
    double dscale;                   // ChebSI
    int blk_lo, blk_hi, n, cap, nb_all, own_rb, own_re, fixed_sweeps;
};

extern __shared__ __align__(16) unsigned char win_smem[];

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ int ld_relaxed_gpu(const int* p) {
    int v;
    asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_relaxed_gpu_add(int* p, int v) {
    asm volatile("red.relaxed.gpu.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_acq_rel_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
// mbarrier wait that reports instead of hanging the GPU if a transaction count were ever wrong
__device__ __forceinline__ void mbar_wait_bounded(uint64_t* bar, uint32_t parity, unsigned long long* err) {
    uint32_t done = 0;
    const uint32_t addr = smem_u32(bar);
    long long t0 = 0;
    for (unsigned it = 0; !done; ++it) {
        asm volatile(
            "{\n"
            " .reg .pred p;\n"
            " mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            " selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
        if (!done && (it & 1023u) == 1023u) {
            if (t0 == 0) t0 = clock64();
            else if (clock64() - t0 > 6000000000ll) { *err = 2ull; return; }
        }
    }
}
// Non-blocking variant for the lane-parallel roles: try_wait may suspend the whole warp, which would stall the other
// lanes' (independent) waits; test_wait only polls.
__device__ __forceinline__ void mbar_poll_bounded(uint64_t* bar, uint32_t parity, unsigned long long* err) {
    uint32_t done = 0;
    const uint32_t addr = smem_u32(bar);
    long long t0 = 0;
    for (unsigned it = 0; !done; ++it) {
        asm volatile(
            "{\n"
            " .reg .pred p;\n"
            " mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            " selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
        if (!done && (it & 4095u) == 4095u) {
            if (t0 == 0) t0 = clock64();
            else if (clock64() - t0 > 6000000000ll) { *err = 2ull; return; }
        }
    }
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

// ---- set-up: the record of every aligned row block ---------------------------------------------------------------
__device__ __forceinline__ int win_class(int off) { return off < -1 ? 0 : (off > 1 ? 2 : 1); }

__global__ void __launch_bounds__(FCT_RB)
k_win_records(const int32_t* __restrict__ rowptr, const uint16_t* __restrict__ code, const int32_t* __restrict__ toff,
              const double* __restrict__ tval, const double* __restrict__ tdiag, int n, unsigned char* __restrict__ recs) {
    __shared__ int smn[3][FCT_RB / 32], smx[3][FCT_RB / 32];
    __shared__ int s_code[FCT_RB], s_list[WIN_LT], s_nlt, s_bad, s_lo[3], s_len[3];
    const int j = blockIdx.x, tid = threadIdx.x;
    const int r0 = j * FCT_RB;
    const int nr = min(FCT_RB, n - r0);
    const int r = r0 + tid;
    unsigned char* rec = recs + (size_t)j * REC_BYTES;
    int mn[3] = {INT_MAX, INT_MAX, INT_MAX}, mx[3] = {INT_MIN, INT_MIN, INT_MIN};
    int t = -1, len = 0;
    if (tid < nr) {
        t = code[r];
        len = rowptr[r + 1] - rowptr[r];
        for (int q = 0; q < FCT_TPL_W && q < len; ++q) {
            const int off = toff[FCT_TPL_W * t + q];
            const int c = r + off, w = win_class(off);
            for (int k = 0; k < 3; ++k)
                if (k == w) { mn[k] = min(mn[k], c); mx[k] = max(mx[k], c); }
        }
        // the row's own entry (padded template slots point at it) is always looked up in the middle window
        mn[1] = min(mn[1], r); mx[1] = max(mx[1], r);
    }
    s_code[tid] = t;
    for (int k = 0; k < 3; ++k) {
        const int a = __reduce_min_sync(0xffffffffu, mn[k]), b = __reduce_max_sync(0xffffffffu, mx[k]);
        if ((tid & 31) == 0) { smn[k][tid >> 5] = a; smx[k][tid >> 5] = b; }
    }
    __syncthreads();
    if (tid == 0) {
        WinMeta m;
        int allmn = INT_MAX, allmx = INT_MIN;
        for (int q = 0; q < 3; ++q) {
            int a = INT_MAX, b = INT_MIN;
            for (int i = 0; i < FCT_RB / 32; ++i) { a = min(a, smn[q][i]); b = max(b, smx[q][i]); }
            int lo = 0, ln = 0;
            if (a <= b) {
                allmn = min(allmn, a); allmx = max(allmx, b);
                lo = a & ~1;
                ln = ((b + 1 - lo) + 1) & ~1;
                if (lo + ln > n) ln = (n - lo) & ~1;            // never read past the vector; the odd tail is an L2 load
                if (ln > WIN_WCAP || ln < 0) ln = 0;            // does not fit: this window is not staged
            }
            m.lo[q] = lo; m.len[q] = ln;
            s_lo[q] = lo; s_len[q] = ln;
        }
        m.dep_lo = allmn <= allmx ? allmn / FCT_RB : j;
        m.dep_hi = allmn <= allmx ? allmx / FCT_RB : j;
        m.k0 = rowptr[r0];
        m.k1 = rowptr[r0 + nr];
        // distinct templates of the block (rows of one mesh line share one): at most WIN_LT on the fast path
        int nlt = 0, bad = 0, prev = -2;
        for (int i = 0; i < nr; ++i) {
            const int c = s_code[i];
            if (c == prev) continue;
            prev = c;
            int f = -1;
            for (int k = 0; k < nlt; ++k) if (s_list[k] == c) f = k;
            if (f < 0) { if (nlt < WIN_LT) s_list[nlt++] = c; else bad = 1; }
        }
        if (m.k1 - (m.k0 & ~1) > 65000) bad = 1;                // row starts are stored as u16
        s_nlt = nlt; s_bad = bad;
        m.fast = 0;
        m.own_ib = r0 - m.lo[1] + WIN_WCAP;
        for (int q = 0; q < 4; ++q) m.pad[q] = 0;
        *reinterpret_cast<WinMeta*>(rec) = m;
    }
    __syncthreads();
    // coverage: every column of every row inside its staged window?
    int inside = 1;
    if (tid < nr) {
        for (int q = 0; q < FCT_TPL_W && q < len; ++q) {
            const int off = toff[FCT_TPL_W * t + q];
            const int w = win_class(off);
            const int idx = r + off - s_lo[w];
            if (idx < 0 || idx >= s_len[w]) inside = 0;
        }
        const int io = r - s_lo[1];
        if (io < 0 || io >= s_len[1]) inside = 0;
    }
    const int all_inside = __syncthreads_and(inside);
    const int fast = all_inside && !s_bad;
    const int nlt = s_nlt;
    int lc = 0;
    if (fast && tid < nr)
        for (int k = 0; k < nlt; ++k) if (s_list[k] == t) lc = k;
    rec[REC_LCODE + tid] = (unsigned char)lc;
    if (tid < WIN_LT * 8) {
        const int lt = tid >> 3, q = tid & 7;
        int ib = 0;
        if (fast && lt < nlt) {
            const int off = toff[FCT_TPL_W * s_list[lt] + q];
            const int w = win_class(off);
            ib = r0 + off - s_lo[w] + w * WIN_WCAP;
        }
        reinterpret_cast<int*>(rec + REC_IB)[tid] = ib;
    }
    if (tid < WIN_LT * 9) {
        const int lt = tid / 9, q = tid - lt * 9;
        double v = 0.0;
        if (fast && lt < nlt) v = q < 8 ? tval[FCT_TPL_W * s_list[lt] + q] : tdiag[s_list[lt]];
        reinterpret_cast<double*>(rec + REC_CHEB)[tid] = v;
    }
    {
        const int ka = rowptr[r0] & ~1;
        uint16_t* rs = reinterpret_cast<uint16_t*>(rec + REC_RS);
        rs[tid] = (uint16_t)((tid <= nr && !s_bad) ? rowptr[r0 + tid] - ka : 0);
        if (tid < 8) rs[FCT_RB + tid] = (uint16_t)((tid == 0 && nr == FCT_RB && !s_bad) ? rowptr[r0 + FCT_RB] - ka : 0);
    }
    if (tid == 0) reinterpret_cast<WinMeta*>(rec)->fast = fast;
}

// ---- the wavefront kernel ------------------------------------------------------------------------------------------
#define CHEB_STAGE_BYTES (ST_X + 2048)
__host__ __device__ __forceinline__ size_t jac_stage_bytes(int cap) { return (size_t)ST_X + (size_t)cap * 8; }

// gathered iterate: window hit -> shared memory, else L2
__device__ __forceinline__ double win_get(const double* __restrict__ sw, const double* __restrict__ xin, int c, int off,
                                          int lo0, int lo1, int lo2, int n0, int n1, int n2) {
    const int lo = off < -1 ? lo0 : (off > 1 ? lo2 : lo1);
    const int nn = off < -1 ? n0 : (off > 1 ? n2 : n1);
    const int base = off < -1 ? 0 : (off > 1 ? 2 * WIN_WCAP : WIN_WCAP);
    const int idx = c - lo;
    if ((unsigned)idx < (unsigned)nn) return sw[base + idx];
    return __ldcg(xin + c);
}

// ticket -> work item (s, j): w = t * S + s, j = blk_lo + t - s * lag; tickets are dealt cyclically to the CTAs
struct WinIter {
    int w, t, s, j;
    int total, S, lag, blk_lo, blk_hi, gt, gs, G;
    __device__ __forceinline__ bool seek() {
        for (; w < total; ) {
            j = blk_lo + t - s * lag;
            if (j >= blk_lo && j < blk_hi) return true;
            step();
        }
        return false;
    }
    __device__ __forceinline__ void step() {
        w += G; t += gt; s += gs;
        if (s >= S) { s -= S; ++t; }
    }
    __device__ __forceinline__ bool first(int total_, int S_, int lag_, int lo_, int hi_) {
        total = total_; S = S_; lag = lag_; blk_lo = lo_; blk_hi = hi_;
        G = (int)gridDim.x; gt = G / S; gs = G - gt * S;
        w = (int)blockIdx.x; t = w / S; s = w - t * S; j = 0;
        return seek();
    }
    __device__ __forceinline__ bool next() { step(); return seek(); }
};

template <int KIND, int NST>
__global__ void __launch_bounds__(WIN_THREADS)
k_win(const __grid_constant__ WinArgs a) {
    __shared__ __align__(8) uint64_t full[NST], empty[NST], freeb[NST];
    const WinSweeps& P = a.P;
    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int i = 0; i < NST; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], WIN_NCONS); mbar_init(&freeb[i], 1); }
        mbar_fence_init();
    }
    __syncthreads();
    unsigned long long* const ERRP = a.jstate + 13;
    int S = P.nsweeps;
    if (KIND == WIN_JAC) {
        // the sweep count of this launch is decided on the device (k_jacobi_win_decide adapts it from solve to solve)
        if (a.fixed_sweeps > 0) {
            S = a.fixed_sweeps;
        } else {
            S = (int)a.jstate[12];
            if (S < 2) S = 14;
        }
        if (S > P.nsweeps) S = P.nsweeps;
    }
    const int blk_lo = a.blk_lo, blk_hi = a.blk_hi, n = a.n;
    const int total = (blk_hi - blk_lo + (S - 1) * P.lag) * S;
    const size_t sbytes = KIND == WIN_CHEB ? (size_t)CHEB_STAGE_BYTES : jac_stage_bytes(a.cap);
    const int warp = tid >> 5, lane = tid & 31;
    WinIter it;
    bool valid = it.first(total, S, P.lag, blk_lo, blk_hi);
    int cnt = 0;

    // Loader and signaller warps are lane-parallel: lane k (< NST) owns ring stage k and handles every NST-th item of
    // this CTA, so NST items are in flight through each role's latency chain (meta / counters / fences / copy issue).
    const int chunk_lo = blk_lo >> 5;
    if (warp == WIN_NCONS + 1) {
        // ---------------- signaller: publish finished items ----------------
        if (lane >= NST) return;
        for (int q = 0; q < lane && valid; ++q) valid = it.next();
        for (int k = 0; valid; ++k) {
            mbar_poll_bounded(&empty[lane], (uint32_t)(k & 1), ERRP);      // all consumer warps have stored their rows
            mbar_arrive(&freeb[lane]);                                      // the stage can be refilled at once ...
            if (S > 1) {                                                    // ... while the rows are published
                if (!(P.dbg & 8)) fence_acq_rel_gpu();
                red_relaxed_gpu_add(a.flags + (size_t)it.s * a.nchunks + ((it.j >> 5) - chunk_lo), 1);
            }
            for (int q = 0; q < NST && valid; ++q) valid = it.next();
        }
        return;
    }

    if (warp == WIN_NCONS) {
        // ---------------- loader ----------------
        if (lane >= NST) return;
        for (int q = 0; q < lane && valid; ++q) valid = it.next();
        // dependencies of (s, j): sweep s-1 complete on the 32-block chunks that hold blocks dep_lo..dep_hi (counters)
        struct Dep { const int* f; int c0, nc, v0, v1, v2; };
        auto chunk_need = [&](int c) { return min(blk_hi, (c + 1) << 5) - max(blk_lo, c << 5); };
        auto dep_issue = [&](int s, const WinMeta& mm) {
            Dep d{nullptr, 0, 0, 0, 0, 0};
            if (s > 0) {
                const int dlo = max(mm.dep_lo, blk_lo), dhi = min(mm.dep_hi, blk_hi - 1);
                d.f = a.flags + (size_t)(s - 1) * a.nchunks - chunk_lo;
                d.c0 = dlo >> 5; d.nc = (dhi >> 5) - d.c0 + 1;
                d.v0 = ld_relaxed_gpu(d.f + d.c0);
                if (d.nc > 1) d.v1 = ld_relaxed_gpu(d.f + d.c0 + 1);
                if (d.nc > 2) d.v2 = ld_relaxed_gpu(d.f + d.c0 + 2);
            }
            return d;
        };
        auto dep_ok = [&](const Dep& d) {
            if (d.nc > 3) return false;
            bool ok = d.v0 >= chunk_need(d.c0);
            if (d.nc > 1) ok = ok && d.v1 >= chunk_need(d.c0 + 1);
            if (d.nc > 2) ok = ok && d.v2 >= chunk_need(d.c0 + 2);
            return ok;
        };
        auto meta_of = [&](int j) { return *reinterpret_cast<const WinMeta*>(a.rec + (size_t)j * REC_BYTES); };
        WinMeta m;
        Dep dep{nullptr, 0, 0, 0, 0, 0};
        if (valid) { m = meta_of(it.j); dep = dep_issue(it.s, m); }
        for (int k = 0; valid; ++k) {
            WinIter nx = it;
            bool valid2 = true;
            for (int q = 0; q < NST && valid2; ++q) valid2 = nx.next();
            WinMeta m2;
            if (valid2) m2 = meta_of(nx.j);             // in flight while this item is issued
            const int stage = lane;
            const int s = it.s, j = it.j;
            unsigned char* sb = win_smem + (size_t)stage * sbytes;
            const unsigned char* grec = a.rec + (size_t)j * REC_BYTES;
            double* sw = reinterpret_cast<double*>(sb + ST_WIN);
            double* sown = reinterpret_cast<double*>(sb + ST_OWN);
            const int r0 = j * FCT_RB;
            const int nr = min(FCT_RB, n - r0);
            const int vcnt = nr & ~1;                   // own-row vector elements the copy engine brings
            const uint32_t cbytes = (uint32_t)(((nr + 7) & ~7) * 2);
            const bool first = KIND == WIN_CHEB && P.first[s];
            const bool has_old = KIND == WIN_CHEB && !first && P.old[s] != nullptr;
            const int rpcnt = min(260, (nr + 1 + 3) & ~3);     // rowptr entries staged (allocation has 8 entries of slack)
            const int ka = m.k0 & ~1;
            const int lcnt = ((m.k1 - ka) + 1) & ~1;          // Lvals carries 8 entries of slack
            const bool slow = !m.fast;
            // dependencies first (while the stage is still being consumed), copies after: a fence issued behind this lane's
            // own bulk copies would wait for them
            if (!first && s > 0 && !(P.dbg & 1)) {
                if (!dep_ok(dep)) {
                    // not there yet (or an unstructured dependency range): poll, bounded
                    const long long t0 = clock64();
                    for (int c = dep.c0; c < dep.c0 + dep.nc; ++c) {
                        while (ld_relaxed_gpu(dep.f + c) < chunk_need(c)) {
                            if (*reinterpret_cast<volatile unsigned long long*>(ERRP)) break;
                            if (clock64() - t0 > 4000000000ll) { *ERRP = 1ull; break; }   // ~2 s: report, do not hang
                            __nanosleep(100);
                        }
                    }
                }
                fence_acq_rel_gpu();
                if (P.dbg & 4) asm volatile("fence.proxy.async.global;" ::: "memory");
                else if (!(P.dbg & 2)) fence_proxy_async_all();
            }
            mbar_poll_bounded(&freeb[stage], (uint32_t)((k & 1) ^ 1), ERRP);
            uint32_t bytes = (KIND == WIN_CHEB ? (uint32_t)REC_RS : (uint32_t)REC_CHEB + (uint32_t)(REC_BYTES - REC_RS)) +
                             (uint32_t)vcnt * 8u;
            if (slow) bytes += cbytes + (KIND == WIN_JAC ? (uint32_t)rpcnt * 4u : 0u);
            if (!first) bytes += (uint32_t)(m.len[0] + m.len[1] + m.len[2]) * 8u;
            if (has_old) bytes += (uint32_t)vcnt * 8u;
            if (KIND == WIN_JAC) bytes += (uint32_t)lcnt * 8u;
            if (P.dbg & 32) bytes += 48u;
            mbar_expect_tx(&full[stage], bytes);
            if (KIND == WIN_CHEB) {
                tma_load_1d(sb, grec, (uint32_t)REC_RS, &full[stage]);                     // meta, codes, index bases, values
            } else {
                tma_load_1d(sb, grec, (uint32_t)REC_CHEB, &full[stage]);                   // meta, codes, index bases
                tma_load_1d(sb + REC_RS, grec + REC_RS, (uint32_t)(REC_BYTES - REC_RS), &full[stage]);   // row starts
                if (lcnt) tma_load_1d(sb + ST_X, a.Lv + ka, (uint32_t)lcnt * 8u, &full[stage]);
            }
            if (slow) {
                tma_load_1d(sb + ST_CODE16, a.code + r0, cbytes, &full[stage]);
                if (KIND == WIN_JAC) tma_load_1d(sb + ST_RP, a.rowptr + r0, (uint32_t)rpcnt * 4u, &full[stage]);
            }
            if (vcnt) tma_load_1d(sown, a.own + r0, (uint32_t)vcnt * 8u, &full[stage]);
            if (!first) {
                const double* xin = P.in[s];
#pragma unroll
                for (int q = 0; q < 3; ++q) {
                    if (!m.len[q]) continue;
                    if (P.dbg & 64) {        // experiment: same bytes in twice as many copies
                        const int h = (m.len[q] / 2) & ~1;
                        if (h) tma_load_1d(sw + q * WIN_WCAP, xin + m.lo[q], (uint32_t)h * 8u, &full[stage]);
                        tma_load_1d(sw + q * WIN_WCAP + h, xin + m.lo[q] + h, (uint32_t)(m.len[q] - h) * 8u, &full[stage]);
                    } else {
                        tma_load_1d(sw + q * WIN_WCAP, xin + m.lo[q], (uint32_t)m.len[q] * 8u, &full[stage]);
                    }
                }
                if (has_old && vcnt) tma_load_1d(sb + ST_X, P.old[s] + r0, (uint32_t)vcnt * 8u, &full[stage]);
            }
            if (P.dbg & 32) {        // experiment: three extra 16-byte copies (op count up, bytes unchanged)
                tma_load_1d(sb + ST_CODE16, a.code + r0, 16u, &full[stage]);
                tma_load_1d(sb + ST_CODE16 + 16, a.code + r0 + 64, 16u, &full[stage]);
                tma_load_1d(sb + ST_CODE16 + 32, a.code + r0 + 128, 16u, &full[stage]);
            }
            it = nx; valid = valid2;
            if (valid2) { m = m2; dep = dep_issue(it.s, m); }
        }
        return;
    }

    // ---------------- consumer warps ----------------
    const int s_last = S - 1, s_early = S - 3;
    double dl = 0.0, xl = 0.0, de = 0.0, xe = 0.0;      // Jacobi: ||dx||, ||x|| of the last and of the early-check sweep
    while (valid) {
        const int stage = cnt % NST;
        const uint32_t par = (uint32_t)((cnt / NST) & 1);
        ++cnt;
        const int s = it.s, j = it.j;
        const unsigned char* sb = win_smem + (size_t)stage * sbytes;
        const WinMeta* sm = reinterpret_cast<const WinMeta*>(sb);
        const double* sw = reinterpret_cast<const double*>(sb + ST_WIN);
        const double* sown = reinterpret_cast<const double*>(sb + ST_OWN);
        const int r0 = j * FCT_RB;
        const int nr = min(FCT_RB, n - r0);
        const int vcnt = nr & ~1;
        const int r = r0 + tid;
        mbar_wait_bounded(&full[stage], par, ERRP);
        if (tid < nr && !(P.dbg & 16)) {
            const double own = tid < vcnt ? sown[tid] : __ldcg(a.own + r);
            double xnew;
            if (KIND == WIN_CHEB && P.first[s]) {
                // iteration k == 1: y1 = omega1 * g / (dscale * Md)
                const double md = sm->fast ? reinterpret_cast<const double*>(sb + REC_CHEB)[9 * sb[REC_LCODE + tid] + 8]
                                           : __ldg(a.tdiag + reinterpret_cast<const uint16_t*>(sb + ST_CODE16)[tid]);
                const double z = own / (a.dscale * md);
                xnew = P.omega[s] * z;
            } else if (sm->fast) {
                // every operand in shared memory: gather index = tid + base of (local template, slot)
                const int lc = sb[REC_LCODE + tid];
                const int4 i0 = *reinterpret_cast<const int4*>(sb + REC_IB + 32 * lc);
                const int4 i1 = *reinterpret_cast<const int4*>(sb + REC_IB + 32 * lc + 16);
                const double* swt = sw + tid;
                const double x0 = swt[i0.x], x1 = swt[i0.y], x2 = swt[i0.z], x3 = swt[i0.w];
                const double x4 = swt[i1.x], x5 = swt[i1.y], x6 = swt[i1.z], x7 = swt[i1.w];
                if (KIND == WIN_CHEB) {
                    const double* vv = reinterpret_cast<const double*>(sb + REC_CHEB) + 9 * lc;
                    const double ym = swt[sm->own_ib];
                    const double yo = P.old[s] ? (tid < vcnt ? reinterpret_cast<const double*>(sb + ST_X)[tid] : __ldcg(P.old[s] + r))
                                               : 0.0;
                    double acc = 0.0;
                    acc += vv[0] * x0; acc += vv[1] * x1; acc += vv[2] * x2; acc += vv[3] * x3;
                    acc += vv[4] * x4; acc += vv[5] * x5; acc += vv[6] * x6; acc += vv[7] * x7;
                    const double z = (own - acc) / (a.dscale * vv[8]);
                    xnew = P.omega[s] * (z + ym - yo) + yo;
                } else {
                    const uint16_t* rs = reinterpret_cast<const uint16_t*>(sb + REC_RS);
                    const int k0r = rs[tid];
                    const int len = (int)rs[tid + 1] - k0r;
                    const double* sLr = reinterpret_cast<const double*>(sb + ST_X) + k0r;
                    const double xv[8] = {x0, x1, x2, x3, x4, x5, x6, x7};
                    double acc = 0.0;
#pragma unroll
                    for (int q = 0; q < 8; ++q) acc += ((q < len) ? sLr[q] : 0.0) * xv[q];
                    xnew = own - acc;
                    if ((s == s_last || s == s_early) && r >= a.own_rb && r < a.own_re) {
                        const double d = fabs(xnew - swt[sm->own_ib]), ax = fabs(xnew);
                        if (s == s_last) { dl = fmax(dl, d); xl = fmax(xl, ax); }
                        else { de = fmax(de, d); xe = fmax(xe, ax); }
                    }
                }
            } else {
                // generic path: global template tables, columns outside the staged windows come from L2
                const int tc = reinterpret_cast<const uint16_t*>(sb + ST_CODE16)[tid];
                const double* xin = P.in[s];
                const int lo0 = sm->lo[0], lo1 = sm->lo[1], lo2 = sm->lo[2];
                const int n0 = sm->len[0], n1 = sm->len[1], n2 = sm->len[2];
                const int4 o0 = __ldg(reinterpret_cast<const int4*>(a.toff + FCT_TPL_W * tc));
                const int4 o1 = __ldg(reinterpret_cast<const int4*>(a.toff + FCT_TPL_W * tc) + 1);
                const int off[8] = {o0.x, o0.y, o0.z, o0.w, o1.x, o1.y, o1.z, o1.w};
                double xv[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) xv[q] = win_get(sw, xin, r + off[q], off[q], lo0, lo1, lo2, n0, n1, n2);
                const double xo = win_get(sw, xin, r, 0, lo0, lo1, lo2, n0, n1, n2);
                if (KIND == WIN_CHEB) {
                    const double md = __ldg(a.tdiag + tc);
                    const double2 v0 = __ldg(reinterpret_cast<const double2*>(a.tval + FCT_TPL_W * tc));
                    const double2 v1 = __ldg(reinterpret_cast<const double2*>(a.tval + FCT_TPL_W * tc) + 1);
                    const double2 v2 = __ldg(reinterpret_cast<const double2*>(a.tval + FCT_TPL_W * tc) + 2);
                    const double2 v3 = __ldg(reinterpret_cast<const double2*>(a.tval + FCT_TPL_W * tc) + 3);
                    double yo = 0.0;
                    if (P.old[s]) yo = tid < vcnt ? reinterpret_cast<const double*>(sb + ST_X)[tid] : __ldcg(P.old[s] + r);
                    double acc = 0.0;
                    acc += v0.x * xv[0]; acc += v0.y * xv[1]; acc += v1.x * xv[2]; acc += v1.y * xv[3];
                    acc += v2.x * xv[4]; acc += v2.y * xv[5]; acc += v3.x * xv[6]; acc += v3.y * xv[7];
                    const double z = (own - acc) / (a.dscale * md);
                    xnew = P.omega[s] * (z + xo - yo) + yo;
                } else {
                    const int32_t* srp = reinterpret_cast<const int32_t*>(sb + ST_RP);
                    const int rpcnt = min(260, (nr + 1 + 3) & ~3);
                    const int ka = sm->k0 & ~1;
                    const int k0r = srp[tid];
                    const int len = ((tid + 1 < rpcnt) ? srp[tid + 1] : a.rowptr[r + 1]) - k0r;
                    const double* sLr = reinterpret_cast<const double*>(sb + ST_X) + (k0r - ka);
                    double acc = 0.0;
#pragma unroll
                    for (int q = 0; q < 8; ++q) acc += ((q < len) ? sLr[q] : 0.0) * xv[q];
                    xnew = own - acc;
                    if ((s == s_last || s == s_early) && r >= a.own_rb && r < a.own_re) {
                        const double d = fabs(xnew - xo), ax = fabs(xnew);
                        if (s == s_last) { dl = fmax(dl, d); xl = fmax(xl, ax); }
                        else { de = fmax(de, d); xe = fmax(xe, ax); }
                    }
                }
            }
            if (r >= P.rb[s] && r < P.re[s]) P.out[s][r] = xnew;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[stage]);
        valid = it.next();
    }
    if (KIND == WIN_JAC) {
        dl = warp_max(dl); xl = warp_max(xl); de = warp_max(de); xe = warp_max(xe);
        if (lane == 0) {
            atomicMax(a.jstate + 0, (unsigned long long)__double_as_longlong(dl));
            atomicMax(a.jstate + 1, (unsigned long long)__double_as_longlong(xl));
            atomicMax(a.jstate + 16, (unsigned long long)__double_as_longlong(de));
            atomicMax(a.jstate + 17, (unsigned long long)__double_as_longlong(xe));
        }
        if (blockIdx.x == 0 && tid == 0) { atomicAdd(a.jstate + 4, (unsigned long long)S); a.jstate[18] = (unsigned long long)S; }
    }
}

// After the fused sweeps: stopping test of the last sweep (as k_jacobi_decide) and adaptation of the sweep count of the
// next solve: two sweeps fewer if the test had already passed two sweeps earlier, two more if it failed (the remaining
// sweeps of THIS solve are done by the one-sweep-per-launch loop that follows).
__global__ void k_jacobi_win_decide(unsigned long long* __restrict__ jstate, double rtol, int smax) {
    const double delta = __longlong_as_double((long long)jstate[0]);
    const double xm = __longlong_as_double((long long)jstate[1]);
    const double de = __longlong_as_double((long long)jstate[16]);
    const double xe = __longlong_as_double((long long)jstate[17]);
    const int S = (int)jstate[18];
    jstate[5] = jstate[0];
    jstate[6] = jstate[1];
    int next = S;
    if (delta <= rtol * xm) {
        jstate[3] = 1ull;
        jstate[10] = jstate[4] > 2ull ? jstate[4] - 2ull : 0ull;
        if (S >= 4 && xe > 0.0 && de <= rtol * xe) next = S - 2;
    } else {
        next = S + 2;
        jstate[11] += 1ull;
    }
    if (next < 2) next = 2;
    if (next > smax) next = smax;
    jstate[12] = (unsigned long long)next;
    jstate[0] = 0ull; jstate[1] = 0ull; jstate[16] = 0ull; jstate[17] = 0ull;
}

// ================================================================================================================
// host side
// ================================================================================================================
struct fct_win {
    unsigned char* rec = nullptr;
    int* flags = nullptr;
    int nb_all = 0;
    int grid_cheb = 0, grid_jac = 0;
    int nst_cheb = 3, nst_jac = 2;
    int lag_cheb = 0, lag_jac = 0;
    int max_dep = 0;
    int nfast = 0;                   // blocks on the all-shared-memory path
    bool cheb_on = true, jac_on = true;
    bool single = false;             // FCT_WIN_SINGLE=1: one launch per sweep through the windowed kernel
};

void fct_win_free(fct_ctx* ctx) {
    if (!ctx->win) return;
    cudaFree(ctx->win->rec);
    cudaFree(ctx->win->flags);
    delete ctx->win;
    ctx->win = nullptr;
}

static int env_int(const char* name, int dflt) {
    const char* e = getenv(name);
    return e ? atoi(e) : dflt;
}

typedef void (*win_kernel_t)(const WinArgs);
static win_kernel_t win_kernel(int kind, int nst) {
    if (kind == WIN_CHEB) return nst == 2 ? k_win<WIN_CHEB, 2> : nst == 4 ? k_win<WIN_CHEB, 4> : nst == 5 ? k_win<WIN_CHEB, 5> : k_win<WIN_CHEB, 3>;
    return nst == 3 ? k_win<WIN_JAC, 3> : nst == 4 ? k_win<WIN_JAC, 4> : k_win<WIN_JAC, 2>;
}

// Called after the row templates exist (fct_templates_build).  Never fails the caller: without windows the
// one-sweep-per-launch kernels are used.
int fct_win_build(fct_ctx* ctx) {
    fct_win_free(ctx);
    if (!ctx->tpl_count) return 0;
    if (env_int("FCT_WIN", 0) == 0) return 0;      // opt-in (FCT_WIN=1) until it beats the one-sweep-per-launch kernels
    fct_win* W = new fct_win();
    const int n = ctx->n;
    W->nb_all = (n + FCT_RB - 1) / FCT_RB;
    bool ok = false;
    do {
        if ((long long)W->nb_all * WIN_SMAX * 4 > 2000000000ll) break;       // tickets are 32-bit
        if (cudaMalloc((void**)&W->rec, (size_t)REC_BYTES * (size_t)W->nb_all + 64) != cudaSuccess) break;
        if (cudaMalloc((void**)&W->flags, sizeof(int) * ((size_t)W->nb_all / 32 + 4) * WIN_SMAX) != cudaSuccess) break;
        k_win_records<<<W->nb_all, FCT_RB, 0, ctx->stream>>>(ctx->rowptr, ctx->tpl_code, ctx->tpl_off, ctx->tpl_val, ctx->tpl_diag, n,
                                                              W->rec);
        ctx->launches++;
        // largest dependency reach (in blocks) decides the minimum lag
        WinMeta* h = (WinMeta*)malloc(sizeof(WinMeta) * (size_t)W->nb_all);
        if (!h) break;
        if (cudaMemcpy2DAsync(h, sizeof(WinMeta), W->rec, REC_BYTES, sizeof(WinMeta), (size_t)W->nb_all, cudaMemcpyDeviceToHost,
                              ctx->stream) != cudaSuccess ||
            cudaStreamSynchronize(ctx->stream) != cudaSuccess) { free(h); break; }
        int reach = 0;
        for (int j = 0; j < W->nb_all; ++j) {
            reach = reach > h[j].dep_hi - j ? reach : h[j].dep_hi - j;
            reach = reach > j - h[j].dep_lo ? reach : j - h[j].dep_lo;
            W->nfast += h[j].fast ? 1 : 0;
        }
        free(h);
        W->max_dep = reach;
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, ctx->device) != cudaSuccess) break;
        W->nst_cheb = env_int("FCT_WIN_NST_CHEB", 3);
        if (W->nst_cheb < 2 || W->nst_cheb > 5) W->nst_cheb = 3;
        W->nst_jac = env_int("FCT_WIN_NST_JAC", 2);
        if (W->nst_jac < 2 || W->nst_jac > 4) W->nst_jac = 2;
        const size_t smc = (size_t)W->nst_cheb * CHEB_STAGE_BYTES;
        const size_t smj = (size_t)W->nst_jac * jac_stage_bytes(ctx->cap);
        if (smj > (size_t)FCT_SMEM_OPTIN) W->jac_on = false;
        if (smc > (size_t)FCT_SMEM_OPTIN) W->cheb_on = false;
        int occ_c = 0, occ_j = 0;
        if (W->cheb_on) {
            cudaFuncSetAttribute(win_kernel(WIN_CHEB, W->nst_cheb), cudaFuncAttributeMaxDynamicSharedMemorySize, FCT_SMEM_OPTIN);
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_c, win_kernel(WIN_CHEB, W->nst_cheb), WIN_THREADS, smc);
        }
        if (W->jac_on) {
            cudaFuncSetAttribute(win_kernel(WIN_JAC, W->nst_jac), cudaFuncAttributeMaxDynamicSharedMemorySize, FCT_SMEM_OPTIN);
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_j, win_kernel(WIN_JAC, W->nst_jac), WIN_THREADS, smj);
        }
        if (cudaGetLastError() != cudaSuccess) break;
        const int oc = env_int("FCT_WIN_OCC_CHEB", 0), oj = env_int("FCT_WIN_OCC_JAC", 0);
        if (oc >= 1 && oc < occ_c) occ_c = oc;
        if (oj >= 1 && oj < occ_j) occ_j = oj;
        if (occ_c < 1) W->cheb_on = false;
        if (occ_j < 1) W->jac_on = false;
        W->grid_cheb = prop.multiProcessorCount * (occ_c > 0 ? occ_c : 1);
        W->grid_jac = prop.multiProcessorCount * (occ_j > 0 ? occ_j : 1);
        // lag: dependency reach + the tickets that are in flight at any time (grid x stages), in time steps of S tickets
        // each; FCT_WIN_LAG_* override (blocks)
        W->lag_cheb = env_int("FCT_WIN_LAG_CHEB", 0);
        W->lag_jac = env_int("FCT_WIN_LAG_JAC", 0);
        W->single = env_int("FCT_WIN_SINGLE", 0) != 0;
        if (env_int("FCT_WIN_CHEB", 1) == 0) W->cheb_on = false;
        if (env_int("FCT_WIN_JAC", 1) == 0) W->jac_on = false;
        ok = true;
    } while (0);
    cudaGetLastError();
    if (!ok) { cudaFree(W->rec); cudaFree(W->flags); delete W; return 0; }
    ctx->win = W;
    return 0;
}

static inline bool aligned16(const void* p) { return ((uintptr_t)p & 15u) == 0; }

static int win_lag(const fct_win* W, int forced, int grid, int nst, int S) {
    // dependencies are tracked per chunk of 32 blocks: a whole chunk beyond the reach must be done
    if (forced > 0) return forced > W->max_dep + 32 ? forced : W->max_dep + 33;
    const int inflight = (grid * (nst + 1) + S - 1) / S;        // time steps covered by the tickets in flight
    return W->max_dep + 33 + inflight;
}

static void win_common_args(fct_ctx* ctx, WinArgs& a) {
    fct_win* W = ctx->win;
    a.code = ctx->tpl_code; a.toff = ctx->tpl_off; a.tval = ctx->tpl_val; a.tdiag = ctx->tpl_diag; a.rec = W->rec;
    a.rowptr = ctx->rowptr; a.jstate = ctx->jstate; a.flags = W->flags;
    a.blk_lo = ctx->cur_rb / FCT_RB; a.blk_hi = (ctx->cur_re + FCT_RB - 1) / FCT_RB;
    a.n = ctx->n; a.cap = ctx->cap; a.nb_all = W->nb_all; a.own_rb = ctx->row_begin; a.own_re = ctx->row_end;
    a.nchunks = ((a.blk_hi + 31) >> 5) - (a.blk_lo >> 5) + 1;
    a.P.dbg = env_int("FCT_WIN_DBG", 0);
}

// ChebSI through the wavefront kernel; returns 1 if it ran, 0 if the caller must use the per-iteration kernels, < 0 on error
int fct_win_chebsi(fct_ctx* ctx, const double* b, double* y, int iters, double lmin, double lmax) {
    fct_win* W = ctx->win;
    if (!W || !W->cheb_on || ctx->comm || iters < 1 || iters > WIN_SMAX) return 0;
    if (!aligned16(b)) return 0;
    const double rho = (lmax - lmin) / (lmax + lmin);
    WinArgs a;
    memset(&a, 0, sizeof(a));
    win_common_args(ctx, a);
    a.dscale = (lmin + lmax) / 2;
    a.own = b;
    WinSweeps& P = a.P;
    double* buf[3] = {ctx->w[0], ctx->w[1], ctx->w[2]};
    double omega = 0.0;
    const double* ymid = nullptr;
    const double* yold = nullptr;
    for (int k = 1; k <= iters; ++k) {
        if (k == 2) omega = 1 / (1 - rho * rho / 2);
        else omega = 1 / (1 - (omega * rho * rho) / 4);
        double* ynew = (k == iters) ? y : buf[k % 3];
        const int s = k - 1;
        P.in[s] = ymid; P.old[s] = yold; P.out[s] = ynew; P.omega[s] = omega;
        P.rb[s] = ctx->cur_rb; P.re[s] = ctx->cur_re; P.first[s] = (k == 1);
        yold = ymid; ymid = ynew;
    }
    P.nsweeps = iters;
    P.lag = win_lag(W, W->lag_cheb, W->grid_cheb, W->nst_cheb, iters);
    if (a.blk_hi <= a.blk_lo) return 1;
    if (iters > 1) {
        if (cudaMemsetAsync(W->flags, 0, sizeof(int) * (size_t)a.nchunks * iters, ctx->stream) != cudaSuccess) return -1;
    }
    const long long total = (long long)(a.blk_hi - a.blk_lo + (iters - 1) * P.lag) * iters;
    if (total > 2000000000ll) return 0;
    const int grid = (long long)W->grid_cheb < total ? W->grid_cheb : (int)total;
    if (W->single) {
        // one launch per iteration through the same kernel (all operands staged by the copy engine, no dependencies)
        WinArgs b1 = a;
        b1.P.nsweeps = 1;
        for (int s = 0; s < iters; ++s) {
            b1.P.in[0] = a.P.in[s]; b1.P.old[0] = a.P.old[s]; b1.P.out[0] = a.P.out[s]; b1.P.omega[0] = a.P.omega[s];
            b1.P.first[0] = a.P.first[s];
            const int nblk = a.blk_hi - a.blk_lo;
            win_kernel(WIN_CHEB, W->nst_cheb)<<<W->grid_cheb < nblk ? W->grid_cheb : nblk, WIN_THREADS,
                                               (size_t)W->nst_cheb * CHEB_STAGE_BYTES, ctx->stream>>>(b1);
            ctx->launches++;
        }
        return 1;
    }
    win_kernel(WIN_CHEB, W->nst_cheb)<<<grid, WIN_THREADS, (size_t)W->nst_cheb * CHEB_STAGE_BYTES, ctx->stream>>>(a);
    ctx->launches++;
    return 1;
}

static bool win_jacobi_args(fct_ctx* ctx, WinArgs& a, const double* Lv, const double* b, double* x, double* tmp, int smax,
                            int fixed) {
    fct_win* W = ctx->win;
    memset(&a, 0, sizeof(a));
    win_common_args(ctx, a);
    a.own = b; a.Lv = Lv; a.fixed_sweeps = fixed;
    WinSweeps& P = a.P;
    for (int s = 0; s < smax; ++s) {
        P.in[s] = (s & 1) ? tmp : x;
        P.out[s] = (s & 1) ? x : tmp;
        P.rb[s] = ctx->cur_rb; P.re[s] = ctx->cur_re;
    }
    P.nsweeps = smax;
    P.lag = win_lag(W, W->lag_jac, W->grid_jac, W->nst_jac, fixed > 0 ? fixed : 14);
    const long long total = (long long)(a.blk_hi - a.blk_lo + (smax - 1) * P.lag) * smax;
    return total <= 2000000000ll;
}

// Fused Jacobi sweeps of the row-scaled low-order system (jac_mode 2): up to `smax` sweeps (even), the count of this
// launch is jstate[12] (adapted on the device).  Returns 1 if enqueued, 0 if unavailable.
int fct_win_jacobi(fct_ctx* ctx, const double* Lv, const double* b, double* x, double* tmp, double rtol, int smax) {
    fct_win* W = ctx->win;
    if (!W || !W->jac_on || ctx->comm || ctx->jac_mode != 2) return 0;
    if (!aligned16(Lv) || !aligned16(b) || !aligned16(x) || !aligned16(tmp)) return 0;
    if (smax > WIN_SMAX) smax = WIN_SMAX;
    smax &= ~1;
    if (smax < 2) return 0;
    WinArgs a;
    if (!win_jacobi_args(ctx, a, Lv, b, x, tmp, smax, 0)) return 0;
    if (a.blk_hi <= a.blk_lo) return 1;
    if (cudaMemsetAsync(W->flags, 0, sizeof(int) * (size_t)a.nchunks * smax, ctx->stream) != cudaSuccess) return -1;
    win_kernel(WIN_JAC, W->nst_jac)<<<W->grid_jac, WIN_THREADS, (size_t)W->nst_jac * jac_stage_bytes(ctx->cap), ctx->stream>>>(a);
    k_jacobi_win_decide<<<1, 1, 0, ctx->stream>>>(ctx->jstate, rtol, smax);
    ctx->launches += 2;
    return 1;
}

// Instrumentation for bench.py: `reps` fused launches of exactly `sweeps` Jacobi sweeps on the low-order system that
// fct_bench_jacobi_sweeps has just built (Lvals / w[3] / w[4] / w[5]); returns the CUDA-event time per sweep.
int fct_win_bench_jacobi(fct_ctx* ctx, int sweeps, int reps, int warm, float* ms_per_sweep) {
    fct_win* W = ctx->win;
    if (!W || !W->jac_on || ctx->comm || ctx->jac_mode != 2) return 0;
    sweeps &= ~1;
    if (sweeps < 2 || sweeps > WIN_SMAX) return 0;
    WinArgs a;
    if (!win_jacobi_args(ctx, a, ctx->Lvals, ctx->w[3], ctx->w[4], ctx->w[5], sweeps, sweeps)) return 0;
    const size_t sm = (size_t)W->nst_jac * jac_stage_bytes(ctx->cap);
    cudaEvent_t e0, e1;
    if (cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess) return -1;
    for (int pass = warm ? 0 : 1; pass < 2; ++pass) {
        if (pass == 1) cudaEventRecord(e0, ctx->stream);
        for (int i = 0; i < reps; ++i) {
            if (W->single) {
                // experiment: the windowed kernel as a one-sweep-per-launch kernel (no dependencies, no flags)
                WinArgs b = a;
                b.fixed_sweeps = 1; b.P.nsweeps = 1;
                for (int q = 0; q < sweeps; ++q) {
                    b.P.in[0] = a.P.in[q]; b.P.out[0] = a.P.out[q];
                    win_kernel(WIN_JAC, W->nst_jac)<<<W->grid_jac, WIN_THREADS, sm, ctx->stream>>>(b);
                }
                continue;
            }
            cudaMemsetAsync(W->flags, 0, sizeof(int) * (size_t)a.nchunks * sweeps, ctx->stream);
            win_kernel(WIN_JAC, W->nst_jac)<<<W->grid_jac, WIN_THREADS, sm, ctx->stream>>>(a);
        }
    }
    cudaEventRecord(e1, ctx->stream);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *ms_per_sweep = ms / (reps * sweeps);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

// Test hook: exactly `sweeps` (even) fused sweeps on the system fct_bench_jacobi_sweeps built, result in w[4].
int fct_win_jacobi_fixed(fct_ctx* ctx, int sweeps) {
    float ms = 0.f;
    return fct_win_bench_jacobi(ctx, sweeps, 1, 0, &ms);
}
