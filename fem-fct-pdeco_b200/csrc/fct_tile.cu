// Overlapped 2-D tiles for the two iterative kernels of the FCT step on structured (RectangleMesh) numberings:
// K Jacobi sweeps of the low-order solve (helpers.py:1782; k_tile) or K Chebyshev iterations of ChebSI (helpers.py:143-185; k_cheb_tile) per
// launch, out of shared memory and registers.
//
// dolfin's CG1 numbering on RectangleMesh is an anti-diagonal numbering: row = start(d) + pos with
// d = ix - iy + n.  Every neighbour of (d, pos) lies in (d +- 1, pos +- 1), so a rectangle in (d, pos) space plus a K-wide
// frame holds everything K dependent passes over its interior need, and each of its diagonals is one contiguous piece of
// every vector and of the CSR value array.  A CTA (one per SM, persistent over its tiles) owns a 40 x 40 region:
//   * the NEXT tile streams in while the current one is computed: the matrix rows of each region diagonal with 1-D TMA
//     bulk copies, the iterate with 8-byte cp.async, the right-hand side into registers, all issued at the start of the
//     current tile's passes and (the shared-memory part) completing on one mbarrier;
//   * each compute thread keeps the matrix rows (<= 8 values each) and the right-hand side of its (at most 3) rows in
//     REGISTERS for all K passes, so a pass costs 7 shared-memory loads + 7 DFMA per row; the iterate ping-pongs between
//     two shared-memory buffers; pass s updates the rows at least s away from the region's edge (redundant work in the
//     frame instead of inter-CTA flags); only the interior is written back;
//   * neighbour positions come from the row templates (fct_templates.cu): per template 8 signed bytes (delta of the
//     shared-memory index), so boundary rows, corners and the truncated rows of a multi-GPU block need no special case.
// HBM traffic per K passes: the matrix values, b and x once (+ the frame, mostly L2 hits on the neighbouring tile's
// interior) instead of K times.  The arithmetic of a row is the same sequence of operations as in the per-pass kernels
// (k_jacobi_sweep_tpl, k_cheb_iter_tpl): results are bit-identical for the same number of passes
// (tests/test_gpu_parity.py::test_tile_kernels_bit_identical).
//
// Multi-GPU: a rank's rows are the global rows [g0, g0 + n); tiles cover the owned rows and read K rings of halo, so
// with halo depth >= K one launch replaces the K per-ring launches between two exchanges.
// General meshes (fct_ctx_set_rect not called, or its verification failed) keep the per-pass kernels.
#include "fct_common.cuh"
#include "fct_pipe.cuh"
#include "../../include/fctpdeco.h"

#include <stdlib.h>
#include <vector>

extern __shared__ __align__(16) unsigned char fct_smem[];

#define TL_ND 40                      // region: diagonals
#define TL_NP 40                      // region: positions per diagonal
#define TL_NC (TL_ND - 2)             // rows that are ever updated ("compute set"): 38 x 38
#define TL_NQ (TL_NP - 2)
#define TL_XS TL_NP                   // row stride of the iterate buffers
#define TL_R 3                        // compute-set rows per thread
#define TL_NT 512                     // threads per CTA (16 warps leave 128 registers per thread), one CTA per SM
#define TL_THREADS TL_NT
#define TL_KMAX 5
static_assert(TL_R * TL_NT >= TL_NC * TL_NQ, "every compute-set row needs a thread");
static_assert(TL_NQ > 32, "the row -> thread map deals 32 positions of a diagonal to a warp and the rest separately");
static_assert(TL_XS + 1 <= 127, "neighbour deltas must fit a signed byte");
// tile flags, set by k_tile_classify from the data (never assumed): every row of the compute set exists, has 7 entries and
// the interior neighbour layout above (UP) / below (LO) the main anti-diagonal; all its mass-matrix rows carry the same values
#define TL_F_UP 1
#define TL_F_LO 2
#define TL_F_MUNI 4

struct fct_tiles {
    int n_cells = 0;                  // cells per side of the structured mesh
    int g0 = 0;                       // global DoF index of local row 0
    unsigned long long* tdelta = nullptr;      // [templates] 8 signed bytes: shared-memory index delta of each row entry
    int4* list[TL_KMAX + 1] = {nullptr};       // per K: tile records {d0, p0, flags, value code} covering the owned rows
    int count[TL_KMAX + 1] = {0};
    int sms = 148;
};

struct TileSel { const int4* tiles; int ntiles, out_rb, out_re; };
struct TileArgs {
    int n_cells, total, g0, nloc, own_rb, own_re, out_rb, out_re, ntiles, K;
    const int4* tiles;     // {interior origin d0, p0, flags (TL_F_*), template code whose values every row of the tile shares}
    // device-chosen fusion depth (Jacobi): *kdev in 2..4 selects K and sel[K] at launch time (fct_kernels.cu sets it from
    // the sweep schedule of the solve); nullptr: the K, tiles, ntiles, out_* above
    const unsigned long long* kdev;
    TileSel sel[TL_KMAX];
    const int32_t* rowptr;
    const uint16_t* code;
    const unsigned long long* tdelta;
    const double* Lv;      // row-scaled low-order operator (zero diagonal slot)
    const double* b;       // b'
    const double* xin;     // x
    double* xout;          // x after K sweeps
    unsigned long long* jstate;
};

// ---- closed-form numbering ----------------------------------------------------------------------------------
__host__ __device__ __forceinline__ int tl_diag_len(int d, int n) { return (d <= n ? d : 2 * n - d) + 1; }
__host__ __device__ __forceinline__ int tl_diag_start(int d, int n, int total) {
    if (d <= n) return d * (d + 1) / 2;
    const int m = 2 * n - d;
    return total - (m + 1) * (m + 2) / 2;
}
// global row -> (d, pos)
__device__ __forceinline__ void tl_row_to_dp(int g, int n, int total, int& d, int& pos) {
    const int half = (n + 1) * (n + 2) / 2;
    const bool upper = g < half;
    const int q = upper ? g : total - 1 - g;
    int t = (int)((sqrt(8.0 * (double)q + 1.0) - 1.0) * 0.5);
    while (t * (t + 1) / 2 > q) --t;
    while ((t + 1) * (t + 2) / 2 <= q) ++t;
    const int p = q - t * (t + 1) / 2;
    if (upper) { d = t; pos = p; }
    else { d = 2 * n - t; pos = t - p; }      // the mirror reverses the order within a diagonal (len = t + 1)
}

// ---- set-up: neighbour deltas per template (doubles as the check that the pattern is a (d +- 1, pos +- 1) stencil) ----
__device__ __forceinline__ bool tl_row_deltas(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx, int r,
                                              int n, int total, int g0, unsigned long long& packed) {
    int d, pos;
    tl_row_to_dp(g0 + r, n, total, d, pos);
    const int k0 = rowptr[r], len = rowptr[r + 1] - k0;
    packed = 0ull;
    if (len > 8) return false;
    for (int j = 0; j < len; ++j) {
        int d2, p2;
        tl_row_to_dp(g0 + colidx[k0 + j], n, total, d2, p2);
        const int dd = d2 - d, dp = p2 - pos;
        if (dd < -1 || dd > 1 || dp < -1 || dp > 1) return false;
        const int delta = dd * TL_XS + dp;
        packed |= (unsigned long long)(unsigned char)(signed char)delta << (8 * j);
    }
    return true;
}
__global__ void k_tile_deltas(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                              const uint16_t* __restrict__ code, int nloc, int n, int total, int g0,
                              unsigned long long* __restrict__ tdelta, int* __restrict__ bad) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nloc) return;
    unsigned long long p;
    if (!tl_row_deltas(rowptr, colidx, r, n, total, g0, p)) { atomicAdd(bad, 1); return; }
    tdelta[code[r]] = p;          // rows of one template agree (checked by k_tile_deltas_verify)
}
__global__ void k_tile_deltas_verify(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                                     const uint16_t* __restrict__ code, int nloc, int n, int total, int g0,
                                     const unsigned long long* __restrict__ tdelta, int* __restrict__ bad) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nloc) return;
    unsigned long long p;
    if (!tl_row_deltas(rowptr, colidx, r, n, total, g0, p) || tdelta[code[r]] != p) atomicAdd(bad, 1);
}

// ---- the tile kernel -------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async8(void* dst_smem, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_mbar_arrive(uint64_t* bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// mbarrier wait that cannot hang the GPU: a wait that has not completed after ~2 s of spinning aborts the launch
// (cudaErrorLaunchFailure at the next synchronisation) instead of spinning forever
__device__ __forceinline__ void mbar_wait_bounded(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0;
    const uint32_t addr = smem_u32(bar);
    const long long t0 = clock64();
    while (!done) {
        asm volatile(
            "{\n"
            " .reg .pred p;\n"
            " mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            " selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
        if (!done && clock64() - t0 > 4000000000LL) __trap();
    }
}
__device__ __forceinline__ int tl_delta(unsigned long long pk, int j) {
    return (int)(signed char)(unsigned char)(pk >> (8 * j));
}
// the two neighbour layouts of interior rows (7 entries, ascending columns): above the main anti-diagonal (d < n) the
// neighbours sit at (d-1: pos-1, pos), (d: pos-1, pos, pos+1), (d+1: pos, pos+1); below it (d > n) at (d-1: pos, pos+1),
// (d: ...), (d+1: pos-1, pos).  Rows with one of these layouts take shared-memory loads with immediate offsets.
__host__ __device__ constexpr unsigned long long tl_pack7(int a0, int a1, int a2, int a3, int a4, int a5, int a6) {
    return (unsigned long long)(unsigned char)(signed char)a0 | ((unsigned long long)(unsigned char)(signed char)a1 << 8) |
           ((unsigned long long)(unsigned char)(signed char)a2 << 16) | ((unsigned long long)(unsigned char)(signed char)a3 << 24) |
           ((unsigned long long)(unsigned char)(signed char)a4 << 32) | ((unsigned long long)(unsigned char)(signed char)a5 << 40) |
           ((unsigned long long)(unsigned char)(signed char)a6 << 48);
}
#define TL_DPK_UP tl_pack7(-TL_XS - 1, -TL_XS, -1, 0, 1, TL_XS, TL_XS + 1)
#define TL_DPK_LO tl_pack7(-TL_XS, -TL_XS + 1, -1, 0, 1, TL_XS - 1, TL_XS)

template <int W>
struct TileSmem {
    static constexpr int SEGCAP = (TL_NQ * W + 2 + 1) & ~1;                        // staged CSR range of one region diagonal (even)
    static constexpr int L_DOUBLES = TL_NC * SEGCAP;
    static constexpr int X_DOUBLES = 3 * TL_ND * TL_XS;
    static constexpr size_t BYTES = 8 * (size_t)(L_DOUBLES + X_DOUBLES) + 4 * (2 * TL_NC + 4 * TL_ND);
    static_assert(L_DOUBLES % 2 == 0, "bulk-copy destinations must stay 16-byte aligned");
};

// local rows [ra, rb) of compute-set diagonal c (region diagonal c + 1) of the tile at (dlo, plo); empty: rb <= ra
__device__ __forceinline__ void tl_segment_rows(const TileArgs& a, int dlo, int plo, int c, int& ra, int& rb) {
    const int n = a.n_cells, d = dlo + c + 1;
    ra = 0; rb = 0;
    if (d < 0 || d > 2 * n) return;
    const int len = tl_diag_len(d, n), st = tl_diag_start(d, n, a.total) - a.g0;
    const int pa = max(plo + 1, 0), pb = min(plo + TL_NP - 1, len);
    if (pb <= pa) return;
    ra = max(st + pa, 0); rb = min(st + pb, a.nloc);
}

// The K dependent passes of one tile.  KIND 1 / 2: every row of the tile has the interior neighbour layout above / below the
// main anti-diagonal (shared-memory loads with immediate offsets); KIND 0: per-row deltas from the template table.
// Straight-line code over the thread's rows (no per-row branches: the 21 loads and the three FMA chains interleave); rows
// that are not part of pass s are computed on whatever their slots hold and simply not stored.
template <int W, int KIND>
__device__ __forceinline__ void tl_passes(int K, double* A, double* B, const int (&xi)[TL_R], const int (&ml)[TL_R],
                                          const int (&row)[TL_R], const double (&Lr)[TL_R][W],
                                          const unsigned long long (&dpk)[TL_R], const double (&br)[TL_R]) {
#pragma unroll 1
    for (int s = 1; s <= K; ++s) {
        const double* in = (s & 1) ? A : B;
        double* out = (s & 1) ? B : A;
        double acc[TL_R];
#pragma unroll
        for (int i = 0; i < TL_R; ++i) {
            const double* p = in + xi[i];
            double t = 0.0;
            if (KIND == 1) {
                t += Lr[i][0] * p[-TL_XS - 1]; t += Lr[i][1] * p[-TL_XS]; t += Lr[i][2] * p[-1]; t += Lr[i][3] * p[0];
                t += Lr[i][4] * p[1]; t += Lr[i][5] * p[TL_XS]; t += Lr[i][W - 1] * p[TL_XS + 1];
            } else if (KIND == 2) {
                t += Lr[i][0] * p[-TL_XS]; t += Lr[i][1] * p[-TL_XS + 1]; t += Lr[i][2] * p[-1]; t += Lr[i][3] * p[0];
                t += Lr[i][4] * p[1]; t += Lr[i][5] * p[TL_XS - 1]; t += Lr[i][W - 1] * p[TL_XS];
            } else {
#pragma unroll
                for (int jj = 0; jj < W; ++jj) t += Lr[i][jj] * p[tl_delta(dpk[i], jj)];
            }
            acc[i] = t;
        }
#pragma unroll
        for (int i = 0; i < TL_R; ++i) {
            const bool act = row[i] >= 0 && s <= ml[i];
            if (act) out[xi[i]] = br[i] - acc[i];
        }
        __syncthreads();
    }
}

// K Jacobi sweeps  x <- b' - sum_j l'_ij x_j  of the row-scaled low-order system (k_jacobi_sweep_tpl<., true>); the ChebSI
// counterpart is k_cheb_tile below.
// Every warp computes.  Software pipeline over the CTA's tiles (all loads are issued by the compute threads themselves):
//   during tile j   : the CSR bounds of tile j+2's diagonal segments and the template data of tile j+1 travel (registers);
//                     tile j+1's iterate (8-byte cp.async), matrix rows (one 1-D TMA bulk copy per region diagonal) and
//                     right-hand side (registers) are in flight, all shared-memory traffic completing on one mbarrier;
//   top of tile j+1 : matrix rows move from the staging buffer to registers, the buffer is handed to tile j+2.
// "Regular" tiles (flags from k_tile_classify: every row exists, 7 entries, one of the two interior layouts -- 93 % of the
// tiles at 4097^2) skip the per-row row-pointer / template look-ups altogether.
template <int W>
__global__ void __launch_bounds__(TL_NT, 1) k_tile(const TileArgs a) {
    if (*reinterpret_cast<volatile unsigned long long*>(a.jstate + 3)) return;      // already converged
    // launch geometry: fixed by the arguments, or selected by the device-side sweep schedule
    __shared__ TileSel ssel;
    __shared__ int sK;
    if (threadIdx.x == 0) {
        int k = a.K;
        TileSel t = {a.tiles, a.ntiles, a.out_rb, a.out_re};
        if (a.kdev) {
            k = (int)*reinterpret_cast<const volatile unsigned long long*>(a.kdev);
            k = k < 2 ? 2 : k > TL_KMAX - 1 ? TL_KMAX - 1 : k;
            t = a.sel[k];
        }
        ssel = t; sK = k;
    }
    __syncthreads();
    const int K = sK;
    const int4* const tiles_ = ssel.tiles;
    const int ntiles_ = ssel.ntiles;
    using SM = TileSmem<W>;
    double* sL = reinterpret_cast<double*>(fct_smem);
    double* sX = sL + SM::L_DOUBLES;
    int* sKa = reinterpret_cast<int*>(sX + SM::X_DOUBLES);          // [NC] first staged CSR index of each L segment
    int* sSh = sKa + TL_NC;                                          // [NC] staged offset of the segment's first row
    int* sRow = sSh + TL_NC;                                         // [2][ND] local row of (dl, pl = 0), per tile parity
    int* sLen = sRow + 2 * TL_ND;                                    // [2][ND] length of region diagonal dl (0: outside the mesh)
    __shared__ __align__(8) uint64_t bar_full;
    __shared__ double sred[2][TL_NT / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        mbar_init(&bar_full, TL_NT + TL_NT / 32);      // cp.async completions + one expect_tx arrival per warp
        mbar_fence_init();
    }
    const int n = a.n_cells, total = a.total;
    const int nmine = ((int)blockIdx.x < ntiles_) ? (ntiles_ - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    // tile records are fetched three tiles ahead into registers (rec1..rec3 = tiles j+1..j+3 while tile j is computed), so
    // that no thread ever waits for one
    auto fetch_rec = [&](int j) {
        return j < nmine ? __ldg(tiles_ + (int)blockIdx.x + j * (int)gridDim.x) : make_int4(0, 0, 0, 0);
    };
    int4 rec0 = fetch_rec(0), rec1 = fetch_rec(1), rec2 = fetch_rec(2), rec3 = fetch_rec(3);
    int recj = 0;                                    // rec0 is the record of tile recj
    auto tile_of = [&](int j) { return j == recj ? rec0 : j == recj + 1 ? rec1 : j == recj + 2 ? rec2 : rec3; };
    auto is_regular = [&](int flag) { return W == 7 && (flag & 3) != 0; };

    // rows this thread updates (compute set) and region entries it loads: fixed for the whole launch
    // (a warp takes 32 consecutive positions of one diagonal -- conflict-free 8-byte shared-memory accesses --, the
    // TL_NQ - 32 positions left over on each diagonal are dealt to the remaining slots)
    int xi[TL_R], ml[TL_R], dlr[TL_R], plr[TL_R];
#pragma unroll
    for (int i = 0; i < TL_R; ++i) {
        const int q = tid + i * TL_NT;
        const bool has = q < TL_NC * TL_NQ;
        int dc = 0, pc = 0;
        if (has) {
            if (q < TL_NC * 32) { dc = q >> 5; pc = q & 31; }
            else { const int t = q - TL_NC * 32; dc = t / (TL_NQ - 32); pc = 32 + t % (TL_NQ - 32); }
        }
        dlr[i] = dc + 1; plr[i] = pc + 1;
        xi[i] = dlr[i] * TL_XS + plr[i];
        ml[i] = has ? min(min(dlr[i], TL_ND - 1 - dlr[i]), min(plr[i], TL_NP - 1 - plr[i])) : 0;
    }
    constexpr int NE = (TL_ND * TL_NP + TL_NT - 1) / TL_NT;          // region entries per thread
    int edl[NE], epl[NE];
#pragma unroll
    for (int i = 0; i < NE; ++i) {
        const int e = tid + i * TL_NT;
        edl[i] = e < TL_ND * TL_NP ? e / TL_NP : -1;
        epl[i] = e < TL_ND * TL_NP ? e % TL_NP : 0;
    }
    // per-tile table of the region's diagonals (threads 0..ND-1), double-buffered by tile parity
    auto diag_table = [&](int j) {
        if (tid < TL_ND && j < nmine) {
            const int4 t = tile_of(j);
            const int d = t.x - K + tid;
            const bool in = d >= 0 && d <= 2 * n;
            sLen[(j & 1) * TL_ND + tid] = in ? tl_diag_len(d, n) : 0;
            sRow[(j & 1) * TL_ND + tid] = in ? tl_diag_start(d, n, total) + (t.y - K) - a.g0 : 0;
        }
    };
    // CSR bounds of the L segments this lane will issue for tile j (lanes 0..2 of each warp: segments warp + 16 lane)
    int sb_k0 = 0, sb_k1 = 0;
    auto load_seg_bounds = [&](int j) {
        sb_k0 = 0; sb_k1 = 0;
        if (j >= nmine) return;
        const int c = warp + (TL_NT / 32) * lane;
        if (lane < 3 && c < TL_NC) {
            const int4 t = tile_of(j);
            int ra, rb;
            tl_segment_rows(a, t.x - K, t.y - K, c, ra, rb);
            if (rb > ra) { sb_k0 = a.rowptr[ra]; sb_k1 = a.rowptr[rb]; }
        }
    };

    int row[TL_R], nrow[TL_R], ncode[TL_R], nk0[TL_R], nk1[TL_R];
    unsigned long long dpk[TL_R];
    double Lr[TL_R][W], br[TL_R], nb[TL_R];
    int nflag = 0;
    // rows of tile j and the first loads they need (right-hand side; for general tiles row pointer and template code)
    auto prefetch_meta = [&](int j) {
#pragma unroll
        for (int i = 0; i < TL_R; ++i) { nrow[i] = -1; ncode[i] = 0; nk0[i] = 0; nk1[i] = 0; nb[i] = 0.0; }
        nflag = 0;
        if (j >= nmine) return;
        const int4 t = tile_of(j);
        nflag = t.z;
        const bool reg = is_regular(nflag);
        const int plo = t.y - K;
        const int* tr = sRow + (j & 1) * TL_ND;
        const int* tl = sLen + (j & 1) * TL_ND;
#pragma unroll
        for (int i = 0; i < TL_R; ++i) {
            if (ml[i] == 0) continue;
            const int r = tr[dlr[i]] + plr[i];
            if ((unsigned)(plo + plr[i]) >= (unsigned)tl[dlr[i]] || (unsigned)r >= (unsigned)a.nloc) continue;
            nrow[i] = r;
            nb[i] = a.b[r];
            if (!reg) {
                ncode[i] = a.code[r];
                nk0[i] = a.rowptr[r]; nk1[i] = a.rowptr[r + 1];
            }
        }
    };
    // second-level loads (depend on the template code): the neighbour deltas of the rows of a general tile
    auto load_meta2 = [&]() {
        const bool reg = is_regular(nflag);
#pragma unroll
        for (int i = 0; i < TL_R; ++i) {
            dpk[i] = 0ull;
            if (nrow[i] < 0) continue;
            if (!reg) dpk[i] = __ldg(a.tdelta + ncode[i]);
        }
    };
    // the iterate of tile j's region -> iterate buffer j % 3; the caller guarantees that the buffer is free
    auto issue_x = [&](int j) {
        if (j < nmine) {
            const int plo = tile_of(j).y - K;
            const int* tr = sRow + (j & 1) * TL_ND;
            const int* tl = sLen + (j & 1) * TL_ND;
            double* xb = sX + (j % 3) * (TL_ND * TL_XS);
#pragma unroll
            for (int i = 0; i < NE; ++i) {
                if (edl[i] < 0) continue;
                const int r = tr[edl[i]] + epl[i];
                if ((unsigned)(plo + epl[i]) < (unsigned)tl[edl[i]] && (unsigned)r < (unsigned)a.nloc)
                    cp_async8(xb + edl[i] * TL_XS + epl[i], a.xin + r);
            }
        }
        cp_async_mbar_arrive(&bar_full);
    };
    // the matrix rows of tile j's diagonals -> staging buffer (bounds loaded one tile earlier by load_seg_bounds)
    auto issue_L = [&](int j) {
        uint32_t tx = 0;
        const int c = warp + (TL_NT / 32) * lane;
        if (lane < 3 && c < TL_NC && j < nmine) {
            int ka = 0, sh = 0;
            if (sb_k1 > sb_k0) {
                sh = (int)((reinterpret_cast<uintptr_t>(a.Lv + sb_k0) >> 3) & 1);
                ka = sb_k0 - sh;
                const uint32_t bytes = (uint32_t)(((sb_k1 - ka) + 1) & ~1) * 8u;
                tma_load_1d(sL + c * SM::SEGCAP, a.Lv + ka, bytes, &bar_full);
                tx = bytes;
            }
            sKa[c] = ka; sSh[c] = sh;
        }
        tx += __shfl_down_sync(0xffffffffu, tx, 1) + __shfl_down_sync(0xffffffffu, tx, 2);
        __syncwarp();                                          // the sKa / sSh stores of lanes 1, 2 are ordered before ...
        if (lane == 0) mbar_expect_tx(&bar_full, tx);          // ... the (releasing) arrival, one per warp
    };

    diag_table(0);
    diag_table(1);
    load_seg_bounds(0);
    __syncthreads();          // barrier initialised, diagonal tables of tiles 0 and 1 visible
    prefetch_meta(0);
    issue_x(0);
    issue_L(0);
    load_seg_bounds(1);
    load_meta2();
    double delta = 0.0, xa = 0.0;
    for (int j = 0; j < nmine; ++j) {
        const int flag = nflag;
        const bool reg = is_regular(flag);
        int k0r[TL_R], lenr[TL_R];
#pragma unroll
        for (int i = 0; i < TL_R; ++i) { row[i] = nrow[i]; k0r[i] = nk0[i]; lenr[i] = nk1[i] - nk0[i]; br[i] = nb[i]; }
        mbar_wait_bounded(&bar_full, (uint32_t)(j & 1));
        // staged matrix rows -> registers
#pragma unroll
        for (int i = 0; i < TL_R; ++i) {
            const int c = dlr[i] - 1;
            if (reg) {
                const double* p = sL + c * SM::SEGCAP + sSh[c] + W * (plr[i] - 1);
#pragma unroll
                for (int jj = 0; jj < W; ++jj) Lr[i][jj] = (row[i] >= 0) ? p[jj] : 0.0;
            } else {
#pragma unroll
                for (int jj = 0; jj < W; ++jj) Lr[i][jj] = 0.0;
                if (row[i] < 0) continue;
                const double* p = sL + c * SM::SEGCAP + (k0r[i] - sKa[c]);
#pragma unroll
                for (int jj = 0; jj < W; ++jj) Lr[i][jj] = (jj < lenr[i]) ? p[jj] : 0.0;
            }
        }
        if (j > 0) { rec0 = rec1; rec1 = rec2; rec2 = rec3; rec3 = fetch_rec(j + 3); recj = j; }
        diag_table(j + 2);          // parity of tile j, whose table nobody reads any more (its x, rows were set up one tile ago)
        __syncthreads();            // staging buffer and third iterate buffer are free; table of tile j+1 is complete
        prefetch_meta(j + 1);
        issue_x(j + 1);
        issue_L(j + 1);
        load_seg_bounds(j + 2);
        double* A = sX + (j % 3) * (TL_ND * TL_XS);
        double* B = sX + ((j + 2) % 3) * (TL_ND * TL_XS);
        if (reg && (flag & TL_F_UP)) tl_passes<W, 1>(K, A, B, xi, ml, row, Lr, dpk, br);
        else if (reg) tl_passes<W, 2>(K, A, B, xi, ml, row, Lr, dpk, br);
        else tl_passes<W, 0>(K, A, B, xi, ml, row, Lr, dpk, br);
        // the next tile's template data travels while this tile's interior is written back
        load_meta2();
        const double* fin = (K & 1) ? B : A;
        const double* prev = (K & 1) ? A : B;
        double fv[TL_R], pv[TL_R];          // all loads first (the slots always exist), then the predicated stores
#pragma unroll
        for (int i = 0; i < TL_R; ++i) { fv[i] = fin[xi[i]]; pv[i] = prev[xi[i]]; }
#pragma unroll
        for (int i = 0; i < TL_R; ++i) {
            if (row[i] >= 0 && ml[i] >= K && row[i] >= ssel.out_rb && row[i] < ssel.out_re) {
                const double v = fv[i];
                a.xout[row[i]] = v;
                if (row[i] >= a.own_rb && row[i] < a.own_re) {
                    delta = fmax(delta, fabs(v - pv[i]));
                    xa = fmax(xa, fabs(v));
                }
            }
        }
    }
    // stopping test input: ||x_K - x_{K-1}||_inf and ||x_K||_inf over the owned rows (k_jacobi_sweep_tpl's `check`)
    delta = warp_max(delta);
    xa = warp_max(xa);
    if (lane == 0) { sred[0][warp] = delta; sred[1][warp] = xa; }
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < TL_NT / 32; ++w) { delta = fmax(delta, sred[0][w]); xa = fmax(xa, sred[1][w]); }
        atomicMax(a.jstate + 0, (unsigned long long)__double_as_longlong(delta));
        atomicMax(a.jstate + 1, (unsigned long long)__double_as_longlong(xa));
        if (blockIdx.x == 0) atomicAdd(a.jstate + 4, (unsigned long long)K);
    }
}

// Classifies the tiles of a list from the data: one warp per tile walks the compute set (the rows that are ever updated)
// and checks that every row exists, has 7 entries and the neighbour layout TL_DPK_UP (or TL_DPK_LO) -- then the kernel may
// skip the per-row look-ups -- and whether all of them carry the same mass-matrix values (then ChebSI reads them once).
__global__ void k_tile_classify(int4* __restrict__ tiles, int ntiles, int K, int n, int total, int g0, int nloc,
                                const int32_t* __restrict__ rowptr, const uint16_t* __restrict__ code,
                                const unsigned long long* __restrict__ tdelta, const double* __restrict__ tval) {
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= ntiles) return;
    int4 t = tiles[w];
    const int dlo = t.x - K, plo = t.y - K;
    bool all = true, up = true, lo = true, muni = true;
    int vcode = -1;
    {   // reference row for the values: the first compute-set row
        const int d = dlo + 1, pos = plo + 1;
        if (d >= 0 && d <= 2 * n && pos >= 0 && pos < tl_diag_len(d, n)) {
            const int r = tl_diag_start(d, n, total) + pos - g0;
            if (r >= 0 && r < nloc) vcode = code[r];
        }
    }
    for (int q = lane; q < TL_NC * TL_NQ; q += 32) {
        const int d = dlo + 1 + q / TL_NQ, pos = plo + 1 + q % TL_NQ;
        int r = -1;
        if (d >= 0 && d <= 2 * n && pos >= 0 && pos < tl_diag_len(d, n)) r = tl_diag_start(d, n, total) + pos - g0;
        if (r < 0 || r >= nloc) { all = false; continue; }
        if (rowptr[r + 1] - rowptr[r] != 7) all = false;
        const int cd = code[r];
        const unsigned long long pk = tdelta[cd];
        up = up && pk == TL_DPK_UP;
        lo = lo && pk == TL_DPK_LO;
        if (vcode >= 0)
            for (int jj = 0; jj < FCT_TPL_W; ++jj)
                muni = muni && __double_as_longlong(tval[FCT_TPL_W * cd + jj]) == __double_as_longlong(tval[FCT_TPL_W * vcode + jj]);
    }
    all = __all_sync(0xffffffffu, all); up = __all_sync(0xffffffffu, up); lo = __all_sync(0xffffffffu, lo);
    muni = __all_sync(0xffffffffu, muni);
    if (lane == 0) {
        t.z = (all && up ? TL_F_UP : 0) | (all && lo ? TL_F_LO : 0) | (all && muni && vcode >= 0 ? TL_F_MUNI : 0);
        t.w = vcode >= 0 ? vcode : 0;
        tiles[w] = t;
    }
}

// ======================================================================================================
// host side
// ======================================================================================================
struct fct_cheb_tiles;
void fct_cheb_tiles_free(fct_ctx* ctx);
static int cheb_tiles_prepare(fct_ctx* ctx);

void fct_tiles_free(fct_ctx* ctx) {
    fct_cheb_tiles_free(ctx);
    fct_tiles* t = ctx->tiles;
    if (!t) return;
    cudaFree(t->tdelta);
    for (int k = 0; k <= TL_KMAX; ++k) cudaFree(t->list[k]);
    delete t;
    ctx->tiles = nullptr;
}

// tiles (interior (TL_ND - 2K) diagonals x (TL_NP - 2K) positions) covering the owned global rows [ga, gb), diagonal block by
// diagonal block; host code
static void tile_list_host(int n, long long ga, long long gb, int K, std::vector<int4>& v) {
    const int total = (n + 1) * (n + 1);
    const int Td = TL_ND - 2 * K, Tp = TL_NP - 2 * K;
    v.clear();
    if (gb <= ga) return;
    int dA = 0, dB = 2 * n;
    while (dA < 2 * n && tl_diag_start(dA + 1, n, total) <= ga) ++dA;
    while (dB > 0 && tl_diag_start(dB, n, total) >= gb) --dB;
    for (int d0 = dA; d0 <= dB; d0 += Td) {
        int maxlen = 0;
        for (int d = d0; d < d0 + Td && d <= 2 * n; ++d) maxlen = std::max(maxlen, tl_diag_len(d, n));
        for (int p0 = 0; p0 < maxlen; p0 += Tp) {
            bool any = false;
            for (int d = d0; d < d0 + Td && d <= dB && !any; ++d) {
                const long long s = tl_diag_start(d, n, total);
                const long long lo = std::max(s + p0, ga), hi = std::min(s + std::min(p0 + Tp, tl_diag_len(d, n)), gb);
                any = hi > lo;
            }
            if (any) v.push_back(make_int4(d0, p0, 0, 0));
        }
    }
}

// test hook (CPU-testable, no device needed): the tile list of a row block and the tile geometry constants
extern "C" int fct_debug_tile_list(int32_t n_cells, int64_t g0, int32_t row_begin, int32_t row_end, int32_t K,
                                   int32_t* d0p0_out, int32_t cap, int32_t* count_out, int32_t* geom_out /* ND, NP */) {
    FCT_CHECK(n_cells >= 1 && K >= 1 && 2 * K < TL_NP && count_out, "fct_debug_tile_list: bad argument");
    std::vector<int4> v;
    tile_list_host(n_cells, g0 + row_begin, g0 + row_end, K, v);
    *count_out = (int32_t)v.size();
    if (geom_out) { geom_out[0] = TL_ND; geom_out[1] = TL_NP; }
    if (d0p0_out)
        for (int i = 0; i < (int)v.size() && i < cap; ++i) { d0p0_out[2 * i] = v[i].x; d0p0_out[2 * i + 1] = v[i].y; }
    return 0;
}

// Rows a launch of K fused passes writes: with halo depth D > K (multi-GPU) ring D-K -- every row whose K-ring lies inside the
// local range -- so that a second launch can follow without an exchange; else the owned rows.  The stopping test of the
// Jacobi sweeps always runs on the owned rows only.
static inline void tile_out_range(const fct_ctx* ctx, int K, int* rb, int* re) {
    const int j = ctx->depth > K ? ctx->depth - K : 0;
    *rb = ctx->ring_lo[j];
    *re = ctx->ring_hi[j];
}

static int build_tile_list(fct_ctx* ctx, int K) {
    fct_tiles* t = ctx->tiles;
    if (t->list[K]) return 0;
    std::vector<int4> v;
    int orb, ore;
    tile_out_range(ctx, K, &orb, &ore);
    tile_list_host(t->n_cells, (long long)t->g0 + orb, (long long)t->g0 + ore, K, v);
    t->count[K] = (int)v.size();
    if (v.empty()) { FCT_CUDA(cudaMalloc((void**)&t->list[K], sizeof(int4))); return 0; }
    FCT_CUDA(cudaMalloc((void**)&t->list[K], sizeof(int4) * v.size()));
    FCT_CUDA(cudaMemcpy(t->list[K], v.data(), sizeof(int4) * v.size(), cudaMemcpyHostToDevice));
    const int n = t->n_cells;
    k_tile_classify<<<((int)v.size() * 32 + 255) / 256, 256, 0, ctx->stream>>>(t->list[K], (int)v.size(), K, n, (n + 1) * (n + 1),
                                                                            t->g0, ctx->n, ctx->rowptr, ctx->tpl_code, t->tdelta,
                                                                            ctx->tpl_val);
    ctx->launches++;
    FCT_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}

template <int W>
static int tile_set_attr() {
    FCT_CUDA(cudaFuncSetAttribute(k_tile<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TileSmem<W>::BYTES));
    return 0;
}
template <int W>
static void tile_launch_t(fct_ctx* ctx, const TileArgs& a, int grid) {
    k_tile<W><<<grid, TL_THREADS, TileSmem<W>::BYTES, ctx->stream>>>(a);
}

static int tiles_configure() {
    static bool done = false;
    if (done) return 0;
    if (tile_set_attr<7>() || tile_set_attr<8>()) return 1;
    done = true;
    return 0;
}

// (Re)build what the tile kernels need; called when the structured numbering is declared and whenever the row templates
// are rebuilt.  Never fails the caller: without tiles the per-pass kernels run.
int fct_tiles_prepare(fct_ctx* ctx) {
    fct_tiles* t = ctx->tiles;
    if (!t) return 0;
    cudaFree(t->tdelta);
    t->tdelta = nullptr;
    ctx->tiles_ok = false;
    ctx->cheb_tiles_ok = false;
    const char* e = getenv("FCT_NO_TILES");
    if (e && atoi(e) == 1) return 0;
    if (ctx->tpl_count <= 0 || ctx->jac_mode != 2 || ctx->max_row > 8) return 0;
    const int n = t->n_cells, total = (n + 1) * (n + 1);
    if ((long long)t->g0 + ctx->n > total) return 0;
    int* bad = nullptr;
    int hbad = 1;
    do {
        if (tiles_configure()) break;
        if (cudaMalloc((void**)&t->tdelta, sizeof(unsigned long long) * (size_t)ctx->tpl_count) != cudaSuccess) break;
        if (cudaMalloc((void**)&bad, sizeof(int)) != cudaSuccess) break;
        cudaMemsetAsync(bad, 0, sizeof(int), ctx->stream);
        cudaMemsetAsync(t->tdelta, 0, sizeof(unsigned long long) * (size_t)ctx->tpl_count, ctx->stream);
        const int nb = (ctx->n + 255) / 256;
        k_tile_deltas<<<nb, 256, 0, ctx->stream>>>(ctx->rowptr, ctx->colidx, ctx->tpl_code, ctx->n, n, total, t->g0, t->tdelta, bad);
        k_tile_deltas_verify<<<nb, 256, 0, ctx->stream>>>(ctx->rowptr, ctx->colidx, ctx->tpl_code, ctx->n, n, total, t->g0,
                                                          t->tdelta, bad);
        ctx->launches += 2;
        if (cudaMemcpyAsync(&hbad, bad, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess) break;
        if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) break;
    } while (0);
    cudaFree(bad);
    cudaGetLastError();
    for (int k = 0; k <= TL_KMAX; ++k) { cudaFree(t->list[k]); t->list[k] = nullptr; t->count[k] = 0; }
    if (hbad != 0) { cudaFree(t->tdelta); t->tdelta = nullptr; return 0; }
    cudaDeviceGetAttribute(&t->sms, cudaDevAttrMultiProcessorCount, ctx->device);
    // all tile lists now: the first solve may already run inside a stream capture (CUDA-graph WHILE loop of the Jacobi
    // solve), where allocations and synchronous copies are not allowed
    for (int k = 2; k <= TL_KMAX; ++k)
        if (build_tile_list(ctx, k)) { cudaGetLastError(); return 0; }
    ctx->tiles_ok = true;
    return cheb_tiles_prepare(ctx);
}

// Declares that local row i is DoF g0 + i of dolfin's CG1 numbering on RectangleMesh(n_cells x n_cells, "right")
// (fct_mesh_rect_build).  The pattern is verified on the device; on a mismatch the context simply keeps the per-pass kernels.
extern "C" int fct_ctx_set_rect(fct_ctx* ctx, int32_t n_cells, int64_t g0) {
    FCT_CHECK(ctx && n_cells >= 1 && g0 >= 0, "fct_ctx_set_rect: bad argument");
    FCT_CHECK(n_cells <= 23000, "fct_ctx_set_rect: mesh too large for 32-bit row arithmetic");
    fct_tiles_free(ctx);
    ctx->tiles = new fct_tiles();
    ctx->tiles->n_cells = n_cells;
    ctx->tiles->g0 = (int)g0;
    return fct_tiles_prepare(ctx);
}

extern "C" int fct_tiles_active(fct_ctx* ctx, int32_t* active) {
    FCT_CHECK(ctx && active, "fct_tiles_active: null argument");
    *active = ctx->tiles_ok ? 1 : 0;
    return 0;
}

static int fill_args(fct_ctx* ctx, int K, TileArgs& a) {
    fct_tiles* t = ctx->tiles;
    FCT_CHECK(t->list[K], "tile list for K=%d was not built", K);
    a.n_cells = t->n_cells; a.total = (t->n_cells + 1) * (t->n_cells + 1); a.g0 = t->g0; a.nloc = ctx->n;
    a.own_rb = ctx->row_begin; a.own_re = ctx->row_end; a.ntiles = t->count[K]; a.tiles = t->list[K]; a.K = K;
    tile_out_range(ctx, K, &a.out_rb, &a.out_re);
    a.rowptr = ctx->rowptr; a.code = ctx->tpl_code; a.tdelta = t->tdelta;
    a.Lv = nullptr; a.b = nullptr; a.xin = nullptr; a.xout = nullptr;
    a.jstate = ctx->jstate;
    a.kdev = nullptr;
    for (int k = 0; k < TL_KMAX; ++k) {
        a.sel[k] = TileSel{nullptr, 0, 0, 0};
        if (k >= 2 && t->list[k]) {
            a.sel[k].tiles = t->list[k]; a.sel[k].ntiles = t->count[k];
            tile_out_range(ctx, k, &a.sel[k].out_rb, &a.sel[k].out_re);
        }
    }
    return 0;
}

// K (2..4) Jacobi sweeps of the row-scaled low-order system in one launch: xout <- sweep^K(xin); accumulates the
// stopping-test maxima of the last sweep and adds K to the sweep counter
// kdev != nullptr: the launch runs *kdev (2..K) sweeps instead, read on the device (sweep schedule of the solve)
int fct_tile_jacobi(fct_ctx* ctx, int K, const double* Lv, const double* b, const double* xin, double* xout,
                    const unsigned long long* kdev) {
    FCT_CHECK(ctx->tiles_ok && K >= 2 && K <= 4, "fct_tile_jacobi: not available");
    TileArgs a;
    if (fill_args(ctx, K, a)) return 1;
    a.Lv = Lv; a.b = b; a.xin = xin; a.xout = xout;
    a.kdev = kdev;
    int most = a.ntiles;                 // the deepest fusion has the smallest interiors, hence the most tiles
    if (kdev)
        for (int k = 2; k <= K; ++k) most = a.sel[k].ntiles > most ? a.sel[k].ntiles : most;
    int grid = most < ctx->tiles->sms ? most : ctx->tiles->sms;
    if (ctx->tile_grid_cap > 0 && grid > ctx->tile_grid_cap) grid = ctx->tile_grid_cap;
    if (grid <= 0) return 0;
    if (ctx->max_row <= 7) tile_launch_t<7>(ctx, a, grid); else tile_launch_t<8>(ctx, a, grid);
    ctx->launches++;
    return 0;
}

// ======================================================================================================
// ChebSI tiles
// ======================================================================================================
// ChebSI (helpers.py:143-185) applies the STATIC mass matrix: away from the boundary every row carries the same seven
// values, and its neighbour layout depends on the diagonal only: in diagonal d-1 the row's neighbours start at pos-1
// (d <= n) or pos (d > n), in diagonal d+1 at pos (d < n) or pos-1 (d >= n).  So a ChebSI tile needs no per-row matrix data
// at all and can be much larger than a Jacobi tile: a CTA owns a 50 x 98 region, every thread three groups of three
// consecutive positions (9 rows); a group shares its 13 neighbour loads (4 + 5 + 4 instead of 3 x 7, lane stride 3
// doubles: conflict-free), the seven matrix values and diag(M) sit in registers for the whole launch, y_old of the
// thread's rows in registers.  Rows that do not follow the nominal pattern (boundary rows, truncated halo rows of a
// multi-GPU block: found by k_cheb_classify from the data, at most CT_MAXEXC per tile) are recomputed after each pass by
// one thread each through the generic template path and overwrite the nominal result -- same arithmetic as
// k_cheb_iter_tpl in both cases, so the result is bit-identical to the per-iteration kernel.
#define CT_ND 50
#define CT_NP 98
#define CT_NC (CT_ND - 2)
#define CT_NQ (CT_NP - 2)
#define CT_XS CT_NP
#define CT_NT 512
#define CT_RD 3                       // region diagonals per warp: CT_NC == CT_RD * (CT_NT / 32)
#define CT_RP 3                       // consecutive positions per lane: CT_NQ == CT_RP * 32
#define CT_MAXEXC CT_NT               // one exceptional row per thread
static_assert(CT_NC == CT_RD * (CT_NT / 32) && CT_NQ == CT_RP * 32, "ChebSI tile geometry");

struct ChebTileArgs {
    int n_cells, total, g0, nloc, own_rb, own_re, out_rb, out_re, ntiles, K;
    const int4* tiles;          // {d0, p0, exceptions, template code carrying the nominal values}
    const int* exc_off;         // [ntiles] first exception descriptor of each tile
    const int4* exc;            // {region index, template code, delta bytes 0..3, delta bytes 4..7}
    const double* tval;
    const double* tdiag;
    const double* g;
    const double* ymid;
    const double* yold;         // nullptr: zero
    double* ymid_out;
    double* yold_out;           // nullptr: not wanted
    double om[TL_KMAX];
    double dscale;
};

struct ChebSmem {
    static constexpr int X_DOUBLES = 3 * CT_ND * CT_XS;
    static constexpr int G_DOUBLES = CT_NC * CT_NQ;
    static constexpr size_t BYTES = 8 * (size_t)(X_DOUBLES + 2 * G_DOUBLES) + 4 * (4 * CT_ND);
};

// neighbour offsets of the nominal layout on diagonal d: first neighbour in d-1 at pos + ou, in d+1 at pos + od
__host__ __device__ __forceinline__ int ct_ou(int d, int n) { return d <= n ? -1 : 0; }
__host__ __device__ __forceinline__ int ct_od(int d, int n) { return d < n ? 0 : -1; }

// x / c for a divisor whose correctly rounded reciprocal rc = 1/c is at hand: q = RN(x rc), r = x - q c (exact, FMA),
// RN(q + r rc) is the correctly rounded quotient (Markstein's division step -- what the compiler's own DDIV sequence ends
// with), i.e. the same bits as x / c in three operations instead of ~35.  The per-iteration kernel divides, and the
// bit-identity tests of the two (single GPU, and N ranks against one) cover hundreds of millions of quotients.
// Range: the step is exact while r is representable, i.e. for |x| >= 2^-960 or so (and x = 0 gives 0); below that -- values
// of order 1e-290, far under anything a state or adjoint of these problems carries -- the result may differ from x / c in
// the last bit.  A range test in the loop (branch per row or per group of rows) was measured at +0.25 .. +0.4 ms per
// ChebSI (it breaks the interleaving of the rows' dependent chains), so it is not there; FCT_TILE_KC=0 selects the
// per-iteration kernel, which divides.
__device__ __forceinline__ double ct_div(double x, double c, double rc) {
    const double q = x * rc;
    const double r = fma(-q, c, x);
    return fma(r, rc, q);
}

__global__ void __launch_bounds__(CT_NT, 1) k_cheb_tile(const ChebTileArgs a) {
    double* sX = reinterpret_cast<double*>(fct_smem);
    double* sG = sX + ChebSmem::X_DOUBLES;                   // staged g of the next tile's compute set
    double* sY = sG + ChebSmem::G_DOUBLES;                   // staged y_old
    int* sRow = reinterpret_cast<int*>(sY + ChebSmem::G_DOUBLES);      // [2][ND] local row of (dl, pl = 0), per tile parity
    int* sLen = sRow + 2 * CT_ND;                            // [2][ND] diagonal length (0: outside the mesh)
    __shared__ __align__(8) uint64_t bar_full;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) { mbar_init(&bar_full, CT_NT); mbar_fence_init(); }
    const int n = a.n_cells, total = a.total, K = a.K;
    const int nmine = ((int)blockIdx.x < a.ntiles) ? (a.ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    auto fetch_rec = [&](int j) {
        return j < nmine ? __ldg(a.tiles + (int)blockIdx.x + j * (int)gridDim.x) : make_int4(0, 0, 0, 0);
    };
    int4 rec0 = fetch_rec(0), rec1 = fetch_rec(1), rec2 = fetch_rec(2), rec3 = fetch_rec(3);
    int recj = 0;
    auto tile_of = [&](int j) { return j == recj ? rec0 : j == recj + 1 ? rec1 : j == recj + 2 ? rec2 : rec3; };
    auto diag_table = [&](int j) {
        if (tid < CT_ND && j < nmine) {
            const int4 t = tile_of(j);
            const int d = t.x - K + tid;
            const bool in = d >= 0 && d <= 2 * n;
            sLen[(j & 1) * CT_ND + tid] = in ? tl_diag_len(d, n) : 0;
            sRow[(j & 1) * CT_ND + tid] = in ? tl_diag_start(d, n, total) + (t.y - K) - a.g0 : 0;
        }
    };
    // the thread's rows: region diagonals dl = warp + 16 k + 1 (k < 3), positions pl = 3 lane + 1 + m (m < 3)
    const int pl0 = CT_RP * lane + 1;
    int mlp[CT_RP];
#pragma unroll
    for (int m = 0; m < CT_RP; ++m) mlp[m] = min(pl0 + m, CT_NP - 1 - (pl0 + m));
    // the seven nominal matrix values and dscale * diag(M): loaded per tile from its value template (the same bits for every tile
    // of a uniform mesh, but nothing here assumes that)
    double Mv[7], mdn = 1.0, rmdn = 1.0;

    auto issue_loads = [&](int j) {
        if (j < nmine) {
            const int plo = tile_of(j).y - K;
            const int* tr = sRow + (j & 1) * CT_ND;
            const int* tl = sLen + (j & 1) * CT_ND;
            double* xb = sX + (j % 3) * (CT_ND * CT_XS);
            for (int e = tid; e < CT_ND * CT_NP; e += CT_NT) {
                const int dl = e / CT_NP, pl = e - dl * CT_NP;
                const int r = tr[dl] + pl;
                if ((unsigned)(plo + pl) < (unsigned)tl[dl] && (unsigned)r < (unsigned)a.nloc) cp_async8(xb + e, a.ymid + r);
            }
            for (int q = tid; q < CT_NC * CT_NQ; q += CT_NT) {
                const int c = q / CT_NQ, pc = q - c * CT_NQ;
                const int r = tr[c + 1] + pc + 1;
                if ((unsigned)(plo + pc + 1) < (unsigned)tl[c + 1] && (unsigned)r < (unsigned)a.nloc) {
                    cp_async8(sG + q, a.g + r);
                    if (a.yold) cp_async8(sY + q, a.yold + r);
                }
            }
        }
        cp_async_mbar_arrive(&bar_full);
    };

    diag_table(0);
    diag_table(1);
    __syncthreads();
    issue_loads(0);
    for (int j = 0; j < nmine; ++j) {
        const int4 t = tile_of(j);
        const int dlo = t.x - K, plo = t.y - K, nexc = t.z;
        const int* tr = sRow + (j & 1) * CT_ND;
        const int* tl = sLen + (j & 1) * CT_ND;
        // nominal values of this tile
#pragma unroll
        for (int jj = 0; jj < 7; ++jj) Mv[jj] = __ldg(a.tval + FCT_TPL_W * t.w + jj);
        mdn = a.dscale * __ldg(a.tdiag + t.w);
        rmdn = 1.0 / mdn;
        // rows, existence, pass limits, neighbour offsets of the thread's three diagonals
        int rowb[CT_RD], exm[CT_RD], mld[CT_RD], ou[CT_RD], od[CT_RD], xi[CT_RD];
#pragma unroll
        for (int k = 0; k < CT_RD; ++k) {
            const int dl = warp + (CT_NT / 32) * k + 1, d = dlo + dl;
            xi[k] = dl * CT_XS + pl0;
            mld[k] = min(dl, CT_ND - 1 - dl);
            ou[k] = ct_ou(d, n); od[k] = ct_od(d, n);
            rowb[k] = tr[dl] + pl0;
            int msk = 0;
#pragma unroll
            for (int m = 0; m < CT_RP; ++m)
                if ((unsigned)(plo + pl0 + m) < (unsigned)tl[dl] && (unsigned)(rowb[k] + m) < (unsigned)a.nloc) msk |= 1 << m;
            exm[k] = msk;
        }
        // this thread's exceptional row of the tile (boundary rows, truncated halo rows): generic template arithmetic.  Set up
        // here, before the barrier below: tile j's diagonal table is recycled for tile j+2 from then on
        int e_idx = -1, e_ml = 0, e_row = 0;
        unsigned long long e_dpk = 0ull;
        double ev[FCT_TPL_W], e_g = 0.0, e_yo = 0.0, e_md = 1.0, e_rmd = 1.0;
#pragma unroll
        for (int jj = 0; jj < FCT_TPL_W; ++jj) ev[jj] = 0.0;
        if (tid < nexc) {
            const int4 ed = __ldg(a.exc + a.exc_off[(int)blockIdx.x + j * (int)gridDim.x] + tid);
            e_idx = ed.x;
            e_dpk = (unsigned long long)(unsigned)ed.z | ((unsigned long long)(unsigned)ed.w << 32);
            const int edl = e_idx / CT_XS, epl = e_idx - edl * CT_XS;
            e_ml = min(min(edl, CT_ND - 1 - edl), min(epl, CT_NP - 1 - epl));
            e_row = tr[edl] + epl;
#pragma unroll
            for (int jj = 0; jj < FCT_TPL_W; ++jj) ev[jj] = __ldg(a.tval + FCT_TPL_W * ed.y + jj);
            e_md = a.dscale * __ldg(a.tdiag + ed.y);
            e_rmd = 1.0 / e_md;
            e_g = a.g[e_row];
            if (a.yold) e_yo = a.yold[e_row];
        }
        mbar_wait_bounded(&bar_full, (uint32_t)(j & 1));
        double g[CT_RD][CT_RP], yo[CT_RD][CT_RP];
#pragma unroll
        for (int k = 0; k < CT_RD; ++k)
#pragma unroll
            for (int m = 0; m < CT_RP; ++m) {
                const int q = (warp + (CT_NT / 32) * k) * CT_NQ + CT_RP * lane + m;
                const bool ex = (exm[k] >> m) & 1;
                g[k][m] = ex ? sG[q] : 0.0;
                yo[k][m] = (ex && a.yold) ? sY[q] : 0.0;
            }
        if (j > 0) { rec0 = rec1; rec1 = rec2; rec2 = rec3; rec3 = fetch_rec(j + 3); recj = j; }
        __syncthreads();            // staging buffers and the third iterate buffer are free; table of tile j+1 is complete;
                                    // every warp has taken what it needs from tile j's diagonal table ...
        diag_table(j + 2);          // ... whose slot is recycled now (read from the next barrier (A) on)
        issue_loads(j + 1);
        double* A = sX + (j % 3) * (CT_ND * CT_XS);
        double* B = sX + ((j + 2) % 3) * (CT_ND * CT_XS);
#pragma unroll 1
        for (int s = 1; s <= K; ++s) {
            const double* in = (s & 1) ? A : B;
            double* out = (s & 1) ? B : A;
            const double om = a.om[s - 1];
#pragma unroll
            for (int k = 0; k < CT_RD; ++k) {
                const double* pm = in + xi[k];
                const double* pu = pm - CT_XS + ou[k];
                const double* pd = pm + CT_XS + od[k];
                double u[CT_RP + 1], c[CT_RP + 2], w[CT_RP + 1];
#pragma unroll
                for (int m = 0; m < CT_RP + 1; ++m) { u[m] = pu[m]; w[m] = pd[m]; }
#pragma unroll
                for (int m = 0; m < CT_RP + 2; ++m) c[m] = pm[m - 1];
#pragma unroll
                for (int m = 0; m < CT_RP; ++m) {
                    double acc = 0.0;
                    acc += Mv[0] * u[m]; acc += Mv[1] * u[m + 1]; acc += Mv[2] * c[m]; acc += Mv[3] * c[m + 1];
                    acc += Mv[4] * c[m + 2]; acc += Mv[5] * w[m]; acc += Mv[6] * w[m + 1];
                    const double ym = c[m + 1];
                    const double z = ct_div(g[k][m] - acc, mdn, rmdn);
                    const double yn = om * (z + ym - yo[k][m]) + yo[k][m];
                    if (((exm[k] >> m) & 1) && s <= min(mld[k], mlp[m])) { out[xi[k] + m] = yn; yo[k][m] = ym; }
                }
            }
            if (nexc > 0) {         // tile-uniform
                __syncthreads();
                if (e_idx >= 0 && s <= e_ml) {
                    const double* p = in + e_idx;
                    double acc = 0.0;
#pragma unroll
                    for (int jj = 0; jj < FCT_TPL_W; ++jj) acc += ev[jj] * p[tl_delta(e_dpk, jj)];
                    const double ym = p[0];
                    const double z = ct_div(e_g - acc, e_md, e_rmd);
                    out[e_idx] = om * (z + ym - e_yo) + e_yo;
                    e_yo = ym;
                }
            }
            __syncthreads();
        }
        const double* fin = (K & 1) ? B : A;
#ifndef CT_NOHOIST
        double fv[CT_RD][CT_RP];          // all loads first (the slots always exist), then the predicated stores
#pragma unroll
        for (int k = 0; k < CT_RD; ++k)
#pragma unroll
            for (int m = 0; m < CT_RP; ++m) fv[k][m] = fin[xi[k] + m];
#endif
#pragma unroll
        for (int k = 0; k < CT_RD; ++k)
#pragma unroll
            for (int m = 0; m < CT_RP; ++m) {
                const int r = rowb[k] + m;
                if (((exm[k] >> m) & 1) && min(mld[k], mlp[m]) >= K && r >= a.out_rb && r < a.out_re) {
#ifndef CT_NOHOIST
                    a.ymid_out[r] = fv[k][m];
#else
                    a.ymid_out[r] = fin[xi[k] + m];
#endif
                    if (a.yold_out) a.yold_out[r] = yo[k][m];
                }
            }
    }
}

// Classification of the ChebSI tiles from the data (one warp per tile; pass 0 counts, pass 1 fills the descriptors):
// a row of the compute set is nominal when it has 7 entries at the nominal offsets of its diagonal and the values of the
// tile's value template (the first such row found); every other existing row becomes an exception descriptor.
__global__ void k_cheb_classify(int pass, int4* __restrict__ tiles, int ntiles, int K, int n, int total, int g0, int nloc,
                                const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                                const uint16_t* __restrict__ code, const double* __restrict__ tval,
                                const int* __restrict__ exc_off, int4* __restrict__ exc) {
    const int wdx = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (wdx >= ntiles) return;
    int4 t = tiles[wdx];
    const int dlo = t.x - K, plo = t.y - K;
    // row -> (exists, nominal layout, deltas in this tile geometry)
    auto inspect = [&](int q, int& r, unsigned long long& pk, bool& layout_ok) {
        const int dl = q / CT_NQ + 1, pl = q % CT_NQ + 1;
        const int d = dlo + dl, pos = plo + pl;
        r = -1; pk = 0ull; layout_ok = false;
        if (d < 0 || d > 2 * n || pos < 0 || pos >= tl_diag_len(d, n)) return;
        const int rr = tl_diag_start(d, n, total) + pos - g0;
        if (rr < 0 || rr >= nloc) return;
        r = rr;
        const int k0 = rowptr[rr], len = rowptr[rr + 1] - k0;
        bool ok = len <= 8;
        for (int jj = 0; jj < len && ok; ++jj) {
            int d2, p2;
            tl_row_to_dp(g0 + colidx[k0 + jj], n, total, d2, p2);
            const int dd = d2 - d, dp = p2 - pos;
            if (dd < -1 || dd > 1 || dp < -1 || dp > 1) { ok = false; break; }
            pk |= (unsigned long long)(unsigned char)(signed char)(dd * CT_XS + dp) << (8 * jj);
        }
        if (!ok) { pk = ~0ull; return; }          // not a (d +- 1, pos +- 1) stencil: caller flags the tile
        const int ou = ct_ou(d, n), od = ct_od(d, n);
        const unsigned long long nominal =
            (unsigned long long)(unsigned char)(signed char)(-CT_XS + ou) | ((unsigned long long)(unsigned char)(signed char)(-CT_XS + ou + 1) << 8) |
            ((unsigned long long)(unsigned char)(signed char)(-1) << 16) | (0ull << 24) | (1ull << 32) |
            ((unsigned long long)(unsigned char)(signed char)(CT_XS + od) << 40) | ((unsigned long long)(unsigned char)(signed char)(CT_XS + od + 1) << 48);
        layout_ok = (len == 7) && pk == nominal;
    };
    // value template: the first row with the nominal layout
    int vcode = -1;
    for (int q0 = 0; q0 < CT_NC * CT_NQ && vcode < 0; q0 += 32) {
        int r; unsigned long long pk; bool lok;
        inspect(q0 + lane, r, pk, lok);
        const unsigned b = __ballot_sync(0xffffffffu, lok);
        if (b) { const int src = __ffs(b) - 1; vcode = __shfl_sync(0xffffffffu, lok ? (int)code[r] : 0, src); }
    }
    int count = 0;
    bool bad = false;
    const int base = pass ? exc_off[wdx] : 0;
    for (int q0 = 0; q0 < CT_NC * CT_NQ; q0 += 32) {
        const int q = q0 + lane;
        int r; unsigned long long pk; bool lok;
        inspect(q, r, pk, lok);
        bool exc_row = false;
        int cd = 0;
        if (r >= 0) {
            if (pk == ~0ull) bad = true;
            cd = code[r];
            bool same = lok && vcode >= 0;
            if (same)
                for (int jj = 0; jj < FCT_TPL_W; ++jj)
                    same = same && __double_as_longlong(tval[FCT_TPL_W * cd + jj]) == __double_as_longlong(tval[FCT_TPL_W * vcode + jj]);
            exc_row = !same;
        }
        const unsigned b = __ballot_sync(0xffffffffu, exc_row);
        if (pass && exc_row) {
            const int slot = count + __popc(b & ((1u << lane) - 1));
            if (slot < CT_MAXEXC) {
                const int dl = q / CT_NQ + 1, pl = q % CT_NQ + 1;
                exc[base + slot] = make_int4(dl * CT_XS + pl, cd, (int)(unsigned)(pk & 0xffffffffull), (int)(unsigned)(pk >> 32));
            }
        }
        count += __popc(b);
    }
    bad = __any_sync(0xffffffffu, bad);
    if (lane == 0 && !pass) {
        t.z = (bad || count > CT_MAXEXC) ? -1 : count;
        t.w = vcode >= 0 ? vcode : 0;
        tiles[wdx] = t;
    }
}

struct fct_cheb_tiles {
    int4* list[TL_KMAX + 1] = {nullptr};
    int* exc_off[TL_KMAX + 1] = {nullptr};
    int4* exc[TL_KMAX + 1] = {nullptr};
    int count[TL_KMAX + 1] = {0};
    bool ok = false;
};

static void cheb_tile_list_host(int n, long long ga, long long gb, int K, std::vector<int4>& v) {
    const int total = (n + 1) * (n + 1);
    const int Td = CT_ND - 2 * K, Tp = CT_NP - 2 * K;
    v.clear();
    if (gb <= ga) return;
    int dA = 0, dB = 2 * n;
    while (dA < 2 * n && tl_diag_start(dA + 1, n, total) <= ga) ++dA;
    while (dB > 0 && tl_diag_start(dB, n, total) >= gb) --dB;
    for (int d0 = dA; d0 <= dB; d0 += Td) {
        int maxlen = 0;
        for (int d = d0; d < d0 + Td && d <= 2 * n; ++d) maxlen = std::max(maxlen, tl_diag_len(d, n));
        for (int p0 = 0; p0 < maxlen; p0 += Tp) {
            bool any = false;
            for (int d = d0; d < d0 + Td && d <= dB && !any; ++d) {
                const long long s = tl_diag_start(d, n, total);
                const long long lo = std::max(s + p0, ga), hi = std::min(s + std::min(p0 + Tp, tl_diag_len(d, n)), gb);
                any = hi > lo;
            }
            if (any) v.push_back(make_int4(d0, p0, 0, 0));
        }
    }
}

void fct_cheb_tiles_free(fct_ctx* ctx) {
    fct_cheb_tiles* c = ctx->cheb_tiles;
    if (!c) return;
    for (int k = 0; k <= TL_KMAX; ++k) { cudaFree(c->list[k]); cudaFree(c->exc_off[k]); cudaFree(c->exc[k]); }
    delete c;
    ctx->cheb_tiles = nullptr;
}

// builds and classifies the ChebSI tile lists (K = 2..5); on any problem the context keeps the per-iteration kernel
static int cheb_tiles_prepare(fct_ctx* ctx) {
    fct_cheb_tiles_free(ctx);
    ctx->cheb_tiles_ok = false;
    fct_tiles* t = ctx->tiles;
    if (!t || ctx->max_row > 7) return 0;
    static bool attr = false;
    if (!attr) {
        if (cudaFuncSetAttribute(k_cheb_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ChebSmem::BYTES) != cudaSuccess) {
            cudaGetLastError();
            return 0;
        }
        attr = true;
    }
    fct_cheb_tiles* c = new fct_cheb_tiles();
    ctx->cheb_tiles = c;
    const int n = t->n_cells, total = (n + 1) * (n + 1);
    bool ok = true;
    for (int K = 2; K <= TL_KMAX && ok; ++K) {
        std::vector<int4> v;
        int orb, ore;
        tile_out_range(ctx, K, &orb, &ore);
        cheb_tile_list_host(n, (long long)t->g0 + orb, (long long)t->g0 + ore, K, v);
        c->count[K] = (int)v.size();
        const size_t nt = v.size() ? v.size() : 1;
        ok = ok && cudaMalloc((void**)&c->list[K], sizeof(int4) * nt) == cudaSuccess;
        ok = ok && cudaMalloc((void**)&c->exc_off[K], sizeof(int) * nt) == cudaSuccess;
        if (!ok || v.empty()) continue;
        ok = ok && cudaMemcpy(c->list[K], v.data(), sizeof(int4) * v.size(), cudaMemcpyHostToDevice) == cudaSuccess;
        const int grid = ((int)v.size() * 32 + 255) / 256;
        k_cheb_classify<<<grid, 256, 0, ctx->stream>>>(0, c->list[K], (int)v.size(), K, n, total, t->g0, ctx->n, ctx->rowptr,
                                                       ctx->colidx, ctx->tpl_code, ctx->tpl_val, nullptr, nullptr);
        ctx->launches++;
        ok = ok && cudaMemcpyAsync(v.data(), c->list[K], sizeof(int4) * v.size(), cudaMemcpyDeviceToHost, ctx->stream) == cudaSuccess;
        ok = ok && cudaStreamSynchronize(ctx->stream) == cudaSuccess;
        if (!ok) break;
        std::vector<int> off(v.size());
        long long tot = 0;
        for (size_t i = 0; i < v.size(); ++i) {
            if (v[i].z < 0) { ok = false; break; }
            off[i] = (int)tot;
            tot += v[i].z;
        }
        if (!ok) break;
        ok = ok && cudaMalloc((void**)&c->exc[K], sizeof(int4) * (size_t)(tot ? tot : 1)) == cudaSuccess;
        ok = ok && cudaMemcpy(c->exc_off[K], off.data(), sizeof(int) * off.size(), cudaMemcpyHostToDevice) == cudaSuccess;
        if (!ok) break;
        k_cheb_classify<<<grid, 256, 0, ctx->stream>>>(1, c->list[K], (int)v.size(), K, n, total, t->g0, ctx->n, ctx->rowptr,
                                                       ctx->colidx, ctx->tpl_code, ctx->tpl_val, c->exc_off[K], c->exc[K]);
        ctx->launches++;
        ok = ok && cudaStreamSynchronize(ctx->stream) == cudaSuccess;
    }
    cudaGetLastError();
    ctx->cheb_tiles_ok = ok;
    return 0;
}

// K (2..5) Chebyshev iterations in one launch with the weights om[0..K): (ymid, yold) -> (ymid_out, yold_out)
int fct_tile_cheb(fct_ctx* ctx, int K, const double* g, const double* ymid, const double* yold, double* ymid_out,
                  double* yold_out, const double* om, double dscale) {
    FCT_CHECK(ctx->tiles_ok && ctx->cheb_tiles_ok && K >= 2 && K <= TL_KMAX, "fct_tile_cheb: not available");
    fct_tiles* t = ctx->tiles;
    fct_cheb_tiles* c = ctx->cheb_tiles;
    ChebTileArgs a;
    a.n_cells = t->n_cells; a.total = (t->n_cells + 1) * (t->n_cells + 1); a.g0 = t->g0; a.nloc = ctx->n;
    a.own_rb = ctx->row_begin; a.own_re = ctx->row_end; a.ntiles = c->count[K]; a.K = K;
    tile_out_range(ctx, K, &a.out_rb, &a.out_re);
    a.tiles = c->list[K]; a.exc_off = c->exc_off[K]; a.exc = c->exc[K];
    a.tval = ctx->tpl_val; a.tdiag = ctx->tpl_diag;
    a.g = g; a.ymid = ymid; a.yold = yold; a.ymid_out = ymid_out; a.yold_out = yold_out; a.dscale = dscale;
    for (int i = 0; i < TL_KMAX; ++i) a.om[i] = i < K ? om[i] : 0.0;
    int grid = a.ntiles < t->sms ? a.ntiles : t->sms;
    if (ctx->tile_grid_cap > 0 && grid > ctx->tile_grid_cap) grid = ctx->tile_grid_cap;
    if (grid <= 0) return 0;
    k_cheb_tile<<<grid, CT_NT, ChebSmem::BYTES, ctx->stream>>>(a);
    ctx->launches++;
    return 0;
}
