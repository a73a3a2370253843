// CSR passes of the FCT step (fp64, HBM-bound): SpMV, Chebyshev iteration, low-order operator build,
// Jacobi sweeps, Zalesak limiter.  See fct_common.cuh for the row-block staging skeleton.
//
// Reference arithmetic replaced here (KarolinaBenkova/FEM-FCT-PDECO):
//   helpers.py:143-185  ChebSI                  -> k_cheb_first / k_cheb_iter
//   helpers.py:206-242  artificial_diffusion_mat -> k_low_build (fused with the next item) / k_art_diff
//   helpers.py:1775-1782 Mat_u_Low, rhs_u_Low, spsolve -> k_low_build + k_jacobi_sweep (+ k_jacobi_decide)
//   helpers.py:1814     rhs_du_dt = -A u_Low + rhs -> k_spmv
//   helpers.py:1818-1851 fluxes, P+-, Q+-, R+-   -> k_flux_limits
//   helpers.py:1860-1870 limited sum + update    -> k_flux_apply
#include "fct_common.cuh"
#include "fct_pipe.cuh"
#include "../../include/fctpdeco.h"

#include <stdio.h>
#include <stdlib.h>
#include <vector>

// dynamic shared memory layout helpers ---------------------------------------------------------------
extern __shared__ __align__(16) unsigned char fct_smem[];

#define FCT_NST 3    // ring stages of the TMA pipeline, kernels staging one fp64 array
#define FCT_NST2 2   // ... kernels staging two fp64 arrays (leaves L1 room for the four gathered vectors)

// Row dot product out of the staged CSR range: sum_k val[k] * x[col[k]] in column order.  Rows of <= 8 entries (every
// P1 row of the structured mesh) take the unrolled path: all eight gathers are issued before the first FMA, so the
// L1/L2 latency is paid once per row instead of once per entry.  DIAG: also return the diagonal value and leave it
// out of the sum (Jacobi).
template <bool DIAG>
__device__ __forceinline__ double row_dot(const double* __restrict__ sA, const int32_t* __restrict__ sC, int ks, int ke,
                                          const double* __restrict__ x, int r, double& diag) {
    const int len = ke - ks;
    double acc = 0.0;
    if (len <= 8) {
        double v[8], xv[8];
        int c[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const bool p = j < len;
            c[j] = p ? sC[ks + j] : r;
            v[j] = p ? sA[ks + j] : 0.0;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) xv[j] = x[c[j]];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (DIAG) {
                const bool isd = (c[j] == r) && (j < len);
                if (isd) diag = v[j];
                acc += isd ? 0.0 : v[j] * xv[j];
            } else {
                acc += v[j] * xv[j];
            }
        }
    } else {
        for (int k = ks; k < ke; ++k) {
            const int c = sC[k];
            const double v = sA[k];
            if (DIAG && c == r) diag = v;
            else acc += v * x[c];
        }
    }
    return acc;
}

// y = alpha * A x + beta * z
__global__ void __launch_bounds__(FCT_RB)
k_spmv(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx, const double* __restrict__ A,
       const double* __restrict__ x, double alpha, double beta, const double* __restrict__ z,
       double* __restrict__ y, int row_begin, int row_end, int64_t nnz, int cap) {
    __shared__ __align__(8) uint64_t bars[FCT_NST];
    RowPipe<1, 1, FCT_NST> pipe{fct_smem, bars, cap, row_begin, row_end, nnz, rowptr, {A}, {colidx}};
    const int nmine = pipe.my_blocks();
    pipe.init();
    pipe.prologue(nmine);
    for (int i = 0; i < nmine; ++i) {
        pipe.prefetch(i, nmine);
        const RowBlock b = pipe.block(i);
        const bool act = (int)threadIdx.x < b.nr;
        const int r = b.r0 + threadIdx.x;
        int ks = 0, ke = 0;
        double zr = 0.0;
        if (act) {
            ks = rowptr[r] - b.ka; ke = rowptr[r + 1] - b.ka;
            if (beta != 0.0) zr = z[r];
        }
        pipe.wait(i, b);
        if (act) {
            const double* sA = pipe.f64(i % FCT_NST, 0);
            const int32_t* sC = pipe.s32(i % FCT_NST, 0);
            double dummy;
            const double acc = row_dot<false>(sA, sC, ks, ke, x, r, dummy);
            double out = alpha * acc;
            if (beta != 0.0) out += beta * zr;
            y[r] = out;
        }
        __syncthreads();
    }
}

// y = alpha * A x + beta * z with the column pattern of A taken from the row templates (every matrix of a context lives
// on the pattern of M): 8 B per entry instead of 12; same products in the same order as k_spmv.
__global__ void __launch_bounds__(FCT_RB)
k_spmv_tc(const int32_t* __restrict__ rowptr, const uint16_t* __restrict__ code, const int32_t* __restrict__ toff,
          const double* __restrict__ A, const double* __restrict__ x, double alpha, double beta, const double* __restrict__ z,
          double* __restrict__ y, int row_begin, int row_end, int64_t nnz, int cap) {
    __shared__ __align__(8) uint64_t bars[FCT_NST];
    RowPipe<1, 0, FCT_NST> pipe{fct_smem, bars, cap, row_begin, row_end, nnz, rowptr, {A}, {nullptr}};
    const int nmine = pipe.my_blocks();
    pipe.init();
    pipe.prologue(nmine);
    struct RowIn { int k0, len; double z; int4 o0, o1; };
    auto row_of = [&](int i) {
        const int blk = (int)blockIdx.x + i * (int)gridDim.x;
        const int r = row_begin + blk * FCT_RB + (int)threadIdx.x;
        return (i < nmine && r < row_end) ? r : -1;
    };
    auto load_code = [&](int i) {
        const int r = row_of(i);
        return r >= 0 ? (int)code[r] : 0;
    };
    auto load_row = [&](int i, int t) {
        RowIn in{0, 0, 0.0, make_int4(0, 0, 0, 0), make_int4(0, 0, 0, 0)};
        const int r = row_of(i);
        if (r >= 0) {
            in.k0 = rowptr[r]; in.len = rowptr[r + 1] - in.k0;
            if (beta != 0.0) in.z = z[r];
            in.o0 = __ldg(reinterpret_cast<const int4*>(toff + FCT_TPL_W * t));
            in.o1 = __ldg(reinterpret_cast<const int4*>(toff + FCT_TPL_W * t) + 1);
        }
        return in;
    };
    RowIn cur = load_row(0, load_code(0));
    int tnext = load_code(1);
    for (int i = 0; i < nmine; ++i) {
        pipe.prefetch(i, nmine);
        const RowIn nxt = load_row(i + 1, tnext);     // the code of a row is fetched two blocks ahead, its offsets one
        tnext = load_code(i + 2);
        const RowBlock b = pipe.block(i);
        pipe.wait(i, b);
        if ((int)threadIdx.x < b.nr) {
            const int r = b.r0 + threadIdx.x;
            const double* sA = pipe.f64(i % FCT_NST, 0) + (cur.k0 - b.ka);
            const int off[8] = {cur.o0.x, cur.o0.y, cur.o0.z, cur.o0.w, cur.o1.x, cur.o1.y, cur.o1.z, cur.o1.w};
            double xv[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) xv[j] = x[r + off[j]];
            double acc = 0.0;
#pragma unroll
            for (int j = 0; j < 8; ++j) acc += ((j < cur.len) ? sA[j] : 0.0) * xv[j];
            double out = alpha * acc;
            if (beta != 0.0) out += beta * cur.z;
            y[r] = out;
        }
        cur = nxt;
        __syncthreads();
    }
}

// ChebSI iteration k == 1: ymid = yold = 0  =>  y1 = omega1 * (b / Md')  with omega1 = 1
__global__ void __launch_bounds__(FCT_RB)
k_cheb_first(const double* __restrict__ g, const double* __restrict__ Md, double dscale, double omega,
             double* __restrict__ ynew, int row_begin, int row_end) {
    const int r = row_begin + blockIdx.x * FCT_RB + threadIdx.x;
    if (r < row_end) {
        const double z = g[r] / (dscale * Md[r]);
        ynew[r] = omega * z;
    }
}

// ChebSI iteration k >= 2 (helpers.py:176-184):
//   r = b - M ymid; z = r / Md'; ynew = omega (z + ymid - yold) + yold
template <int NST>
__global__ void __launch_bounds__(FCT_RB)
k_cheb_iter(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx, const double* __restrict__ Mv,
            const double* __restrict__ Md, const double* __restrict__ g, const double* __restrict__ ymid,
            const double* __restrict__ yold, double* __restrict__ ynew, double omega, double dscale,
            int has_old, int row_begin, int row_end, int64_t nnz, int cap) {
    __shared__ __align__(8) uint64_t bars[NST];
    RowPipe<1, 1, NST> pipe{fct_smem, bars, cap, row_begin, row_end, nnz, rowptr, {Mv}, {colidx}};
    const int nmine = pipe.my_blocks();
    pipe.init();
    pipe.prologue(nmine);
    // the row's own inputs are fetched one block ahead (registers), so their DRAM latency overlaps the previous block
    struct RowIn { int k0, k1; double g, md, ym, yo; };
    auto load_row = [&](int i) {
        RowIn in{0, 0, 0.0, 1.0, 0.0, 0.0};
        if (i < nmine) {
            const int blk = (int)blockIdx.x + i * (int)gridDim.x;
            const int r = row_begin + blk * FCT_RB + (int)threadIdx.x;
            if (r < row_end) {
                in.k0 = rowptr[r]; in.k1 = rowptr[r + 1];
                in.g = g[r]; in.md = Md[r]; in.ym = ymid[r];
                if (has_old) in.yo = yold[r];
            }
        }
        return in;
    };
    RowIn cur = load_row(0);
    for (int i = 0; i < nmine; ++i) {
        pipe.prefetch(i, nmine);
        const RowIn nxt = load_row(i + 1);
        const RowBlock b = pipe.block(i);
        pipe.wait(i, b);
        if ((int)threadIdx.x < b.nr) {
            const int r = b.r0 + threadIdx.x;
            double dummy;
            const double acc = row_dot<false>(pipe.f64(i % NST, 0), pipe.s32(i % NST, 0), cur.k0 - b.ka, cur.k1 - b.ka,
                                              ymid, r, dummy);
            const double z = (cur.g - acc) / (dscale * cur.md);
            ynew[r] = omega * (z + cur.ym - cur.yo) + cur.yo;
        }
        cur = nxt;
        __syncthreads();
    }
}

// sortable-key transform for atomicMin on signed doubles
__device__ __forceinline__ unsigned long long f64_sort_key(double v) {
    unsigned long long u = (unsigned long long)__double_as_longlong(v);
    return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}

// Build the low-order operator and right-hand side in one pass:
//   d_ij = max(0, a_ij, a_ji) (i != j)        [artificial_diffusion_mat(-A), helpers.py:1769]
//   L    = M_L + dt (A - D) (+ dt S)            [helpers.py:1775-1778]
//   b    = M_L u_n + dt rhs                     [helpers.py:1780]
// `sign` folds the legacy FCT_alg convention (A -> -A, old_helpers.py:135-145) into the same kernel.
// Outputs: Lv (off-diagonals of L, diagonal slot 0), dinv = 1/l_ii, Dv (off-diagonals; diagonal slot holds d_ii), b;
// min row sum of L.  scale_rows: Lv and b are stored divided by l_ii (the Jacobi fixed point is unchanged).
// A (and S), colidx, tpos arrive through a 2-stage TMA ring; L and D leave through two staging buffers.
#define FCT_NST_LOW 2
template <int HAS_S>
__global__ void __launch_bounds__(FCT_RB)
k_low_build(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx, const int32_t* __restrict__ tpos,
            const double* __restrict__ A, double sign, const double* __restrict__ S, const double* __restrict__ ML,
            const double* __restrict__ un, const double* __restrict__ rhs, double dt,
            double* __restrict__ Lv, double* __restrict__ Dv, double* __restrict__ bvec, double* __restrict__ dinv,
            unsigned long long* __restrict__ min_rowsum_key, int scale_rows, int row_begin, int row_end, int64_t nnz,
            int cap) {
    __shared__ __align__(8) uint64_t bars[FCT_NST_LOW];
    __shared__ double sred[FCT_RB / 32];
    // the column indices are not staged: the diagonal is the entry that is its own transpose (tpos[k] == k)
    RowPipe<1 + HAS_S, 1, FCT_NST_LOW> pipe{fct_smem, bars, cap, row_begin, row_end, nnz, rowptr, {A}, {tpos}};
    if (HAS_S) pipe.gf[HAS_S] = S;
    double* sL = reinterpret_cast<double*>(fct_smem + FCT_NST_LOW * pipe.stage_bytes());
    double* sD = sL + cap;
    const int nmine = pipe.my_blocks();
    pipe.init();
    pipe.prologue(nmine);
    double rowsum = 1e300;
    struct RowIn { int k0, k1; double ml, un, rhs; };
    auto load_row = [&](int i) {
        RowIn in{0, 0, 0.0, 0.0, 0.0};
        if (i < nmine) {
            const int blk = (int)blockIdx.x + i * (int)gridDim.x;
            const int r = row_begin + blk * FCT_RB + (int)threadIdx.x;
            if (r < row_end) {
                in.k0 = rowptr[r]; in.k1 = rowptr[r + 1];
                in.ml = ML[r]; in.un = un[r];
                if (rhs) in.rhs = rhs[r];
            }
        }
        return in;
    };
    RowIn cur = load_row(0);
    for (int i = 0; i < nmine; ++i) {
        pipe.prefetch(i, nmine);
        const RowIn nxt = load_row(i + 1);
        const RowBlock b = pipe.block(i);
        const bool act = (int)threadIdx.x < b.nr;
        const int r = b.r0 + threadIdx.x;
        const int ks = cur.k0 - b.ka, ke = cur.k1 - b.ka;
        const double ml = cur.ml, unr = cur.un, rr = cur.rhs;
        pipe.wait(i, b);
        if (act) {
            const int st = i % FCT_NST_LOW;
            const double* sA = pipe.f64(st, 0);
            const double* sS = pipe.f64(st, HAS_S ? 1 : 0);
            const int32_t* sT = pipe.s32(st, 0);
            const int kg = cur.k0 - ks;              // global CSR index of staged slot 0
            double dsum = 0.0, lsum = 0.0;
            int kd = ks;
            const int len = ke - ks;
            if (len <= 8) {
                // all transposed-entry gathers a_ji = A[tpos] are issued before the first use
                double av[8], atv[8];
                int tv[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const bool p = j < len;
                    tv[j] = p ? sT[ks + j] : -1;
                    av[j] = p ? sA[ks + j] : 0.0;
                    atv[j] = p ? A[tv[j]] : 0.0;
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    if (j < len) {
                        if (tv[j] == kg + ks + j) { kd = ks + j; continue; }
                        const double a = sign * av[j];
                        const double at = sign * atv[j];
                        const double d = fmax(0.0, fmax(a, at));
                        dsum += d;
                        double l = dt * (a - d);
                        if (HAS_S) l += dt * sS[ks + j];
                        lsum += l;
                        sL[ks + j] = l;
                        sD[ks + j] = d;
                    }
                }
            } else {
                for (int k = ks; k < ke; ++k) {
                    if (sT[k] == kg + k) { kd = k; continue; }
                    const double a = sign * sA[k];
                    const double at = sign * A[sT[k]];
                    const double d = fmax(0.0, fmax(a, at));
                    dsum += d;
                    double l = dt * (a - d);
                    if (HAS_S) l += dt * sS[k];
                    lsum += l;
                    sL[k] = l;
                    sD[k] = d;
                }
            }
            // diagonal: d_ii = -sum_j d_ij
            const double a = sign * sA[kd];
            const double dii = -dsum;
            double l = ml + dt * (a - dii);
            if (HAS_S) l += dt * sS[kd];
            lsum += l;
            sL[kd] = 0.0;          // the diagonal travels as 1/l_ii in `dinv`: the Jacobi row loop is branch-free
            const double di = 1.0 / l;
            if (dinv) dinv[r] = di;
            sD[kd] = dii;
            rowsum = fmin(rowsum, lsum);
            double br = ml * unr + (rhs ? dt * rr : 0.0);
            if (scale_rows) {      // Jacobi on the row-scaled system (k_jacobi_sweep_tpl<.., true>): x_new = b' - sum l'_ij x_j
                for (int k = ks; k < ke; ++k) sL[k] *= di;
                br *= di;
            }
            bvec[r] = br;
        }
        cur = nxt;
        __syncthreads();
        unstage_f64(Lv, sL, b);
        unstage_f64(Dv, sD, b);
        __syncthreads();
    }
    const double m = block_min(rowsum, sred);
    if (threadIdx.x == 0) atomicMin(min_rowsum_key, f64_sort_key(m));
}

// artificial_diffusion_mat as a stand-alone operation (helpers.py:206-242): D from `mat`
__global__ void __launch_bounds__(FCT_RB)
k_art_diff(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx, const int32_t* __restrict__ tpos,
           const double* __restrict__ mat, double* __restrict__ Dv, int row_begin, int row_end, int64_t nnz, int cap) {
    double* sA = reinterpret_cast<double*>(fct_smem);
    int32_t* sC = reinterpret_cast<int32_t*>(sA + cap);
    int32_t* sT = sC + cap;
    FCT_FOR_ROW_BLOCKS(blk, row_begin, row_end) {
    const RowBlock b = row_block(rowptr, row_begin, row_end, blk);
    stage_f64(sA, mat, b, nnz);
    stage_s32(sC, colidx, b, nnz);
    stage_s32(sT, tpos, b, nnz);
    __syncthreads();
    if ((int)threadIdx.x < b.nr) {
        const int r = b.r0 + threadIdx.x;
        const int ks = rowptr[r] - b.ka, ke = rowptr[r + 1] - b.ka;
        double dsum = 0.0;
        int kd = -1;
        for (int k = ks; k < ke; ++k) {
            if (sC[k] == r) { kd = k; continue; }
            const double d = fmax(0.0, fmax(-sA[k], -mat[sT[k]]));
            dsum += d;
            sA[k] = d;
        }
        sA[kd] = -dsum;
    }
    __syncthreads();
    unstage_f64(Dv, sA, b);
    __syncthreads();
    }
}

// One Jacobi sweep x_new = (b - sum_{j != i} l_ij x_j) / l_ii.  Skipped once jstate[3] (converged) is set.
// When `check` is set the sweep also accumulates ||x_new - x||_inf and ||x_new||_inf into jstate[0..1].
// `dinv` != nullptr: Lv holds the off-diagonals only (diagonal slot 0) and dinv = 1/diag (the FCT low-order system,
// written that way by k_low_build); dinv == nullptr: general matrix, the diagonal is picked out of the row.
template <int NST, bool SEP>
__device__ __forceinline__ void
jacobi_sweep_body(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx, const double* __restrict__ Lv,
               const double* __restrict__ bvec, const double* __restrict__ dinv, const double* __restrict__ x,
               double* __restrict__ xnew, unsigned long long* __restrict__ jstate, int check, int own_rb, int own_re,
               int row_begin, int row_end, int64_t nnz, int cap) {
    if (*reinterpret_cast<volatile unsigned long long*>(jstate + 3)) return;
    __shared__ __align__(8) uint64_t bars[NST];
    __shared__ double sred[FCT_RB / 32];
    RowPipe<1, 1, NST> pipe{fct_smem, bars, cap, row_begin, row_end, nnz, rowptr, {Lv}, {colidx}};
    const int nmine = pipe.my_blocks();
    pipe.init();
    pipe.prologue(nmine);
    struct RowIn { int k0, k1; double b, x, di; };
    auto load_row = [&](int i) {
        RowIn in{0, 0, 0.0, 0.0, 1.0};
        if (i < nmine) {
            const int blk = (int)blockIdx.x + i * (int)gridDim.x;
            const int r = row_begin + blk * FCT_RB + (int)threadIdx.x;
            if (r < row_end) {
                in.k0 = rowptr[r]; in.k1 = rowptr[r + 1];
                in.b = bvec[r];
                if (SEP) in.di = dinv[r];
                if (check) in.x = x[r];
            }
        }
        return in;
    };
    RowIn cur = load_row(0);
    double delta = 0.0, xa = 0.0;
    for (int i = 0; i < nmine; ++i) {
        pipe.prefetch(i, nmine);
        const RowIn nxt = load_row(i + 1);
        const RowBlock b = pipe.block(i);
        pipe.wait(i, b);
        if ((int)threadIdx.x < b.nr) {
            const int r = b.r0 + threadIdx.x;
            double xn;
            if (SEP) {
                double dummy;
                const double acc = row_dot<false>(pipe.f64(i % NST, 0), pipe.s32(i % NST, 0), cur.k0 - b.ka,
                                                  cur.k1 - b.ka, x, r, dummy);
                xn = (cur.b - acc) * cur.di;
            } else {
                double diag = 1.0;
                const double acc = row_dot<true>(pipe.f64(i % NST, 0), pipe.s32(i % NST, 0), cur.k0 - b.ka,
                                                 cur.k1 - b.ka, x, r, diag);
                xn = (cur.b - acc) / diag;
            }
            xnew[r] = xn;
            if (check && r >= own_rb && r < own_re) {      // the stopping test looks at the owned rows only
                delta = fmax(delta, fabs(xn - cur.x));
                xa = fmax(xa, fabs(xn));
            }
        }
        cur = nxt;
        __syncthreads();
    }
    if (check) {
        const double dm = block_max(delta, sred);
        const double xm = block_max(xa, sred);
        if (threadIdx.x == 0) {
            atomicMax(jstate + 0, (unsigned long long)__double_as_longlong(dm));
            atomicMax(jstate + 1, (unsigned long long)__double_as_longlong(xm));
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(jstate + 4, 1ull);
}

template <int NST>
__global__ void __launch_bounds__(FCT_RB, 5)
k_jacobi_sweep(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx, const double* __restrict__ Lv,
               const double* __restrict__ bvec, const double* __restrict__ dinv, const double* __restrict__ x,
               double* __restrict__ xnew, unsigned long long* __restrict__ jstate, int check, int own_rb, int own_re,
               int row_begin, int row_end, int64_t nnz, int cap) {
    jacobi_sweep_body<NST, true>(rowptr, colidx, Lv, bvec, dinv, x, xnew, jstate, check, own_rb, own_re, row_begin, row_end,
                                 nnz, cap);
}
template <int NST>
__global__ void __launch_bounds__(FCT_RB)
k_jacobi_sweep_gen(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx, const double* __restrict__ Lv,
                   const double* __restrict__ bvec, const double* __restrict__ dinv, const double* __restrict__ x,
                   double* __restrict__ xnew, unsigned long long* __restrict__ jstate, int check, int own_rb, int own_re,
                   int row_begin, int row_end, int64_t nnz, int cap) {
    jacobi_sweep_body<NST, false>(rowptr, colidx, Lv, bvec, dinv, x, xnew, jstate, check, own_rb, own_re, row_begin,
                                  row_end, nnz, cap);
}

// Jacobi sweep of the FCT low-order system with the column pattern taken from the row templates of M (every matrix of a
// context lives on the same pattern): per row a 16-bit code replaces the 4 B/entry column indices, the TMA ring carries
// the matrix values only.  SCALED = false: x_new = (b - sum l_ij x_j) / l_ii with dinv = 1/l_ii -- the same operations in
// the same order as k_jacobi_sweep, bit-identical.  SCALED = true: k_low_build has divided the row and b by l_ii, so
// x_new = b' - sum l'_ij x_j and dinv is not read.  The code of a row is fetched two blocks ahead and its offsets one
// block ahead, so the code -> offsets -> gather chain never waits on DRAM.
template <int NST, bool SCALED, int MINB = 5>
__global__ void __launch_bounds__(FCT_RB, MINB)
k_jacobi_sweep_tpl(const int32_t* __restrict__ rowptr, const uint16_t* __restrict__ code, const int32_t* __restrict__ toff,
                   const double* __restrict__ Lv, const double* __restrict__ bvec, const double* __restrict__ dinv,
                   const double* __restrict__ x, double* __restrict__ xnew, unsigned long long* __restrict__ jstate, int check,
                   int own_rb, int own_re, int row_begin, int row_end, int64_t nnz, int cap) {
    if (*reinterpret_cast<volatile unsigned long long*>(jstate + 3)) return;
    __shared__ __align__(8) uint64_t bars[NST];
    __shared__ double sred[FCT_RB / 32];
    RowPipe<1, 0, NST> pipe{fct_smem, bars, cap, row_begin, row_end, nnz, rowptr, {Lv}, {nullptr}};
    const int nmine = pipe.my_blocks();
    pipe.init();
    pipe.prologue(nmine);
    struct RowIn { int k0, len; double b, x, di; int4 o0, o1; };
    auto row_of = [&](int i) {
        const int blk = (int)blockIdx.x + i * (int)gridDim.x;
        const int r = row_begin + blk * FCT_RB + (int)threadIdx.x;
        return (i < nmine && r < row_end) ? r : -1;
    };
    auto load_code = [&](int i) {
        const int r = row_of(i);
        return r >= 0 ? (int)code[r] : 0;
    };
    auto load_row = [&](int i, int t) {
        RowIn in{0, 0, 0.0, 0.0, 1.0, make_int4(0, 0, 0, 0), make_int4(0, 0, 0, 0)};
        const int r = row_of(i);
        if (r >= 0) {
            in.k0 = rowptr[r]; in.len = rowptr[r + 1] - in.k0;
            in.b = bvec[r];
            if (!SCALED) in.di = dinv[r];
            if (check) in.x = x[r];
            in.o0 = __ldg(reinterpret_cast<const int4*>(toff + FCT_TPL_W * t));
            in.o1 = __ldg(reinterpret_cast<const int4*>(toff + FCT_TPL_W * t) + 1);
        }
        return in;
    };
    RowIn cur = load_row(0, load_code(0));
    int tnext = load_code(1);
    double delta = 0.0, xa = 0.0;
    for (int i = 0; i < nmine; ++i) {
        pipe.prefetch(i, nmine);
        const RowIn nxt = load_row(i + 1, tnext);
        tnext = load_code(i + 2);
        const RowBlock b = pipe.block(i);
        pipe.wait(i, b);
        if ((int)threadIdx.x < b.nr) {
            const int r = b.r0 + threadIdx.x;
            const double* sL = pipe.f64(i % NST, 0) + (cur.k0 - b.ka);
            const int off[8] = {cur.o0.x, cur.o0.y, cur.o0.z, cur.o0.w, cur.o1.x, cur.o1.y, cur.o1.z, cur.o1.w};
            double xv[8], v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) xv[j] = x[r + off[j]];          // padded slots: offset 0
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = (j < cur.len) ? sL[j] : 0.0;
            double acc = 0.0;
#pragma unroll
            for (int j = 0; j < 8; ++j) acc += v[j] * xv[j];            // diagonal slot holds 0
            const double xn = SCALED ? (cur.b - acc) : (cur.b - acc) * cur.di;
            xnew[r] = xn;
            if (check && r >= own_rb && r < own_re) {
                delta = fmax(delta, fabs(xn - cur.x));
                xa = fmax(xa, fabs(xn));
            }
        }
        cur = nxt;
        __syncthreads();
    }
    if (check) {
        const double dm = block_max(delta, sred);
        const double xm = block_max(xa, sred);
        if (threadIdx.x == 0) {
            atomicMax(jstate + 0, (unsigned long long)__double_as_longlong(dm));
            atomicMax(jstate + 1, (unsigned long long)__double_as_longlong(xm));
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(jstate + 4, 1ull);
}

// After a checked sweep (and, multi-GPU, after the max-allreduce of jstate[0..1]): decide convergence.
// jstate[10] = "check_from": sweeps before which the stopping test is not attempted.  The sweep count of the
// low-order solve is nearly constant from one time step to the next, so the early tests (each of which is an
// all-to-all between the ranks in the multi-GPU case) are skipped; jstate[11] counts the failed tests of this solve.
__device__ __forceinline__ void jacobi_note_convergence(unsigned long long* jstate) {
    const unsigned long long s = jstate[4];
    const unsigned long long back = jstate[11] ? 2ull : 4ull;
    jstate[10] = s > back ? s - back : 0ull;
}
__global__ void k_jacobi_decide(unsigned long long* __restrict__ jstate, double rtol, unsigned long long max_sweeps) {
    if (jstate[3]) return;
    // same schedule as the multi-GPU test; a learnt check_from beyond max_sweeps must not suppress every test
    if (jstate[4] < jstate[10] && jstate[4] < max_sweeps) { jstate[0] = 0ull; jstate[1] = 0ull; return; }
    const double delta = __longlong_as_double((long long)jstate[0]);
    const double xm = __longlong_as_double((long long)jstate[1]);
    jstate[5] = jstate[0];
    jstate[6] = jstate[1];
    if (delta <= rtol * xm) { jstate[3] = 1ull; jacobi_note_convergence(jstate); }
    else jstate[11] += 1ull;
    jstate[0] = 0ull;
    jstate[1] = 0ull;
}

// the same decision as the body of a CUDA-graph WHILE node: keeps looping until converged or out of sweeps
__global__ void k_jacobi_decide_cond(unsigned long long* __restrict__ jstate, double rtol, unsigned long long max_sweeps,
                                     cudaGraphConditionalHandle handle) {
    if (jstate[3]) { cudaGraphSetConditional(handle, 0u); return; }     // already converged (fused launch)
    if (jstate[4] < jstate[10] && jstate[4] < max_sweeps) {
        jstate[0] = 0ull; jstate[1] = 0ull;
        cudaGraphSetConditional(handle, 1u);
        return;
    }
    const double delta = __longlong_as_double((long long)jstate[0]);
    const double xm = __longlong_as_double((long long)jstate[1]);
    jstate[5] = jstate[0];
    jstate[6] = jstate[1];
    const bool conv = delta <= rtol * xm;
    if (conv) { jstate[3] = 1ull; jacobi_note_convergence(jstate); }
    else jstate[11] += 1ull;
    jstate[0] = 0ull;
    jstate[1] = 0ull;
    cudaGraphSetConditional(handle, (conv || jstate[4] >= max_sweeps) ? 0u : 1u);
}

// stopping test after a fused K-sweep launch (fct_tile.cu); `which` = 1 when that launch wrote the scratch iterate, so
// that the solve can copy it back (k_copy_if) when the loop ends there
__global__ void k_tile_decide(unsigned long long* __restrict__ jstate, double rtol, unsigned long long max_sweeps,
                              unsigned long long which, int use_handle, cudaGraphConditionalHandle handle,
                              unsigned long long kmax) {
    if (jstate[3]) { if (use_handle) cudaGraphSetConditional(handle, 0u); return; }
    // the same test schedule as the multi-GPU decision (k_p2p_max2_decide): no test before the sweep count learnt from the
    // previous solve -- on one GPU the test is free, but N ranks must stop after the same number of sweeps as one
    if (jstate[4] < jstate[10] && jstate[4] < max_sweeps) {
        jstate[0] = 0ull; jstate[1] = 0ull;
        if (kmax) jstate[13] = tile_next_k(jstate[4], jstate[16], kmax);
        if (use_handle) cudaGraphSetConditional(handle, 1u);
        return;
    }
    const double delta = __longlong_as_double((long long)jstate[0]);
    const double xm = __longlong_as_double((long long)jstate[1]);
    jstate[5] = jstate[0];
    jstate[6] = jstate[1];
    const bool conv = delta <= rtol * xm;
    if (conv) { jstate[3] = 1ull; jstate[12] = which; jacobi_note_convergence(jstate); if (kmax) tile_schedule_converged(jstate, rtol * xm, delta); }
    else { jstate[11] += 1ull; if (kmax) tile_schedule_failed(jstate, delta, kmax); }
    jstate[0] = 0ull;
    jstate[1] = 0ull;
    if (use_handle) cudaGraphSetConditional(handle, (conv || jstate[4] >= max_sweeps) ? 0u : 1u);
}
__global__ void k_copy_if(const unsigned long long* __restrict__ flag, const double* __restrict__ src, double* __restrict__ dst,
                          int n) {
    if (*flag == 0ull) return;
    const int stride = (int)gridDim.x * (int)blockDim.x;
    for (int i = (int)blockIdx.x * (int)blockDim.x + (int)threadIdx.x; i < n; i += stride) dst[i] = src[i];
}

__global__ void k_jacobi_reset(unsigned long long* __restrict__ jstate, unsigned long long tile_kmax) {
    if (tile_kmax) tile_schedule_begin(jstate, tile_kmax);
    jstate[0] = 0ull; jstate[1] = 0ull; jstate[2] = 0ull; jstate[3] = 0ull; jstate[4] = 0ull;
    jstate[5] = 0ull; jstate[6] = 0ull;
    jstate[12] = 0ull;                   // 1: the converged iterate sits in the scratch vector (fused tile sweeps)
    jstate[7] = 0xFFFFFFFFFFFFFFFFull;   // min row-sum key
    jstate[11] = 0ull;                   // failed stopping tests of this solve (jstate[10], check_from, persists)
}

// Zalesak limiter, pass 1 (helpers.py:1818-1851): raw fluxes f_ij = m_ij (ud_i - ud_j) + d_ij (ul_i - ul_j),
// P+- = sums of positive/negative fluxes, Q+- = distance to the local extrema of u_low, R+- nodal factors.
// TPL: the mass-matrix entries come from the row templates (fct_templates.cu) instead of the TMA ring -- one fp64
// array less to stream.
template <bool TPL>
__global__ void __launch_bounds__(FCT_RB, 3)
k_flux_limits(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx, const double* __restrict__ Mv,
              const uint16_t* __restrict__ tcode, const int32_t* __restrict__ toff, const double* __restrict__ tval, const double* __restrict__ Dv, const double* __restrict__ ML, const double* __restrict__ udot,
              const double* __restrict__ ulow, double dt, double* __restrict__ Rpos, double* __restrict__ Rneg,
              int row_begin, int row_end, int64_t nnz, int cap) {
    __shared__ __align__(8) uint64_t bars[FCT_NST2];
    RowPipe<(TPL ? 1 : 2), (TPL ? 0 : 1), FCT_NST2> pipe{fct_smem, bars, cap, row_begin, row_end, nnz, rowptr, {TPL ? Dv : Mv},
                                                             {TPL ? nullptr : colidx}};
    if (!TPL) pipe.gf[TPL ? 0 : 1] = Dv;
    const int nmine = pipe.my_blocks();
    pipe.init();
    pipe.prologue(nmine);
    struct RowIn { int k0, k1, t; double ud, ul, ml; };
    auto load_row = [&](int i) {
        RowIn in{0, 0, 0, 0.0, 0.0, 1.0};
        if (i < nmine) {
            const int blk = (int)blockIdx.x + i * (int)gridDim.x;
            const int r = row_begin + blk * FCT_RB + (int)threadIdx.x;
            if (r < row_end) {
                in.k0 = rowptr[r]; in.k1 = rowptr[r + 1];
                if (TPL) in.t = tcode[r];
                in.ud = udot[r]; in.ul = ulow[r]; in.ml = ML[r];
            }
        }
        return in;
    };
    RowIn cur = load_row(0);
    for (int i = 0; i < nmine; ++i) {
        pipe.prefetch(i, nmine);
        const RowIn nxt = load_row(i + 1);
        const RowBlock b = pipe.block(i);
        pipe.wait(i, b);
        if ((int)threadIdx.x < b.nr) {
            const int r = b.r0 + threadIdx.x;
            const int ks = cur.k0 - b.ka, ke = cur.k1 - b.ka, len = ke - ks;
            const double* sM = pipe.f64(i % FCT_NST2, 0);
            const double* sD = pipe.f64(i % FCT_NST2, TPL ? 0 : 1);
            double tm[8];
            int toffs[8];
            if (TPL) {
                // m_ij and the column offsets of the row come from its template (code fetched one block ahead)
                const double2* tp = reinterpret_cast<const double2*>(tval + 8 * cur.t);
#pragma unroll
                for (int j = 0; j < 4; ++j) { const double2 v = __ldg(tp + j); tm[2 * j] = v.x; tm[2 * j + 1] = v.y; }
                const int4 o0 = __ldg(reinterpret_cast<const int4*>(toff + 8 * cur.t));
                const int4 o1 = __ldg(reinterpret_cast<const int4*>(toff + 8 * cur.t) + 1);
                toffs[0] = o0.x; toffs[1] = o0.y; toffs[2] = o0.z; toffs[3] = o0.w;
                toffs[4] = o1.x; toffs[5] = o1.y; toffs[6] = o1.z; toffs[7] = o1.w;
            }
            const int32_t* sC = TPL ? nullptr : pipe.s32(i % FCT_NST2, 0);
            const double udi = cur.ud, uli = cur.ul;
            double pp = 0.0, pn = 0.0, umax = uli, umin = uli;
            if (len <= 8) {
                int c[8];
                double udj[8], ulj[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) c[j] = (j < len) ? (TPL ? r + toffs[j] : sC[ks + j]) : r;
#pragma unroll
                for (int j = 0; j < 8; ++j) { udj[j] = udot[c[j]]; ulj[j] = ulow[c[j]]; }
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    if (j < len && c[j] != r) {
                        const double f = (TPL ? tm[j] : sM[ks + j]) * (udi - udj[j]) + sD[ks + j] * (uli - ulj[j]);
                        pp += fmax(f, 0.0);
                        pn += fmin(f, 0.0);
                        umax = fmax(umax, ulj[j]);
                        umin = fmin(umin, ulj[j]);
                    }
                }
            } else {
                for (int k = ks; k < ke; ++k) {
                    const int c = sC[k];
                    if (c == r) continue;
                    const double ulj = ulow[c];
                    const double f = sM[k] * (udi - udot[c]) + sD[k] * (uli - ulj);
                    pp += fmax(f, 0.0);
                    pn += fmin(f, 0.0);
                    umax = fmax(umax, ulj);
                    umin = fmin(umin, ulj);
                }
            }
            const double qp = umax - uli, qn = umin - uli;
            Rpos[r] = (pp != 0.0) ? fmin(1.0, cur.ml * qp / (dt * pp)) : 1.0;
            Rneg[r] = (pn != 0.0) ? fmin(1.0, cur.ml * qn / (dt * pn)) : 1.0;
        }
        cur = nxt;
        __syncthreads();
    }
}

// Zalesak limiter, pass 2 (helpers.py:1860-1870): alpha_ij, limited sum, explicit correction.
template <bool TPL>
__global__ void __launch_bounds__(FCT_RB, 3)
k_flux_apply(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx, const double* __restrict__ Mv,
             const uint16_t* __restrict__ tcode, const int32_t* __restrict__ toff, const double* __restrict__ tval, const double* __restrict__ Dv, const double* __restrict__ ML, const double* __restrict__ udot,
             const double* __restrict__ ulow, const double* __restrict__ Rpos, const double* __restrict__ Rneg,
             double dt, double* __restrict__ uout, int row_begin, int row_end, int64_t nnz, int cap) {
    __shared__ __align__(8) uint64_t bars[FCT_NST2];
    // column indices stay staged here: taking them from the templates was measured slower (0.64 vs 0.55 ms)
    RowPipe<(TPL ? 1 : 2), 1, FCT_NST2> pipe{fct_smem, bars, cap, row_begin, row_end, nnz, rowptr, {TPL ? Dv : Mv}, {colidx}};
    if (!TPL) pipe.gf[TPL ? 0 : 1] = Dv;
    const int nmine = pipe.my_blocks();
    pipe.init();
    pipe.prologue(nmine);
    struct RowIn { int k0, k1, t; double ud, ul, ml, rp, rn; };
    auto load_row = [&](int i) {
        RowIn in{0, 0, 0, 0.0, 0.0, 1.0, 1.0, 1.0};
        if (i < nmine) {
            const int blk = (int)blockIdx.x + i * (int)gridDim.x;
            const int r = row_begin + blk * FCT_RB + (int)threadIdx.x;
            if (r < row_end) {
                in.k0 = rowptr[r]; in.k1 = rowptr[r + 1];
                if (TPL) in.t = tcode[r];
                in.ud = udot[r]; in.ul = ulow[r]; in.ml = ML[r]; in.rp = Rpos[r]; in.rn = Rneg[r];
            }
        }
        return in;
    };
    RowIn cur = load_row(0);
    for (int i = 0; i < nmine; ++i) {
        pipe.prefetch(i, nmine);
        const RowIn nxt = load_row(i + 1);
        const RowBlock b = pipe.block(i);
        pipe.wait(i, b);
        if ((int)threadIdx.x < b.nr) {
            const int r = b.r0 + threadIdx.x;
            const int ks = cur.k0 - b.ka, ke = cur.k1 - b.ka, len = ke - ks;
            const double* sM = pipe.f64(i % FCT_NST2, 0);
            const double* sD = pipe.f64(i % FCT_NST2, TPL ? 0 : 1);
            double tm[8];
            if (TPL) {
                // m_ij and the column offsets of the row come from its template (code fetched one block ahead)
                const double2* tp = reinterpret_cast<const double2*>(tval + 8 * cur.t);
#pragma unroll
                for (int j = 0; j < 4; ++j) { const double2 v = __ldg(tp + j); tm[2 * j] = v.x; tm[2 * j + 1] = v.y; }
            }
            const int32_t* sC = pipe.s32(i % FCT_NST2, 0);
            const double udi = cur.ud, uli = cur.ul, rpi = cur.rp, rni = cur.rn;
            double fbar = 0.0;
            if (len <= 8) {
                // two half-rows of four: keeps 16 gathers in flight without blowing the register budget
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    int c[4];
                    double udj[4], ulj[4], rpj[4], rnj[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) c[j] = (4 * h + j < len) ? sC[ks + 4 * h + j] : r;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        udj[j] = udot[c[j]]; ulj[j] = ulow[c[j]]; rpj[j] = Rpos[c[j]]; rnj[j] = Rneg[c[j]];
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int k = ks + 4 * h + j;
                        if (4 * h + j < len && c[j] != r) {
                            const double f = (TPL ? tm[4 * h + j] : sM[k]) * (udi - udj[j]) + sD[k] * (uli - ulj[j]);
                            const double alpha = (f > 0.0) ? fmin(rpi, rnj[j]) : fmin(rni, rpj[j]);
                            fbar += alpha * f;
                        }
                    }
                }
            } else {
                for (int k = ks; k < ke; ++k) {
                    const int c = sC[k];
                    if (c == r) continue;
                    const double f = sM[k] * (udi - udot[c]) + sD[k] * (uli - ulow[c]);
                    const double alpha = (f > 0.0) ? fmin(rpi, Rneg[c]) : fmin(rni, Rpos[c]);
                    fbar += alpha * f;
                }
            }
            uout[r] = uli + dt * fbar / cur.ml;
        }
        cur = nxt;
        __syncthreads();
    }
}

// out_i = sum_j mat_ij  (row_lump, helpers.py:309-328); optionally also the diagonal
__global__ void __launch_bounds__(FCT_RB)
k_row_lump(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx, const double* __restrict__ A,
           double* __restrict__ out, double* __restrict__ diag_out, int row_begin, int row_end, int64_t nnz, int cap) {
    double* sA = reinterpret_cast<double*>(fct_smem);
    int32_t* sC = reinterpret_cast<int32_t*>(sA + cap);
    FCT_FOR_ROW_BLOCKS(blk, row_begin, row_end) {
    const RowBlock b = row_block(rowptr, row_begin, row_end, blk);
    stage_f64(sA, A, b, nnz);
    stage_s32(sC, colidx, b, nnz);
    __syncthreads();
    if ((int)threadIdx.x < b.nr) {
        const int r = b.r0 + threadIdx.x;
        const int ks = rowptr[r] - b.ka, ke = rowptr[r + 1] - b.ka;
        double acc = 0.0, dg = 0.0;
        for (int k = ks; k < ke; ++k) {
            acc += sA[k];
            if (sC[k] == r) dg = sA[k];
        }
        if (out) out[r] = acc;
        if (diag_out) diag_out[r] = dg;
    }
    __syncthreads();
    }
}

// partial[blockIdx] = sum over the CTA's rows of x_i (M y)_i   (deterministic two-stage reduction)
__global__ void __launch_bounds__(FCT_RB)
k_dot_M(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx, const double* __restrict__ Mv,
        const double* __restrict__ x, const double* __restrict__ xt, const double* __restrict__ y,
        const double* __restrict__ yt, double* __restrict__ partial, int row_begin, int row_end, int64_t nnz, int cap) {
    __shared__ __align__(8) uint64_t bars[FCT_NST];
    __shared__ double sred[FCT_RB / 32];
    RowPipe<1, 1, FCT_NST> pipe{fct_smem, bars, cap, row_begin, row_end, nnz, rowptr, {Mv}, {colidx}};
    const int nmine = pipe.my_blocks();
    pipe.init();
    pipe.prologue(nmine);
    double v = 0.0;
    for (int i = 0; i < nmine; ++i) {
        pipe.prefetch(i, nmine);
        const RowBlock b = pipe.block(i);
        const bool act = (int)threadIdx.x < b.nr;
        const int r = b.r0 + threadIdx.x;
        int ks = 0, ke = 0;
        double xr = 0.0;
        if (act) {
            ks = rowptr[r] - b.ka; ke = rowptr[r + 1] - b.ka;
            xr = xt ? (x[r] - xt[r]) : x[r];
        }
        pipe.wait(i, b);
        if (act) {
            const double* sA = pipe.f64(i % FCT_NST, 0);
            const int32_t* sC = pipe.s32(i % FCT_NST, 0);
            double acc = 0.0;
            for (int k = ks; k < ke; ++k) {
                const int c = sC[k];
                const double yc = yt ? (y[c] - yt[c]) : y[c];
                acc += sA[k] * yc;
            }
            v += xr * acc;
        }
        __syncthreads();
    }
    const double s = block_sum(v, sred);
    if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

// the same reduction with M applied through its row templates (2 B per row instead of 12 B per entry): same grid, same
// row -> thread assignment, same order of operations as k_dot_M, hence the same bits
__global__ void __launch_bounds__(FCT_RB)
k_dot_M_tpl(const uint16_t* __restrict__ code, const int32_t* __restrict__ toff, const double* __restrict__ tval,
            const double* __restrict__ x, const double* __restrict__ xt, const double* __restrict__ y,
            const double* __restrict__ yt, double* __restrict__ partial, int row_begin, int row_end) {
    __shared__ double sred[FCT_RB / 32];
    const int nblk = (row_end - row_begin + FCT_RB - 1) / FCT_RB;
    double v = 0.0;
    for (int blk = blockIdx.x; blk < nblk; blk += gridDim.x) {
        const int r = row_begin + blk * FCT_RB + (int)threadIdx.x;
        if (r < row_end) {
            const double xr = xt ? (x[r] - xt[r]) : x[r];
            const int t = code[r];
            const int4 o0 = __ldg(reinterpret_cast<const int4*>(toff + FCT_TPL_W * t));
            const int4 o1 = __ldg(reinterpret_cast<const int4*>(toff + FCT_TPL_W * t) + 1);
            const int off[8] = {o0.x, o0.y, o0.z, o0.w, o1.x, o1.y, o1.z, o1.w};
            double yc[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) yc[j] = yt ? (y[r + off[j]] - yt[r + off[j]]) : y[r + off[j]];
            double acc = 0.0;
#pragma unroll
            for (int j = 0; j < 8; ++j) acc += __ldg(tval + FCT_TPL_W * t + j) * yc[j];     // padded slots: value 0
            v += xr * acc;
        }
    }
    const double s = block_sum(v, sred);
    if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

// out[slot] (+)= scale * sum(partial[0..m))  -- single block, fixed order => deterministic
__global__ void __launch_bounds__(FCT_RB)
k_reduce_partials(const double* __restrict__ partial, int m, double scale, double* __restrict__ out, int accumulate) {
    __shared__ double sred[FCT_RB / 32];
    double v = 0.0;
    for (int i = threadIdx.x; i < m; i += FCT_RB) v += partial[i];
    const double s = block_sum(v, sred);
    if (threadIdx.x == 0) *out = (accumulate ? *out : 0.0) + scale * s;
}

__global__ void k_axpby(int64_t len, double a, const double* __restrict__ x, double b, const double* __restrict__ y,
                        double* __restrict__ out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += stride)
        out[i] = a * x[i] + (y ? b * y[i] : 0.0);
}

__global__ void k_clip_axpy(int64_t len, const double* __restrict__ x, double s, const double* __restrict__ d,
                            double lo, double hi, double* __restrict__ out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += stride)
        out[i] = fmin(fmax(__dadd_rn(x[i], __dmul_rn(s, d[i])), lo), hi);      // np.clip(c + s*d): product rounded first
}

// tpos[k] = position of (j,i) for entry k = (i,j): binary search in row j
__global__ void k_build_tpos(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx, int n,
                             int32_t* __restrict__ tpos, int* __restrict__ err) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    bool has_diag = false;
    for (int k = rowptr[r]; k < rowptr[r + 1]; ++k) {
        const int c = colidx[k];
        if (c == r) has_diag = true;
        int lo = rowptr[c], hi = rowptr[c + 1] - 1, pos = -1;
        while (lo <= hi) {
            const int mid = (lo + hi) >> 1;
            const int cc = colidx[mid];
            if (cc == r) { pos = mid; break; }
            if (cc < r) lo = mid + 1; else hi = mid - 1;
        }
        if (pos < 0) { atomicExch(err, 1); pos = k; }
        tpos[k] = pos;
    }
    if (!has_diag) atomicExch(err, 2);
}

// ======================================================================================================
// host side
// ======================================================================================================
static inline size_t smem_bytes(const fct_ctx* c, int nf64, int ns32) {
    return (size_t)c->cap * (8 * (size_t)nf64 + 4 * (size_t)ns32);
}

#define LAUNCH_ROWS(ctx, kern, nf64, ns32, ...)                                                          \
    do {                                                                                                 \
        const int nb__ = fct_grid(ctx, fct_nblocks(ctx));                                                \
        if (nb__ > 0) {                                                                                  \
            kern<<<nb__, FCT_RB, smem_bytes(ctx, nf64, ns32), (ctx)->stream>>>(__VA_ARGS__);             \
            (ctx)->launches++;                                                                           \
        }                                                                                                \
    } while (0)

// Launch with the programmatic-dependent-launch attribute (see pdl_wait in fct_pipe.cuh) unless disabled or the stream is
// being captured into the Jacobi WHILE graph.
template <typename... KArgs, typename... Args>
static inline void launch_pipe(fct_ctx* ctx, void (*kern)(KArgs...), int grid, size_t smem, Args... args) {
    if (ctx->use_pdl && !ctx->capturing) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)grid);
        cfg.blockDim = dim3(FCT_RB);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = ctx->stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
    } else {
        kern<<<grid, FCT_RB, smem, ctx->stream>>>(KArgs(args)...);
    }
    ctx->launches++;
}

// TMA-ring kernels: FCT_NST stages of (nf64 fp64 + ns32 int32) staged arrays; persistent grid = SMs x resident CTAs
static inline int pipe_grid(const fct_ctx* c, int nf64) {
    const int cap = (nf64 >= 2) ? c->grid_pipe2 : c->grid_pipe1;
    const int nb = fct_nblocks(c);
    return nb < cap ? nb : cap;
}
#define LAUNCH_PIPE(ctx, kern, nf64, ns32, ...)                                                          \
    do {                                                                                                 \
        const int nb__ = pipe_grid(ctx, nf64);                                                           \
        if (nb__ > 0)                                                                                    \
            launch_pipe(ctx, kern, nb__, ((nf64) >= 2 ? FCT_NST2 : FCT_NST) * smem_bytes(ctx, nf64, ns32), __VA_ARGS__); \
    } while (0)

// k_cheb_iter / k_jacobi_sweep are instantiated for 2, 3 and 4 ring stages; the context picks one (FCT_NST env var,
// default 2) together with the matching persistent grid
#define LAUNCH_PIPE_NST(ctx, kern, ...)                                                                  \
    do {                                                                                                 \
        const int nb0__ = fct_nblocks(ctx);                                                              \
        const int nb__ = nb0__ < (ctx)->grid_nst1 ? nb0__ : (ctx)->grid_nst1;                            \
        if (nb__ > 0) {                                                                                  \
            const size_t sm__ = (size_t)(ctx)->nst1 * smem_bytes(ctx, 1, 1);                             \
            if ((ctx)->nst1 == 2) launch_pipe(ctx, kern<2>, nb__, sm__, __VA_ARGS__);                    \
            else if ((ctx)->nst1 == 4) launch_pipe(ctx, kern<4>, nb__, sm__, __VA_ARGS__);               \
            else launch_pipe(ctx, kern<3>, nb__, sm__, __VA_ARGS__);                                     \
        }                                                                                                \
    } while (0)

// k_jacobi_sweep_tpl: the ring carries one fp64 array and no index array; grid = SMs x its own occupancy
static inline void launch_jacobi_tpl(fct_ctx* ctx, const double* Lv, const double* b, const double* dinv, const double* xin,
                                     double* xout, int chk) {
    const int nb0 = fct_nblocks(ctx);
    const int nb = nb0 < ctx->grid_jtpl ? nb0 : ctx->grid_jtpl;
    if (nb <= 0) return;
    const size_t sm = (size_t)ctx->nst_jtpl * smem_bytes(ctx, 1, 0);
#define JT_ARGS ctx->rowptr, ctx->tpl_code, ctx->tpl_off, Lv, b, dinv, xin, xout, ctx->jstate, chk, ctx->row_begin, \
                ctx->row_end, ctx->cur_rb, ctx->cur_re, ctx->nnz, ctx->cap
    const bool sc = ctx->jac_mode == 2;
    static int occ6 = -1;            // FCT_JTPL_OCC6=1: 2-stage ring, 6 resident CTAs (40 registers, some spills)
    if (occ6 < 0) { const char* e = getenv("FCT_JTPL_OCC6"); occ6 = (e && atoi(e) == 1) ? 1 : 0; }
    if (occ6 && sc && ctx->nst_jtpl == 2) {
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        const int nb6 = nb0 < 6 * sms ? nb0 : 6 * sms;
        launch_pipe(ctx, k_jacobi_sweep_tpl<2, true, 6>, nb6, sm, JT_ARGS);
        return;
    }
    if (ctx->nst_jtpl == 2) { if (sc) launch_pipe(ctx, k_jacobi_sweep_tpl<2, true>, nb, sm, JT_ARGS); else launch_pipe(ctx, k_jacobi_sweep_tpl<2, false>, nb, sm, JT_ARGS); }
    else if (ctx->nst_jtpl == 4) { if (sc) launch_pipe(ctx, k_jacobi_sweep_tpl<4, true>, nb, sm, JT_ARGS); else launch_pipe(ctx, k_jacobi_sweep_tpl<4, false>, nb, sm, JT_ARGS); }
    else { if (sc) launch_pipe(ctx, k_jacobi_sweep_tpl<3, true>, nb, sm, JT_ARGS); else launch_pipe(ctx, k_jacobi_sweep_tpl<3, false>, nb, sm, JT_ARGS); }
#undef JT_ARGS
}

int fct_launch_error(fct_ctx* ctx, const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        fct_set_error("%s: kernel launch failed: %s", what, cudaGetErrorString(e));
        return 1;
    }
    return 0;
}

int fct_kernels_configure(fct_ctx* ctx) {
    // opt in to the dynamic shared memory the widest kernel needs (3 fp64 + 2 int32 staged arrays)
    const size_t worst = smem_bytes(ctx, 3, 2);
    FCT_CHECK(worst <= 200 * 1024, "row blocks need %zu B of shared memory (max row %d): unsupported pattern",
              worst, ctx->max_row);
    const int w = FCT_SMEM_OPTIN;   // opt-in ceiling only (never lowered by a later, smaller context)
    FCT_CUDA(cudaFuncSetAttribute(k_spmv, cudaFuncAttributeMaxDynamicSharedMemorySize, w));
    FCT_CUDA(cudaFuncSetAttribute(k_spmv_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, w));
    FCT_CUDA(cudaFuncSetAttribute(k_cheb_iter<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, w));
    FCT_CUDA(cudaFuncSetAttribute(k_cheb_iter<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, w));
    FCT_CUDA(cudaFuncSetAttribute(k_cheb_iter<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, w));
    FCT_CUDA(cudaFuncSetAttribute(k_low_build<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, w));
    FCT_CUDA(cudaFuncSetAttribute(k_low_build<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, w));
    FCT_CUDA(cudaFuncSetAttribute(k_art_diff, cudaFuncAttributeMaxDynamicSharedMemorySize, w));
    FCT_CUDA(cudaFuncSetAttribute(k_jacobi_sweep<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, w));
    FCT_CUDA(cudaFuncSetAttribute(k_jacobi_sweep<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, w));
    FCT_CUDA(cudaFuncSetAttribute(k_jacobi_sweep<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, w));
    FCT_CUDA(cudaFuncSetAttribute(k_jacobi_sweep_tpl<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, w));
    FCT_CUDA(cudaFuncSetAttribute(k_jacobi_sweep_tpl<3, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, w));
    FCT_CUDA(cudaFuncSetAttribute(k_jacobi_sweep_tpl<4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, w));
    FCT_CUDA(cudaFuncSetAttribute(k_jacobi_sweep_tpl<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, w));
    FCT_CUDA((cudaFuncSetAttribute(k_jacobi_sweep_tpl<2, true, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, w)));
    FCT_CUDA(cudaFuncSetAttribute(k_jacobi_sweep_tpl<3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, w));
    FCT_CUDA(cudaFuncSetAttribute(k_jacobi_sweep_tpl<4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, w));
    FCT_CUDA(cudaFuncSetAttribute(k_jacobi_sweep_gen<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, w));
    FCT_CUDA(cudaFuncSetAttribute(k_jacobi_sweep_gen<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, w));
    FCT_CUDA(cudaFuncSetAttribute(k_jacobi_sweep_gen<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, w));
    FCT_CUDA(cudaFuncSetAttribute(k_flux_limits<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, w));
    FCT_CUDA(cudaFuncSetAttribute(k_flux_apply<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, w));
    FCT_CUDA(cudaFuncSetAttribute(k_flux_limits<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, w));
    FCT_CUDA(cudaFuncSetAttribute(k_flux_apply<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, w));
    FCT_CUDA(cudaFuncSetAttribute(k_row_lump, cudaFuncAttributeMaxDynamicSharedMemorySize, w));
    FCT_CUDA(cudaFuncSetAttribute(k_dot_M, cudaFuncAttributeMaxDynamicSharedMemorySize, w));
    FCT_CHECK(FCT_NST_LOW * smem_bytes(ctx, 2, 1) + 2 * smem_bytes(ctx, 1, 0) <= (size_t)FCT_SMEM_OPTIN,
              "row blocks need %zu B of shared memory for the TMA ring (max row %d): unsupported pattern",
              FCT_NST_LOW * smem_bytes(ctx, 2, 1) + 2 * smem_bytes(ctx, 1, 0), ctx->max_row);
    cudaDeviceProp prop;
    FCT_CUDA(cudaGetDeviceProperties(&prop, ctx->device));
    int occ1 = 0, occ2 = 0;
    FCT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ1, k_spmv, FCT_RB, FCT_NST * smem_bytes(ctx, 1, 1)));
    {
        const char* e = getenv("FCT_NST");
        ctx->nst1 = (e && (atoi(e) == 3 || atoi(e) == 4)) ? atoi(e) : 2;
        int occn = 0;
        const size_t sm = (size_t)ctx->nst1 * smem_bytes(ctx, 1, 1);
        if (ctx->nst1 == 2) FCT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occn, k_cheb_iter<2>, FCT_RB, sm));
        else if (ctx->nst1 == 4) FCT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occn, k_cheb_iter<4>, FCT_RB, sm));
        else FCT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occn, k_cheb_iter<3>, FCT_RB, sm));
        FCT_CHECK(occn >= 1, "k_cheb_iter does not fit on an SM (cap=%d, stages=%d)", ctx->cap, ctx->nst1);
        const char* o = getenv("FCT_OCC");      // tuning knob: cap the resident CTAs per SM
        if (o && atoi(o) >= 1 && atoi(o) < occn) occn = atoi(o);
        ctx->grid_nst1 = prop.multiProcessorCount * occn;
        int occj = 0;
        const char* ej = getenv("FCT_NST_JTPL");
        ctx->nst_jtpl = (ej && (atoi(ej) == 2 || atoi(ej) == 4)) ? atoi(ej) : 3;
        const size_t smj = (size_t)ctx->nst_jtpl * smem_bytes(ctx, 1, 0);
        if (ctx->nst_jtpl == 2) FCT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occj, k_jacobi_sweep_tpl<2, true>, FCT_RB, smj));
        else if (ctx->nst_jtpl == 4) FCT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occj, k_jacobi_sweep_tpl<4, true>, FCT_RB, smj));
        else FCT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occj, k_jacobi_sweep_tpl<3, true>, FCT_RB, smj));
        if (occj < 1) occj = 1;
        const char* oj = getenv("FCT_OCC_JTPL");
        if (oj && atoi(oj) >= 1 && atoi(oj) < occj) occj = atoi(oj);
        ctx->grid_jtpl = prop.multiProcessorCount * occj;
    }
    FCT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ2, k_flux_apply<false>, FCT_RB, FCT_NST2 * smem_bytes(ctx, 2, 1)));
    int occ2t = 0;
    FCT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ2t, k_flux_apply<true>, FCT_RB, FCT_NST2 * smem_bytes(ctx, 1, 1)));
    ctx->grid_flux_tpl = prop.multiProcessorCount * (occ2t > 0 ? occ2t : 1);
    FCT_CHECK(occ1 >= 1 && occ2 >= 1, "TMA-ring kernels do not fit on an SM (cap=%d)", ctx->cap);
    int occl0 = 0, occl1 = 0;
    FCT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(
        &occl0, k_low_build<0>, FCT_RB, FCT_NST_LOW * smem_bytes(ctx, 1, 1) + 2 * smem_bytes(ctx, 1, 0)));
    FCT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(
        &occl1, k_low_build<1>, FCT_RB, FCT_NST_LOW * smem_bytes(ctx, 2, 1) + 2 * smem_bytes(ctx, 1, 0)));
    FCT_CHECK(occl0 >= 1 && occl1 >= 1, "k_low_build does not fit on an SM (cap=%d)", ctx->cap);
    {
        const char* ol = getenv("FCT_OCC_LOW");     // tuning knob: resident CTAs per SM of k_low_build
        if (ol && atoi(ol) >= 1) { if (atoi(ol) < occl0) occl0 = atoi(ol); if (atoi(ol) < occl1) occl1 = atoi(ol); }
    }
    ctx->grid_low[0] = prop.multiProcessorCount * occl0;
    ctx->grid_low[1] = prop.multiProcessorCount * occl1;
    ctx->grid_pipe1 = prop.multiProcessorCount * occ1;
    ctx->grid_pipe2 = prop.multiProcessorCount * occ2;
    return 0;
}

int fct_build_tpos(fct_ctx* ctx) {
    int* derr = reinterpret_cast<int*>(ctx->jstate);   // scratch word, reset afterwards
    FCT_CUDA(cudaMemsetAsync(derr, 0, sizeof(int), ctx->stream));
    k_build_tpos<<<(ctx->n + 255) / 256, 256, 0, ctx->stream>>>(ctx->rowptr, ctx->colidx, ctx->n, ctx->tpos, derr);
    ctx->launches++;
    int herr = 0;
    FCT_CUDA(cudaMemcpyAsync(&herr, derr, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    FCT_CUDA(cudaStreamSynchronize(ctx->stream));
    FCT_CUDA(cudaMemsetAsync(derr, 0, sizeof(int), ctx->stream));
    // halo rows of a partitioned pattern are truncated, so asymmetry there is expected: only a missing
    // diagonal is fatal for them; for a full (single-GPU) pattern both are errors.
    if (herr == 2) { fct_set_error("pattern is missing a diagonal entry"); return 2; }
    if (herr == 1 && ctx->row_begin == 0 && ctx->row_end == ctx->n) {
        fct_set_error("pattern is not structurally symmetric");
        return 2;
    }
    return 0;
}

int fct_cheb_iter_tpl(fct_ctx* ctx, const double* Md, const double* g, const double* ymid, const double* yold,
                      double* ynew, double omega, double dscale);       // fct_templates.cu
int fct_spmv_tpl(fct_ctx* ctx, const double* x, double alpha, double beta, const double* z, double* y);

// y = alpha A x + beta z for a per-step matrix: template columns when the context has row templates, else CSR
static int fct_spmv_any(fct_ctx* ctx, const double* A, const double* x, double alpha, double beta, const double* z, double* y) {
    // measured at 4097^2: 0.343 ms with template columns vs 0.333 ms CSR (the longer code -> offsets -> gather chain eats
    // the byte saving), so the CSR kernel stays the default; FCT_SPMV_TC=1 selects the template-column kernel
    static int use_tc = -1;
    if (use_tc < 0) { const char* e = getenv("FCT_SPMV_TC"); use_tc = (e && atoi(e) == 1) ? 1 : 0; }
    if (use_tc && ctx->tpl_count > 0 && ctx->jac_mode > 0) {
        const int nb = pipe_grid(ctx, 1);
        if (nb > 0)
            launch_pipe(ctx, k_spmv_tc, nb, FCT_NST * smem_bytes(ctx, 1, 0), ctx->rowptr, ctx->tpl_code, ctx->tpl_off, A, x, alpha,
                        beta, z, y, ctx->cur_rb, ctx->cur_re, ctx->nnz, ctx->cap);
        return 0;
    }
    LAUNCH_PIPE(ctx, k_spmv, 1, 1, ctx->rowptr, ctx->colidx, A, x, alpha, beta, z, y, ctx->cur_rb, ctx->cur_re,
                ctx->nnz, ctx->cap);
    return 0;
}

extern "C" int fct_spmv(fct_ctx* ctx, const double* A, const double* x, double alpha, double beta, const double* z,
                        double* y) {
    FCT_CHECK(ctx && A && x && y, "fct_spmv: null argument");
    FCT_CHECK(beta == 0.0 || z, "fct_spmv: beta != 0 needs z");
    if (ctx->tpl_count && A == ctx->M) {
        if (fct_spmv_tpl(ctx, x, alpha, beta, z, y)) return 1;
        return fct_launch_error(ctx, "fct_spmv");
    }
    if (fct_spmv_any(ctx, A, x, alpha, beta, z, y)) return 1;
    return fct_launch_error(ctx, "fct_spmv");
}

int fct_halo_exchange_if(fct_ctx* ctx, double* vec);   // no-op without a communicator (fct_comm.cu)
int fct_halo_exchange2_if(fct_ctx* ctx, double* v0, double* v1);
bool fct_p2p_ready(const fct_ctx* ctx);
int fct_p2p_max2_decide(fct_ctx* ctx, double rtol, int max_sweeps, int use_handle, cudaGraphConditionalHandle handle,
                        int which = 0, int kmax = 0);
int fct_halo_allreduce_max2(fct_ctx* ctx, unsigned long long* two_words);

// ChebSI with deep-halo bookkeeping.  `vb`: ring on which the right-hand side b is valid.  Iteration k runs on ring
// min(valid(y_{k-1}) - 1, vb); when that would drop below the owned rows the two live iterates are exchanged in one
// message (valid on ring `depth` again).  Returns in *vy the ring on which the result is valid.
int fct_tile_jacobi(fct_ctx* ctx, int K, const double* Lv, const double* b, const double* xin, double* xout,
                    const unsigned long long* kdev);      // fct_tile.cu
int fct_tile_cheb(fct_ctx* ctx, int K, const double* g, const double* ymid, const double* yold, double* ymid_out,
                  double* yold_out, const double* om, double dscale);

// ChebSI on the overlapped tiles (fct_tile.cu): iteration 1 is the vector kernel, iterations 2..iters run in groups of up
// to tile_kc per launch (multi-GPU: at most `depth`, one two-vector exchange per group).  Same recurrence, same weights,
// same row arithmetic as the per-iteration path below.
static int chebsi_tiles(fct_ctx* ctx, const double* Md, const double* b, double* y, int iters, double lmin, double lmax, int vb,
                        int* vy) {
    const double rho = (lmax - lmin) / (lmax + lmin);
    const double dscale = (lmin + lmax) / 2;
    std::vector<double> om((size_t)iters + 1, 0.0);
    double omega = 0.0;
    for (int k = 1; k <= iters; ++k) {
        if (k == 2) omega = 1 / (1 - rho * rho / 2);
        else omega = 1 / (1 - (omega * rho * rho) / 4);
        om[k] = omega;
    }
    const bool multi = ctx->comm != nullptr;
    const int K = ctx->depth;
    if (vb > K) vb = K;
    double* pair[2][2] = {{ctx->w[0], ctx->w[1]}, {ctx->w[2], ctx->w[5]}};
    fct_set_ring(ctx, multi ? vb : 0);
    {
        const int nb = fct_nblocks(ctx);
        if (nb > 0) {
            k_cheb_first<<<nb, FCT_RB, 0, ctx->stream>>>(b, Md, dscale, om[1], pair[0][0], ctx->cur_rb, ctx->cur_re);
            ctx->launches++;
        }
    }
    fct_set_ring(ctx, 0);
    const double* ymid = pair[0][0];
    const double* yold = nullptr;
    int cur = 0, vmid = multi ? vb : K, it = 2;
    int kmax = ctx->tile_kc;
    if (multi && kmax > K) kmax = K;
    // Multi-GPU: a launch of kk iterations needs its inputs on ring kk and leaves the iterate on ring vmid - kk.  A launch (one
    // pass over the matrix) costs more than an exchange, so the number of launches stays ceil((iters-1)/kmax); its slack is
    // spent on shorter groups that fit the rings still valid, which saves the exchange in front of them.
    int slack = multi ? ((iters - 1 + kmax - 1) / kmax) * kmax - (iters - 1) : 0;
    while (it <= iters) {
        const int rem = iters - it + 1;
        int kk = rem < kmax ? rem : kmax;
        if (multi && vmid < kk) {
            if (vmid >= 2 && kk - vmid <= slack && rem - vmid != 1) {
                slack -= kk - vmid;
                kk = vmid;
            } else {
                if (fct_halo_exchange2_if(ctx, const_cast<double*>(ymid), const_cast<double*>(yold))) return 1;
                vmid = K;
            }
        }
        if (rem - kk == 1 && kk > 2) --kk;               // never leave a single iteration for the last launch
        const bool last = (it + kk - 1 == iters);
        double* out_mid = last ? y : pair[cur ^ 1][0];
        double* out_old = last ? nullptr : pair[cur ^ 1][1];
        if (fct_tile_cheb(ctx, kk, b, ymid, yold, out_mid, out_old, &om[it], dscale)) return 1;
        ymid = out_mid; yold = out_old; cur ^= 1; it += kk;
        vmid = multi ? vmid - kk : K;
    }
    if (vy) *vy = multi ? (vmid > 0 ? vmid : 0) : K;
    return fct_launch_error(ctx, "fct_chebsi");
}

int fct_chebsi_v(fct_ctx* ctx, const double* M, const double* Md, const double* b, double* y, int32_t iters, double lmin,
                 double lmax, int vb, int* vy) {
    if (ctx->tiles_ok && ctx->cheb_tiles_ok && M == ctx->M && Md == ctx->Mdiag && ctx->cheb_mdtab && iters >= 3 && ctx->tile_kc >= 2 &&
        (!ctx->comm || (ctx->depth >= 3 && vb >= 1)))
        return chebsi_tiles(ctx, Md, b, y, iters, lmin, lmax, vb, vy);
    // helpers.py:164-180
    const double rho = (lmax - lmin) / (lmax + lmin);
    const double dscale = (lmin + lmax) / 2;
    double omega = 0.0;
    // rotating buffers: the result of iteration k lands in buf[k % 3]; the last one is redirected to y
    double* buf[3] = {ctx->w[0], ctx->w[1], ctx->w[2]};
    double* ymid = nullptr;
    double* yold = nullptr;
    const int K = ctx->depth;
    if (vb > K) vb = K;
    int vmid = K;
    for (int k = 1; k <= iters; ++k) {
        if (k == 2) omega = 1 / (1 - rho * rho / 2);
        else omega = 1 / (1 - (omega * rho * rho) / 4);
        double* ynew = (k == iters) ? y : buf[k % 3];
        if (k == 1) {
            fct_set_ring(ctx, vb);
            const int nb = fct_nblocks(ctx);
            if (nb > 0) {
                k_cheb_first<<<nb, FCT_RB, 0, ctx->stream>>>(b, Md, dscale, omega, ynew, ctx->cur_rb, ctx->cur_re);
                ctx->launches++;
            }
            vmid = vb;
        } else {
            int r = vmid - 1 < vb ? vmid - 1 : vb;
            if (r < 0) {
                if (fct_halo_exchange2_if(ctx, ymid, yold)) return 1;      // yold may be null (k == 2)
                vmid = K;
                r = K - 1 < vb ? K - 1 : vb;
                if (r < 0) r = 0;
            }
            fct_set_ring(ctx, r);
            if (ctx->tpl_count && M == ctx->M) {
                // static mass matrix with row templates: 2 B per row instead of 12 B per entry (fct_templates.cu)
                if (fct_cheb_iter_tpl(ctx, Md, b, ymid, yold, ynew, omega, dscale)) return 1;
            } else {
                LAUNCH_PIPE_NST(ctx, k_cheb_iter, ctx->rowptr, ctx->colidx, M, Md, b, ymid, yold, ynew, omega, dscale,
                                yold != nullptr, ctx->cur_rb, ctx->cur_re, ctx->nnz, ctx->cap);
            }
            vmid = r;
        }
        yold = ymid;
        ymid = ynew;
    }
    fct_set_ring(ctx, 0);
    if (vy) *vy = vmid;
    return fct_launch_error(ctx, "fct_chebsi");
}

extern "C" int fct_chebsi(fct_ctx* ctx, const double* M, const double* Md, const double* b, double* y, int32_t iters,
                          double lmin, double lmax) {
    FCT_CHECK(ctx && M && Md && b && y, "fct_chebsi: null argument");
    FCT_CHECK(iters >= 1, "fct_chebsi: iters must be >= 1");
    return fct_chebsi_v(ctx, M, Md, b, y, iters, lmin, lmax, 0, nullptr);    // b is only trusted on the owned rows
}

extern "C" int fct_artificial_diffusion(fct_ctx* ctx, const double* mat, double* D) {
    FCT_CHECK(ctx && mat && D, "fct_artificial_diffusion: null argument");
    LAUNCH_ROWS(ctx, k_art_diff, 1, 2, ctx->rowptr, ctx->colidx, ctx->tpos, mat, D, ctx->cur_rb, ctx->cur_re,
                ctx->nnz, ctx->cap);
    return fct_launch_error(ctx, "fct_artificial_diffusion");
}

extern "C" int fct_row_lump(fct_ctx* ctx, const double* mat, double* out) {
    FCT_CHECK(ctx && mat && out, "fct_row_lump: null argument");
    LAUNCH_ROWS(ctx, k_row_lump, 1, 1, ctx->rowptr, ctx->colidx, mat, out, (double*)nullptr, ctx->cur_rb,
                ctx->cur_re, ctx->nnz, ctx->cap);
    return fct_launch_error(ctx, "fct_row_lump");
}

int fct_row_lump_diag(fct_ctx* ctx, const double* mat, double* out, double* diag) {
    LAUNCH_ROWS(ctx, k_row_lump, 1, 1, ctx->rowptr, ctx->colidx, mat, out, diag, ctx->cur_rb, ctx->cur_re,
                ctx->nnz, ctx->cap);
    return fct_launch_error(ctx, "fct_row_lump_diag");
}

// Jacobi solve of Lv x = b, x holds the initial guess on entry and the result on exit (device-side early exit;
// sweeps run in pairs so that the result always lands back in x).
// One Jacobi "cycle": the sweeps that fit between two halo exchanges.  With halo depth k the iterate is valid on ring k
// after an exchange, sweep s of the cycle runs on ring k-1-s, and the cycle ends with one exchange (depth 1: the
// classic pair with an exchange after each sweep).  The stopping test (owned rows, all-reduced max) is evaluated
// after every second sweep.  `handle` != 0: the cycle is the body of a CUDA-graph WHILE node.
static int jacobi_cycle(fct_ctx* ctx, const double* Lv, const double* b, const double* dinv, double* x, double* tmp,
                        double rtol, int max_sweeps, bool p2p, int use_handle, cudaGraphConditionalHandle handle) {
#define JACOBI_LAUNCH(xin, xout, chk)                                                                              \
    do {                                                                                                           \
        if (dinv && ctx->jac_mode > 0 && ctx->tpl_count > 0)                                                       \
            launch_jacobi_tpl(ctx, Lv, b, dinv, xin, xout, chk);                                                   \
        else if (dinv)                                                                                             \
            LAUNCH_PIPE_NST(ctx, k_jacobi_sweep, ctx->rowptr, ctx->colidx, Lv, b, dinv, xin, xout, ctx->jstate, chk, \
                            ctx->row_begin, ctx->row_end, ctx->cur_rb, ctx->cur_re, ctx->nnz, ctx->cap);           \
        else                                                                                                       \
            LAUNCH_PIPE_NST(ctx, k_jacobi_sweep_gen, ctx->rowptr, ctx->colidx, Lv, b, dinv, xin, xout, ctx->jstate, \
                            chk, ctx->row_begin, ctx->row_end, ctx->cur_rb, ctx->cur_re, ctx->nnz, ctx->cap);      \
    } while (0)
    const int K = ctx->depth;
    const int npairs = K >= 2 ? K / 2 : 1;
    int rc = 0;
    for (int q = 0; q < npairs && !rc; ++q) {
        const bool last = (q == npairs - 1);
        fct_set_ring(ctx, K >= 2 ? K - 1 - 2 * q : 0);
        JACOBI_LAUNCH(x, tmp, 0);
        if (K < 2) rc |= fct_halo_exchange_if(ctx, tmp);
        fct_set_ring(ctx, K >= 2 ? K - 2 - 2 * q : 0);
        JACOBI_LAUNCH(tmp, x, 1);
        if (last) rc |= fct_halo_exchange_if(ctx, x);
        // decision: only the last one of a cycle may end the graph loop (the exchange above must have run)
        const int uh = (use_handle && last) ? 1 : 0;
        if (p2p) {
            rc |= fct_p2p_max2_decide(ctx, rtol, max_sweeps, uh, handle);
        } else {
            rc |= fct_halo_allreduce_max2(ctx, ctx->jstate);
            if (uh) k_jacobi_decide_cond<<<1, 1, 0, ctx->stream>>>(ctx->jstate, rtol, (unsigned long long)max_sweeps, handle);
            else k_jacobi_decide<<<1, 1, 0, ctx->stream>>>(ctx->jstate, rtol, (unsigned long long)max_sweeps);
            ctx->launches++;
        }
    }
    fct_set_ring(ctx, 0);
    return rc;
#undef JACOBI_LAUNCH
}

// The same cycle on the overlapped tiles (fct_tile.cu): two fused launches of K sweeps each (x -> tmp -> x), an exchange and
// a stopping test after each.  When the test after the first launch succeeds the second one returns at once and the
// solve copies tmp back (k_copy_if after the loop).
static inline int tile_kj(const fct_ctx* ctx) {
    int K = ctx->tile_kj;
    if (ctx->comm && K > ctx->depth) K = ctx->depth;
    return K;
}
static inline bool jacobi_use_tiles(const fct_ctx* ctx, const double* dinv) {
    return ctx->tiles_ok && dinv && ctx->jac_mode == 2 && tile_kj(ctx) >= 2;
}
int fct_halo_exchange_cond(fct_ctx* ctx, double* vec, const unsigned long long* cond);      // fct_comm.cu
static inline bool jacobi_tiles_deep(const fct_ctx* ctx, bool p2p) {
    return ctx->comm && p2p && ctx->depth >= 2 * tile_kj(ctx);
}
static int jacobi_cycle_tiles(fct_ctx* ctx, const double* Lv, const double* b, double* x, double* tmp, double rtol,
                              int max_sweeps, bool p2p, int use_handle, cudaGraphConditionalHandle handle) {
    const int K = tile_kj(ctx);
    // halo depth >= 2 K (peer mailboxes): the iterate is valid on ring depth-K after the first launch, which is enough for the
    // second one -- one exchange per cycle; if the solve ends after a first half, jacobi_copy_back exchanges the result
    const bool deep = jacobi_tiles_deep(ctx, p2p);
    const bool adapt = ctx->tile_adapt && ctx->tile_sched;      // launch depths from the device-side sweep schedule
    int rc = 0;
    fct_set_ring(ctx, 0);
    for (int half = 0; half < 2 && !rc; ++half) {
        double* xin = half ? tmp : x;
        double* xout = half ? x : tmp;
        rc |= fct_tile_jacobi(ctx, K, Lv, b, xin, xout, adapt ? ctx->jstate + 13 : nullptr);
        if (!(deep && half == 0)) rc |= fct_halo_exchange_if(ctx, xout);
        const int uh = (use_handle && half == 1) ? 1 : 0;
        if (p2p) {
            rc |= fct_p2p_max2_decide(ctx, rtol, max_sweeps, uh, handle, half == 0 ? 1 : 0, adapt ? K : 0);
        } else {
            rc |= fct_halo_allreduce_max2(ctx, ctx->jstate);
            k_tile_decide<<<1, 1, 0, ctx->stream>>>(ctx->jstate, rtol, (unsigned long long)max_sweeps,
                                                    half == 0 ? 1ull : 0ull, uh, handle, adapt ? (unsigned long long)K : 0ull);
            ctx->launches++;
        }
    }
    return rc;
}
static int jacobi_copy_back(fct_ctx* ctx, const double* tmp, double* x, bool p2p) {
    k_copy_if<<<148 * 4, 256, 0, ctx->stream>>>(ctx->jstate + 12, tmp, x, ctx->n);
    ctx->launches++;
    return jacobi_tiles_deep(ctx, p2p) ? fct_halo_exchange_cond(ctx, x, ctx->jstate + 12) : 0;
}

// Jacobi solve of Lv x = b; x holds the initial guess on entry (valid on every local row) and the result on exit
// (valid on every local row: the last cycle ends with an exchange).  dinv != nullptr: Lv has a zero diagonal slot and
// dinv = 1/diag (FCT low-order system); otherwise a general matrix.
int fct_jacobi_solve(fct_ctx* ctx, const double* Lv, const double* b, const double* dinv, double* x, double* tmp,
                     double rtol, int max_sweeps) {
    const int sweeps_per_cycle = ctx->depth >= 2 ? 2 * (ctx->depth / 2) : 2;
    const int cycles = (max_sweeps + sweeps_per_cycle - 1) / sweeps_per_cycle;
    const bool p2p = ctx->comm && fct_p2p_ready(ctx);
    const bool tiles = jacobi_use_tiles(ctx, dinv) && (!ctx->comm || p2p || !ctx->use_graph);
    const int jmode = tiles ? 10 + tile_kj(ctx) + ((ctx->tile_adapt && ctx->tile_sched) ? 100 : 0) : ctx->jac_mode;
    if ((!ctx->comm || p2p) && dinv && ctx->use_graph) {
        // The cycle is the body of a CUDA-graph WHILE node whose condition the decide kernel sets on the device:
        // exactly as many sweeps as needed are launched, with no host round trip and no skipped launches (multi-GPU:
        // the halo exchanges and the all-reduced stopping test are peer-memory kernels inside the same body).  The graph
        // is rebuilt only if an operand pointer or a solver option changes.
        fct_jgraph& jg = ctx->jgraph;
        if (!jg.exec || jg.Lv != Lv || jg.b != b || jg.dinv != dinv || jg.x != x || jg.tmp != tmp || jg.rtol != rtol ||
            jg.max_sweeps != max_sweeps || jg.depth != ctx->depth || jg.mode != jmode || jg.tpl != (const void*)ctx->tpl_code) {
            if (jg.exec) { cudaGraphExecDestroy((cudaGraphExec_t)jg.exec); jg.exec = nullptr; }
            if (jg.graph) { cudaGraphDestroy((cudaGraph_t)jg.graph); jg.graph = nullptr; }
            cudaGraph_t g;
            FCT_CUDA(cudaGraphCreate(&g, 0));
            cudaGraphConditionalHandle h;
            FCT_CUDA(cudaGraphConditionalHandleCreate(&h, g, 1, cudaGraphCondAssignDefault));
            cudaGraphNodeParams np = {cudaGraphNodeTypeConditional};
            np.conditional.handle = h;
            np.conditional.type = cudaGraphCondTypeWhile;
            np.conditional.size = 1;
            cudaGraphNode_t node;
            FCT_CUDA(cudaGraphAddNode(&node, g, nullptr, 0, &np));
            cudaGraph_t body = np.conditional.phGraph_out[0];
            cudaStream_t user = ctx->stream;
            ctx->stream = ctx->copy_stream;          // capture stream: launches below are recorded, not executed
            ctx->capturing = true;
            const int64_t launches0 = ctx->launches;
            FCT_CUDA(cudaStreamBeginCaptureToGraph(ctx->stream, body, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal));
            const int rc = tiles ? jacobi_cycle_tiles(ctx, Lv, b, x, tmp, rtol, max_sweeps, p2p, 1, h)
                                 : jacobi_cycle(ctx, Lv, b, dinv, x, tmp, rtol, max_sweeps, p2p, 1, h);
            cudaError_t ce = cudaStreamEndCapture(ctx->stream, nullptr);
            ctx->stream = user;
            ctx->capturing = false;
            ctx->launches = launches0;
            if (rc || ce != cudaSuccess) {
                if (!rc) fct_set_error("fct_jacobi_solve: graph capture failed: %s", cudaGetErrorString(ce));
                cudaGraphDestroy(g);
                return 1;
            }
            cudaGraphExec_t ex;
            FCT_CUDA(cudaGraphInstantiate(&ex, g, 0));
            jg.graph = g; jg.exec = ex;
            jg.Lv = Lv; jg.b = b; jg.dinv = dinv; jg.x = x; jg.tmp = tmp; jg.rtol = rtol; jg.max_sweeps = max_sweeps;
            jg.depth = ctx->depth;
            jg.mode = jmode;
            jg.tpl = ctx->tpl_code;
        }
        FCT_CUDA(cudaGraphLaunch((cudaGraphExec_t)jg.exec, ctx->stream));
        ctx->launches += 3;      // at least one body iteration; the executed sweeps are counted in jstate[4]
        if (tiles && jacobi_copy_back(ctx, tmp, x, p2p)) return 1;
        return 0;
    }
    if (tiles) {
        // static launch sequence (FCT_NO_GRAPH=1): launches after convergence return at once
        const int per = 2 * tile_kj(ctx);
        for (int c = 0; c < (max_sweeps + per - 1) / per; ++c)
            if (jacobi_cycle_tiles(ctx, Lv, b, x, tmp, rtol, max_sweeps, p2p, 0, 0)) return 1;
        if (jacobi_copy_back(ctx, tmp, x, p2p)) return 1;
        return fct_launch_error(ctx, "fct_jacobi_solve");
    }
    if (ctx->comm) {
        // Multi-GPU without peer mailboxes (NCCL): a skipped sweep would still pay its exchanges, so cycles are
        // enqueued in a budget learnt from the previous solve and the all-reduced (rank-uniform) convergence flag is
        // read back before spending more.
        int done = 0;
        int budget = ctx->last_pairs > 0 ? ctx->last_pairs : 6;
        while (done < cycles) {
            const int todo = (done + budget <= cycles) ? budget : cycles - done;
            for (int c = 0; c < todo; ++c)
                if (jacobi_cycle(ctx, Lv, b, dinv, x, tmp, rtol, max_sweeps, p2p, 0, 0)) return 1;
            done += todo;
            unsigned long long st[2] = {0, 0};     // {converged flag, sweeps executed}
            FCT_CUDA(cudaMemcpyAsync(st, ctx->jstate + 3, sizeof(st), cudaMemcpyDeviceToHost, ctx->stream));
            FCT_CUDA(cudaStreamSynchronize(ctx->stream));
            if (st[0]) { done = (int)((st[1] + sweeps_per_cycle - 1) / sweeps_per_cycle); break; }
            budget = 1;
        }
        ctx->last_pairs = done;      // sweeps per step are very stable: next time enqueue exactly this many first
        return fct_launch_error(ctx, "fct_jacobi_solve");
    }
    // single GPU, static launch sequence with device-side early exit (general matrices, or FCT_NO_GRAPH=1)
    for (int c = 0; c < cycles; ++c)
        if (jacobi_cycle(ctx, Lv, b, dinv, x, tmp, rtol, max_sweeps, false, 0, 0)) return 1;
    return fct_launch_error(ctx, "fct_jacobi_solve");
}

int fct_p2p_check(fct_ctx* ctx, const char* what);      // fct_p2p.cu

int fct_read_step_info(fct_ctx* ctx, fct_step_info* info) {
    if (fct_p2p_check(ctx, "fct_step")) return 1;
    unsigned long long h[8];
    FCT_CUDA(cudaMemcpyAsync(h, ctx->jstate, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
    FCT_CUDA(cudaStreamSynchronize(ctx->stream));
    info->solver_sweeps = (int32_t)h[4];
    info->converged = (int32_t)h[3];
    memcpy(&info->last_delta, &h[5], 8);
    memcpy(&info->x_norm, &h[6], 8);
    unsigned long long key = h[7];
    unsigned long long bits = (key >> 63) ? (key & 0x7FFFFFFFFFFFFFFFull) : ~key;
    memcpy(&info->min_rowsum_low, &bits, 8);
    return 0;
}

// ---- fallback of the low-order solve ----------------------------------------------------------------------------------
// The reference solves the low-order system with a direct solver (helpers.py:1782) and only PRINTS a diagnostic when it is
// not an M-matrix (:1796-1809).  Jacobi needs a contraction factor well below 1; beyond that (large dt, strong reaction
// terms with positive off-diagonals, ...) the solve is completed by Jacobi-preconditioned BiCGStab on the same matrix,
// starting from the Jacobi iterate.  k_low_build left the diagonal slot of L empty (the diagonal travels as 1/l_ii, or
// as 1 after row scaling): put it back first.
__global__ void k_restore_diag(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ tpos, double* __restrict__ Lv,
                               const double* __restrict__ dinv, int n) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    for (int k = rowptr[r]; k < rowptr[r + 1]; ++k)
        if (tpos[k] == k) Lv[k] = dinv ? 1.0 / dinv[r] : 1.0;
}
__global__ void k_set_word(unsigned long long* p, unsigned long long v) { *p = v; }

int fct_solve_ws(fct_ctx* ctx, int32_t kind, const double* mat, const double* b, double* x, double rtol, int32_t maxit,
                 int32_t* its_host, double* res_host, double* const* ws);      // fct_drivers.cu

static int low_order_fallback(fct_ctx* ctx, const double* bvec, const double* un, double* ulow, const double* dinv) {
    FCT_CHECK(!ctx->comm, "low-order Jacobi solve did not converge after %d sweeps and the BiCGStab fallback is single-GPU "
              "(dt violates the M-matrix condition of helpers.py:1795-1809)", ctx->max_sweeps);
    for (int i = 0; i < 12; ++i)
        if (!ctx->fb_w[i]) FCT_CUDA(cudaMalloc((void**)&ctx->fb_w[i], sizeof(double) * ((size_t)ctx->n + 8)));
    k_restore_diag<<<(ctx->n + 255) / 256, 256, 0, ctx->stream>>>(ctx->rowptr, ctx->tpos, ctx->Lvals, dinv, ctx->n);
    ctx->launches++;
    // restart from u_n: a divergent Jacobi iteration leaves an iterate that is worse than the initial guess (or not finite)
    FCT_CUDA(cudaMemcpyAsync(ulow, un, sizeof(double) * ctx->n, cudaMemcpyDeviceToDevice, ctx->stream));
    int32_t its = 0;
    double res = 0.0;
    if (fct_solve_ws(ctx, 2, ctx->Lvals, bvec, ulow, 1e-13, 5000, &its, &res, ctx->fb_w)) {
        char msg[900];
        snprintf(msg, sizeof(msg), "%s", fct_last_error());
        fct_set_error("low-order solve: Jacobi did not converge in %d sweeps and the BiCGStab fallback failed too: %s",
                      ctx->max_sweeps, msg);
        return 1;
    }
    k_set_word<<<1, 1, 0, ctx->stream>>>(ctx->jstate + 3, 1ull);      // solved: later readers see a converged step
    ctx->launches++;
    return 0;
}

int fct_drift_low_build(fct_ctx* ctx, const double* c, double bx, double by, double ascale, double sign, const double* rhs,
                        const double* un, double dt, double* bvec, double* dinv_out);     // fct_assembly.cu

// The FCT step.  drift_c != nullptr: A = ascale * drift operator of the control drift_c is not given but produced,
// together with D, L and b, by the fused assembly pass (fct_assembly.cu) into ctx->Avals; A must then be ctx->Avals.
static int fct_step_impl(fct_ctx* ctx, const double* A, double sign, const double* S, const double* rhs, const double* un,
                         double dt, double* uout, fct_step_info* info, const double* drift_c, double bx, double by,
                         double ascale);

extern "C" int fct_step(fct_ctx* ctx, const double* A, double sign, const double* S, const double* rhs,
                        const double* un, double dt, double* uout, fct_step_info* info) {
    return fct_step_impl(ctx, A, sign, S, rhs, un, dt, uout, info, nullptr, 0.0, 0.0, 0.0);
}

// One FCT step of the drift-control problem with A = ascale * [(b.grad c) u v + (b.grad v) c u] (eps = 0): assembly fused
// into the low-order build when the mesh has geometry templates, else assemble + fct_step.
int fct_step_drift(fct_ctx* ctx, const double* c, double bx, double by, double ascale, const double* rhs, const double* un,
                   double dt, double* uout) {
    return fct_step_impl(ctx, ctx->Avals, 1.0, nullptr, rhs, un, dt, uout, nullptr, c, bx, by, ascale);
}

static int fct_step_impl(fct_ctx* ctx, const double* A, double sign, const double* S, const double* rhs, const double* un,
                         double dt, double* uout, fct_step_info* info, const double* drift_c, double bx, double by,
                         double ascale) {
    FCT_CHECK(ctx && A && un && uout, "fct_step: null argument");
    FCT_CHECK(ctx->mass_set, "fct_step: static mass matrices not set (fct_ctx_set_mass / fct_assemble_static)");
    FCT_CHECK(sign == 1.0 || sign == -1.0, "fct_step: sign must be +1 (FCT_alg_ref) or -1 (FCT_alg)");
    double* bvec = ctx->w[3];
    double* ulow = ctx->w[4];
    double* tmp = ctx->w[5];
    double* g = ctx->w[6];
    double* dinv = ctx->w[6];      // 1/diag(L): dead once the low-order solve is done, before g is formed
    double* udot = ctx->w[7];
    double* Rp = ctx->w[8];
    double* Rn = ctx->w[9];
    // fused tile sweeps: the reset also starts the sweep schedule of this solve (depth of the first launch, probe)
    ctx->tile_sched = ctx->tile_adapt && ctx->in_time_loop && jacobi_use_tiles(ctx, ctx->w[6]) && (!ctx->comm || fct_p2p_ready(ctx) || !ctx->use_graph);
    k_jacobi_reset<<<1, 1, 0, ctx->stream>>>(ctx->jstate, ctx->tile_sched ? (unsigned long long)tile_kj(ctx) : 0ull);
    ctx->launches++;
    // Deep halos: inputs (u_n, rhs, A incl. transposed entries) are valid on ring K (rhs: K-1); every pass below runs
    // on the largest ring its inputs allow, so that only the Jacobi / Chebyshev cycles and the output need exchanges.
    const int K = ctx->depth;
    // 1-2. D, L, b on ring K-1
    fct_set_ring(ctx, K - 1);
    bool built = false;
    if (drift_c) {
        const int rcf = fct_drift_low_build(ctx, drift_c, bx, by, ascale, sign, rhs, un, dt, bvec,
                                            ctx->jac_mode == 2 ? nullptr : dinv);
        if (rcf == 1) {
            built = true;
        } else {
            // no geometry templates: assemble A (all local rows: the halo rows carry the transposed entries), then build
            fct_set_ring(ctx, 0);
            if (fct_assemble_matrix(ctx, FCT_FORM_DRIFT, drift_c, nullptr, nullptr, bx, by, ascale, 0, ctx->Avals)) return 1;
            fct_set_ring(ctx, K - 1);
        }
    }
    if (!built) {
        const int nf = S ? 2 : 1;
        double* dinv_out = ctx->jac_mode == 2 ? nullptr : dinv;      // the row-scaled sweeps never read 1/diag
        const size_t smem = FCT_NST_LOW * smem_bytes(ctx, nf, 1) + 2 * smem_bytes(ctx, 1, 0);
        const int nb = fct_nblocks(ctx) < ctx->grid_low[nf - 1] ? fct_nblocks(ctx) : ctx->grid_low[nf - 1];
        if (nb > 0) {
            if (S)
                k_low_build<1><<<nb, FCT_RB, smem, ctx->stream>>>(ctx->rowptr, ctx->colidx, ctx->tpos, A, sign, S, ctx->ML, un,
                                                                  rhs, dt, ctx->Lvals, ctx->Dvals, bvec, dinv_out, ctx->jstate + 7,
                                                                  ctx->jac_mode == 2, ctx->cur_rb, ctx->cur_re, ctx->nnz, ctx->cap);
            else
                k_low_build<0><<<nb, FCT_RB, smem, ctx->stream>>>(ctx->rowptr, ctx->colidx, ctx->tpos, A, sign, S, ctx->ML, un,
                                                                  rhs, dt, ctx->Lvals, ctx->Dvals, bvec, dinv_out, ctx->jstate + 7,
                                                                  ctx->jac_mode == 2, ctx->cur_rb, ctx->cur_re, ctx->nnz, ctx->cap);
            ctx->launches++;
        }
    }
    fct_set_ring(ctx, 0);
    // low-order solve, initial guess u_n (valid on every local row; so is the result)
    FCT_CUDA(cudaMemcpyAsync(ulow, un, sizeof(double) * ctx->n, cudaMemcpyDeviceToDevice, ctx->stream));
    if (fct_jacobi_solve(ctx, ctx->Lvals, bvec, dinv, ulow, tmp, ctx->rtol, ctx->max_sweeps)) return 1;
    if (info || ctx->checked_steps) {
        // host-visible step (the Python shims, or a time loop repeated after a failed step): complete an unconverged
        // Jacobi solve with BiCGStab -- the reference's direct solve has no such dt restriction
        unsigned long long conv = 1ull;
        FCT_CUDA(cudaMemcpyAsync(&conv, ctx->jstate + 3, sizeof(conv), cudaMemcpyDeviceToHost, ctx->stream));
        FCT_CUDA(cudaStreamSynchronize(ctx->stream));
        if (!conv && low_order_fallback(ctx, bvec, un, ulow, ctx->jac_mode == 2 ? nullptr : dinv)) return 1;
    }
    // 4. g = -(sign A) u_low + rhs on ring K-1; udot = ChebSI(g)
    fct_set_ring(ctx, K - 1);
    if (fct_spmv_any(ctx, A, ulow, -sign, rhs ? 1.0 : 0.0, rhs, g)) return 1;
    int vud = 0;
    if (fct_chebsi_v(ctx, ctx->M, ctx->Mdiag, g, udot, 20, 0.5, 2.0, K - 1, &vud)) return 1;
    // 5-7. fluxes, P, Q, R: on ring 1 when the halo is deep enough (then R+- needs no exchange), else on the owned rows
    const int rflux = K >= 2 ? 1 : 0;
    if (vud < rflux + 1) {
        if (fct_halo_exchange_if(ctx, udot)) return 1;
    }
    fct_set_ring(ctx, rflux);
    const bool ftpl = ctx->tpl_count > 0;
    const int gflux = fct_nblocks(ctx) < ctx->grid_flux_tpl ? fct_nblocks(ctx) : ctx->grid_flux_tpl;
    if (ftpl) {
        if (gflux > 0)
            launch_pipe(ctx, k_flux_limits<true>, gflux, FCT_NST2 * smem_bytes(ctx, 1, 0), ctx->rowptr, ctx->colidx, ctx->M,
                        ctx->tpl_code, ctx->tpl_off, ctx->tpl_val, ctx->Dvals, ctx->ML, udot, ulow, dt, Rp, Rn, ctx->cur_rb, ctx->cur_re,
                        ctx->nnz, ctx->cap);
    } else {
        LAUNCH_PIPE(ctx, k_flux_limits<false>, 2, 1, ctx->rowptr, ctx->colidx, ctx->M, ctx->tpl_code, ctx->tpl_off, ctx->tpl_val, ctx->Dvals,
                    ctx->ML, udot, ulow, dt, Rp, Rn, ctx->cur_rb, ctx->cur_re, ctx->nnz, ctx->cap);
    }
    if (K < 2) {
        if (fct_halo_exchange2_if(ctx, Rp, Rn)) return 1;
    }
    // 8-9. limited sum + update on the owned rows
    fct_set_ring(ctx, 0);
    if (ftpl) {
        const int ga = fct_nblocks(ctx) < ctx->grid_flux_tpl ? fct_nblocks(ctx) : ctx->grid_flux_tpl;
        if (ga > 0)
            launch_pipe(ctx, k_flux_apply<true>, ga, FCT_NST2 * smem_bytes(ctx, 1, 1), ctx->rowptr, ctx->colidx, ctx->M,
                        ctx->tpl_code, ctx->tpl_off, ctx->tpl_val, ctx->Dvals, ctx->ML, udot, ulow, Rp, Rn, dt, uout, ctx->cur_rb, ctx->cur_re,
                        ctx->nnz, ctx->cap);
    } else {
        LAUNCH_PIPE(ctx, k_flux_apply<false>, 2, 1, ctx->rowptr, ctx->colidx, ctx->M, ctx->tpl_code, ctx->tpl_off, ctx->tpl_val, ctx->Dvals,
                    ctx->ML, udot, ulow, Rp, Rn, dt, uout, ctx->cur_rb, ctx->cur_re, ctx->nnz, ctx->cap);
    }
    if (fct_launch_error(ctx, "fct_step")) return 1;
    if (fct_halo_exchange_if(ctx, uout)) return 1;
    if (info) return fct_read_step_info(ctx, info);
    return 0;
}

// Instrumentation for bench.py: builds the low-order system of (A, u_n, dt) and times `reps` back-to-back Jacobi sweeps
// (the FCT step's k_jacobi_sweep, stopping test accumulation on every second one) with CUDA events on the context's
// stream.  Single GPU.
extern "C" int fct_bench_jacobi_sweeps(fct_ctx* ctx, const double* A, const double* un, double dt, int32_t reps,
                                       float* ms_per_sweep_host) {
    FCT_CHECK(ctx && A && un && ms_per_sweep_host && reps >= 2, "fct_bench_jacobi_sweeps: bad argument");
    FCT_CHECK(ctx->mass_set && !ctx->comm, "fct_bench_jacobi_sweeps: single-GPU context with static matrices required");
    double* bvec = ctx->w[3];
    double* x = ctx->w[4];
    double* tmp = ctx->w[5];
    double* dinv = ctx->w[6];
    k_jacobi_reset<<<1, 1, 0, ctx->stream>>>(ctx->jstate, 0ull);
    fct_set_ring(ctx, 0);
    {
        const size_t smem = FCT_NST_LOW * smem_bytes(ctx, 1, 1) + 2 * smem_bytes(ctx, 1, 0);
        const int nb = fct_nblocks(ctx) < ctx->grid_low[0] ? fct_nblocks(ctx) : ctx->grid_low[0];
        k_low_build<0><<<nb, FCT_RB, smem, ctx->stream>>>(ctx->rowptr, ctx->colidx, ctx->tpos, A, 1.0, nullptr, ctx->ML, un,
                                                          nullptr, dt, ctx->Lvals, ctx->Dvals, bvec, dinv, ctx->jstate + 7,
                                                          ctx->jac_mode == 2, ctx->cur_rb, ctx->cur_re, ctx->nnz, ctx->cap);
    }
    FCT_CUDA(cudaMemcpyAsync(x, un, sizeof(double) * ctx->n, cudaMemcpyDeviceToDevice, ctx->stream));
    cudaEvent_t e0, e1;
    FCT_CUDA(cudaEventCreate(&e0));
    FCT_CUDA(cudaEventCreate(&e1));
    const double* Lv = ctx->Lvals;
    for (int pass = 0; pass < 2; ++pass) {          // pass 0: warm-up
        if (pass == 1) FCT_CUDA(cudaEventRecord(e0, ctx->stream));
        for (int i = 0; i < reps; ++i) {
            double* xin = (i & 1) ? tmp : x;
            double* xout = (i & 1) ? x : tmp;
            if (ctx->jac_mode > 0 && ctx->tpl_count > 0)
                launch_jacobi_tpl(ctx, Lv, bvec, dinv, xin, xout, i & 1);
            else
                LAUNCH_PIPE_NST(ctx, k_jacobi_sweep, ctx->rowptr, ctx->colidx, Lv, bvec, dinv, xin, xout, ctx->jstate, i & 1,
                                ctx->row_begin, ctx->row_end, ctx->cur_rb, ctx->cur_re, ctx->nnz, ctx->cap);
        }
    }
    FCT_CUDA(cudaEventRecord(e1, ctx->stream));
    FCT_CUDA(cudaEventSynchronize(e1));
    float ms = 0.f;
    FCT_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *ms_per_sweep_host = ms / reps;
    return fct_launch_error(ctx, "fct_bench_jacobi_sweeps");
}

// Test hook (tests/test_gpu_parity.py): builds the low-order system of (A, u_n, dt) and runs exactly `sweeps` Jacobi
// sweeps from the initial guess u_n, either one launch per sweep (fused = 0) or as fused tile launches of `fused` (2..4)
// sweeps each (fct_tile.cu; sweeps must be a multiple of `fused`); copies the iterate to x_out.  The two must agree bit
// for bit.
extern "C" int fct_debug_jacobi_fixed(fct_ctx* ctx, const double* A, const double* un, double dt, int32_t sweeps,
                                      int32_t fused, double* x_out) {
    FCT_CHECK(ctx && A && un && x_out && sweeps >= 1, "fct_debug_jacobi_fixed: bad argument");
    float dummy = 0.f;
    // (re)build L, b; the timing loops of fct_bench_jacobi_sweeps leave a scratch iterate behind, so restart from u_n
    if (fct_bench_jacobi_sweeps(ctx, A, un, dt, 2, &dummy)) return 1;
    double* x = ctx->w[4];
    double* tmp = ctx->w[5];
    FCT_CUDA(cudaMemcpyAsync(x, un, sizeof(double) * ctx->n, cudaMemcpyDeviceToDevice, ctx->stream));
    k_jacobi_reset<<<1, 1, 0, ctx->stream>>>(ctx->jstate, 0ull);
    int launches = 0;
    if (fused) {
        FCT_CHECK(ctx->tiles_ok && ctx->jac_mode == 2, "fct_debug_jacobi_fixed: tile kernels not available on this context");
        FCT_CHECK(fused >= 2 && fused <= 4 && sweeps % fused == 0, "fct_debug_jacobi_fixed: sweeps must be a multiple of fused (2..4)");
        for (int i = 0; i < sweeps / fused; ++i, ++launches)
            if (fct_tile_jacobi(ctx, fused, ctx->Lvals, ctx->w[3], (i & 1) ? tmp : x, (i & 1) ? x : tmp, nullptr)) return 1;
    } else {
        for (int i = 0; i < sweeps; ++i, ++launches) {
            double* xin = (i & 1) ? tmp : x;
            double* xout = (i & 1) ? x : tmp;
            if (ctx->jac_mode > 0 && ctx->tpl_count > 0)
                launch_jacobi_tpl(ctx, ctx->Lvals, ctx->w[3], ctx->w[6], xin, xout, 0);
            else
                LAUNCH_PIPE_NST(ctx, k_jacobi_sweep, ctx->rowptr, ctx->colidx, ctx->Lvals, ctx->w[3], ctx->w[6], xin, xout,
                                ctx->jstate, 0, ctx->row_begin, ctx->row_end, ctx->cur_rb, ctx->cur_re, ctx->nnz, ctx->cap);
        }
    }
    FCT_CUDA(cudaMemcpyAsync(x_out, (launches & 1) ? tmp : x, sizeof(double) * ctx->n, cudaMemcpyDeviceToDevice, ctx->stream));
    k_jacobi_reset<<<1, 1, 0, ctx->stream>>>(ctx->jstate, 0ull);
    return fct_launch_error(ctx, "fct_debug_jacobi_fixed");
}

// times `reps` fused tile launches of `sweeps` (2..4) Jacobi sweeps each on the low-order system of (A, u_n, dt); returns
// the CUDA-event time per SWEEP
extern "C" int fct_bench_jacobi_fused(fct_ctx* ctx, const double* A, const double* un, double dt, int32_t sweeps,
                                      int32_t reps, float* ms_per_sweep_host) {
    FCT_CHECK(ctx && A && un && ms_per_sweep_host && reps >= 1, "fct_bench_jacobi_fused: bad argument");
    FCT_CHECK(ctx->tiles_ok && ctx->jac_mode == 2 && sweeps >= 2 && sweeps <= 4,
              "fct_bench_jacobi_fused: tile kernels not available on this context (or sweeps not in 2..4)");
    float dummy = 0.f;
    if (fct_bench_jacobi_sweeps(ctx, A, un, dt, 2, &dummy)) return 1;      // builds L, b and leaves an iterate in w[4]
    double* x = ctx->w[4];
    double* tmp = ctx->w[5];
    k_jacobi_reset<<<1, 1, 0, ctx->stream>>>(ctx->jstate, 0ull);
    cudaEvent_t e0, e1;
    FCT_CUDA(cudaEventCreate(&e0));
    FCT_CUDA(cudaEventCreate(&e1));
    for (int pass = 0; pass < 2; ++pass) {          // pass 0: warm-up
        if (pass == 1) FCT_CUDA(cudaEventRecord(e0, ctx->stream));
        for (int i = 0; i < reps; ++i)
            if (fct_tile_jacobi(ctx, sweeps, ctx->Lvals, ctx->w[3], (i & 1) ? tmp : x, (i & 1) ? x : tmp, nullptr)) return 1;
    }
    FCT_CUDA(cudaEventRecord(e1, ctx->stream));
    FCT_CUDA(cudaEventSynchronize(e1));
    float ms = 0.f;
    FCT_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    k_jacobi_reset<<<1, 1, 0, ctx->stream>>>(ctx->jstate, 0ull);
    *ms_per_sweep_host = ms / (reps * sweeps);
    return fct_launch_error(ctx, "fct_bench_jacobi_fused");
}

extern "C" int fct_step_host(fct_ctx* ctx, const double* A, double sign, const double* S, const double* rhs,
                             const double* un, double dt, double* uout, fct_step_info* info) {
    FCT_CHECK(ctx && A && un && uout, "fct_step_host: null argument");
    FCT_CHECK(ctx->row_begin == 0 && ctx->row_end == ctx->n, "fct_step_host: single-GPU contexts only");
    const size_t vb = sizeof(double) * (size_t)ctx->n, mb = sizeof(double) * (size_t)ctx->nnz;
    double* d_un = ctx->w[0];     // w[0..2] are ChebSI buffers: free until fct_chebsi runs, and un is consumed before
    double* d_rhs = ctx->w[1];
    double* d_out = ctx->w[2];
    // un / rhs are needed after ChebSI starts?  un: no (only k_low_build + initial guess).  rhs: yes (k_spmv runs
    // before ChebSI).  Both are consumed before the first Chebyshev buffer is written.
    FCT_CUDA(cudaMemcpyAsync(ctx->Avals, A, mb, cudaMemcpyHostToDevice, ctx->stream));
    if (S) FCT_CUDA(cudaMemcpyAsync(ctx->Svals, S, mb, cudaMemcpyHostToDevice, ctx->stream));
    FCT_CUDA(cudaMemcpyAsync(d_un, un, vb, cudaMemcpyHostToDevice, ctx->stream));
    if (rhs) FCT_CUDA(cudaMemcpyAsync(d_rhs, rhs, vb, cudaMemcpyHostToDevice, ctx->stream));
    // ChebSI's first iteration writes buf[1] = w[1] (rhs) only after g was formed; w[2] at k == 2; w[0] at k == 3.
    // d_out = w[2] is written last by k_flux_apply, after ChebSI finished.
    // host-visible step: the low-order solve is checked and, if Jacobi ran out of sweeps, completed by BiCGStab
    const bool checked0 = ctx->checked_steps;
    ctx->checked_steps = true;
    int rc = fct_step(ctx, ctx->Avals, sign, S ? ctx->Svals : nullptr, rhs ? d_rhs : nullptr, d_un, dt, d_out, nullptr);
    ctx->checked_steps = checked0;
    if (rc) return rc;
    FCT_CUDA(cudaMemcpyAsync(uout, d_out, vb, cudaMemcpyDeviceToHost, ctx->stream));
    if (info) return fct_read_step_info(ctx, info);
    FCT_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}

extern "C" int fct_dot_M(fct_ctx* ctx, const double* M, const double* x, const double* y, double* out_host) {
    FCT_CHECK(ctx && M && x && y && out_host, "fct_dot_M: null argument");
    double* partial = ctx->w[0];
    const int nb = pipe_grid(ctx, 1);
    if (ctx->tpl_count && M == ctx->M && nb > 0) {
        k_dot_M_tpl<<<nb, FCT_RB, 0, ctx->stream>>>(ctx->tpl_code, ctx->tpl_off, ctx->tpl_val, x, (const double*)nullptr, y,
                                                    (const double*)nullptr, partial, ctx->cur_rb, ctx->cur_re);
        ctx->launches++;
    } else {
        LAUNCH_PIPE(ctx, k_dot_M, 1, 1, ctx->rowptr, ctx->colidx, M, x, (const double*)nullptr, y, (const double*)nullptr,
                    partial, ctx->cur_rb, ctx->cur_re, ctx->nnz, ctx->cap);
    }
    k_reduce_partials<<<1, FCT_RB, 0, ctx->stream>>>(partial, nb, 1.0, ctx->red, 0);
    ctx->launches++;
    if (fct_launch_error(ctx, "fct_dot_M")) return 1;
    FCT_CUDA(cudaMemcpyAsync(out_host, ctx->red, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    FCT_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int fct_allreduce_sum_host(fct_ctx* ctx, double* v);   // fct_comm.cu

extern "C" int fct_norm_sq_Q(fct_ctx* ctx, const double* M, const double* phi, const double* target,
                             int32_t num_steps, double dt, double* out_host) {
    FCT_CHECK(ctx && M && phi && out_host && num_steps >= 0, "fct_norm_sq_Q: bad argument");
    double* partial = ctx->w[0];
    const int nb = pipe_grid(ctx, 1);
    // helpers.py:354-359: sum_k w_k phi_k^T M phi_k * dt, w_0 = w_N = 1/2
    for (int k = 0; k <= num_steps; ++k) {
        const double* p = phi + (size_t)k * ctx->n;
        const double* t = target ? target + (size_t)k * ctx->n : nullptr;
        const double wk = (k == 0 || k == num_steps) ? 0.5 : 1.0;
        if (ctx->tpl_count && M == ctx->M && nb > 0) {
            k_dot_M_tpl<<<nb, FCT_RB, 0, ctx->stream>>>(ctx->tpl_code, ctx->tpl_off, ctx->tpl_val, p, t, p, t, partial,
                                                        ctx->row_begin, ctx->row_end);
            ctx->launches++;
        } else {
            LAUNCH_PIPE(ctx, k_dot_M, 1, 1, ctx->rowptr, ctx->colidx, M, p, t, p, t, partial, ctx->row_begin, ctx->row_end,
                        ctx->nnz, ctx->cap);
        }
        k_reduce_partials<<<1, FCT_RB, 0, ctx->stream>>>(partial, nb, wk, ctx->red, k > 0);
        ctx->launches++;
    }
    if (fct_launch_error(ctx, "fct_norm_sq_Q")) return 1;
    double v = 0.0;
    FCT_CUDA(cudaMemcpyAsync(&v, ctx->red, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    FCT_CUDA(cudaStreamSynchronize(ctx->stream));
    if (fct_allreduce_sum_host(ctx, &v)) return 1;
    *out_host = v * dt;
    return 0;
}

extern "C" int fct_axpby(fct_ctx* ctx, int64_t len, double a, const double* x, double b, const double* y, double* out) {
    FCT_CHECK(ctx && x && out && len >= 0, "fct_axpby: bad argument");
    if (len == 0) return 0;
    const int grid = (int)((len + 255) / 256 < 148 * 16 ? (len + 255) / 256 : 148 * 16);
    k_axpby<<<grid, 256, 0, ctx->stream>>>(len, a, x, b, y, out);
    ctx->launches++;
    return fct_launch_error(ctx, "fct_axpby");
}

extern "C" int fct_vals_axpby(fct_ctx* ctx, double a, const double* X, double b, const double* Y, double* out) {
    FCT_CHECK(ctx, "fct_vals_axpby: null context");
    return fct_axpby(ctx, ctx->nnz, a, X, b, Y, out);
}

extern "C" int fct_clip_axpy(fct_ctx* ctx, int64_t len, const double* x, double s, const double* d, double lo,
                             double hi, double* out) {
    FCT_CHECK(ctx && x && d && out && len >= 0, "fct_clip_axpy: bad argument");
    if (len == 0) return 0;
    const int grid = (int)((len + 255) / 256 < 148 * 16 ? (len + 255) / 256 : 148 * 16);
    k_clip_axpy<<<grid, 256, 0, ctx->stream>>>(len, x, s, d, lo, hi, out);
    ctx->launches++;
    return fct_launch_error(ctx, "fct_clip_axpy");
}
