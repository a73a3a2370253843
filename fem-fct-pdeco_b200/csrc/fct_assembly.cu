// P1 element assembly into the fixed CSR pattern, as a row-gather: one thread per matrix row walks the
// (<= 6 on the structured mesh) cells incident to its vertex in ascending cell order (vertex->cell incidence
// storing the two other vertices of each incident cell), evaluates the row of
// each element tensor and adds it into the row's slots held in shared memory; the finished rows leave through
// a coalesced store.  No atomics: the summation order is the cell order (what dolfin's cell loop and the
// oracle's bincount do), so results are bit-reproducible and independent of the GPU count.
//
// Replaces the dolfin `assemble(...)` calls on the reference's hot path (helpers.py:87-104 and the call
// sites listed per form in include/fctpdeco.h).
#include "fct_common.cuh"
#include "../../include/fctpdeco.h"

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_select.cuh>
#include <stdlib.h>
#include <vector>

extern __shared__ __align__(16) unsigned char fct_smem[];

// FIAT default triangle schemes (reference-triangle coordinates (xi,eta), weights sum to 1/2).
// Degree 4: Strang-Fix 6-point with the literal 15-digit constants FIAT ships (the chemotaxis golden depends on
// them); degree 5: Strang-Fix 7-point.
__constant__ double c_q4[6][3] = {
    {0.091576213509771, 0.091576213509771, 0.109951743655322 / 2},
    {0.816847572980459, 0.091576213509771, 0.109951743655322 / 2},
    {0.091576213509771, 0.816847572980459, 0.109951743655322 / 2},
    {0.445948490915965, 0.445948490915965, 0.223381589678011 / 2},
    {0.108103018168070, 0.445948490915965, 0.223381589678011 / 2},
    {0.445948490915965, 0.108103018168070, 0.223381589678011 / 2}};
__constant__ double c_q5[7][3] = {
    {1.0 / 3.0, 1.0 / 3.0, 0.225 / 2},
    {0.10128650732345633, 0.10128650732345633, 0.12593918054482717 / 2},
    {0.79742698535308720, 0.10128650732345633, 0.12593918054482717 / 2},
    {0.10128650732345633, 0.79742698535308720, 0.12593918054482717 / 2},
    {0.47014206410511505, 0.47014206410511505, 0.13239415278850616 / 2},
    {0.05971587178976981, 0.47014206410511505, 0.13239415278850616 / 2},
    {0.47014206410511505, 0.05971587178976981, 0.13239415278850616 / 2}};

struct CellGeom {
    int d[3];          // DoFs
    double gx[3], gy[3];
    double detJ, area;
    double m12;        // area / 12 (mass-matrix scale); comes out of the geometry-template table without the division
};

// Geometry of a cell seen from its vertex r.  The vertex->cell incidence stores, per (vertex, incident cell), the cell's
// two other vertices in the cell's own cyclic order (orientation preserved), so r is always local vertex 0: every
// per-vertex array is indexed with compile-time constants and stays in registers, and the cell list itself is not read.
__device__ __forceinline__ CellGeom cell_geom(const double* __restrict__ xy, int r, int j1, int j2) {
    CellGeom g;
    g.d[0] = r; g.d[1] = j1; g.d[2] = j2;
    const double2 p0 = __ldg(reinterpret_cast<const double2*>(xy) + r);
    const double2 p1 = __ldg(reinterpret_cast<const double2*>(xy) + j1);
    const double2 p2 = __ldg(reinterpret_cast<const double2*>(xy) + j2);
    const double det = (p1.x - p0.x) * (p2.y - p0.y) - (p2.x - p0.x) * (p1.y - p0.y);
    const double inv = 1.0 / det;
    g.gx[1] = (p2.y - p0.y) * inv;  g.gy[1] = -(p2.x - p0.x) * inv;
    g.gx[2] = -(p1.y - p0.y) * inv; g.gy[2] = (p1.x - p0.x) * inv;
    g.gx[0] = -(g.gx[1] + g.gx[2]); g.gy[0] = -(g.gy[1] + g.gy[2]);
    g.detJ = fabs(det);
    g.area = 0.5 * g.detJ;
    g.m12 = g.area / 12.0;
    return g;
}

struct FormArgs {
    const double* f0;
    const double* f1;
    const double* f2;
    const double* f3;
    double s0, s1;
};

// e[b] = A_e[a][b] for the local row a of cell geometry g
template <int KIND>
__device__ __forceinline__ void element_row(const CellGeom& g, int a, const FormArgs& fa, double e[3]) {
    if (KIND == FCT_FORM_MASS) {
        const double m = g.m12;
#pragma unroll
        for (int b = 0; b < 3; ++b) e[b] = (b == a) ? 2.0 * m : m;
    } else if (KIND == FCT_FORM_STIFFNESS) {
#pragma unroll
        for (int b = 0; b < 3; ++b) e[b] = g.area * (g.gx[a] * g.gx[b] + g.gy[a] * g.gy[b]);
    } else if (KIND == FCT_FORM_DRIFT_MASS) {
        const double c0 = fa.f0[g.d[0]], c1 = fa.f0[g.d[1]], c2 = fa.f0[g.d[2]];
        const double gcx = c0 * g.gx[0] + c1 * g.gx[1] + c2 * g.gx[2];
        const double gcy = c0 * g.gy[0] + c1 * g.gy[1] + c2 * g.gy[2];
        const double sm = (fa.s0 * gcx + fa.s1 * gcy) * (g.m12);
#pragma unroll
        for (int b = 0; b < 3; ++b) e[b] = sm * ((b == a) ? 2.0 : 1.0);
    } else if (KIND == FCT_FORM_DRIFT_CONV) {
        const double cc[3] = {fa.f0[g.d[0]], fa.f0[g.d[1]], fa.f0[g.d[2]]};
        const double m = g.m12;
        const double bg = fa.s0 * g.gx[a] + fa.s1 * g.gy[a];
        const double csum = (cc[0] + cc[1]) + cc[2];
#pragma unroll
        for (int b = 0; b < 3; ++b) e[b] = bg * (m * (csum + cc[b]));
    } else if (KIND == FCT_FORM_DRIFT) {
        const double c0 = fa.f0[g.d[0]], c1 = fa.f0[g.d[1]], c2 = fa.f0[g.d[2]];
        const double cc[3] = {c0, c1, c2};
        const double gcx = c0 * g.gx[0] + c1 * g.gx[1] + c2 * g.gx[2];
        const double gcy = c0 * g.gy[0] + c1 * g.gy[1] + c2 * g.gy[2];
        const double s = fa.s0 * gcx + fa.s1 * gcy;           // b . grad c
        const double m = g.m12;
        const double bg = fa.s0 * g.gx[a] + fa.s1 * g.gy[a];  // b . grad phi_a
        const double csum = (c0 + c1) + c2;
#pragma unroll
        for (int b = 0; b < 3; ++b) {
            const double mass = (s * m) * ((b == a) ? 2.0 : 1.0);
            const double conv = bg * (m * (csum + cc[b]));
            e[b] = mass + conv;
        }
    } else if (KIND == FCT_FORM_WIND_P1 || KIND == FCT_FORM_WIND_P1_T) {
        const double wx[3] = {fa.f0[g.d[0]], fa.f0[g.d[1]], fa.f0[g.d[2]]};
        const double wy[3] = {fa.f1[g.d[0]], fa.f1[g.d[1]], fa.f1[g.d[2]]};
        const double sx = (wx[0] + wx[1]) + wx[2], sy = (wy[0] + wy[1]) + wy[2];
        const double m = g.m12;
#pragma unroll
        for (int b = 0; b < 3; ++b) {
            if (KIND == FCT_FORM_WIND_P1) e[b] = g.gx[a] * (m * (sx + wx[b])) + g.gy[a] * (m * (sy + wy[b]));
            else e[b] = g.gx[b] * (m * (sx + wx[a])) + g.gy[b] * (m * (sy + wy[a]));
        }
    } else if (KIND == FCT_FORM_DIVW_MASS) {
        // div(w_h) u v with a P1 wind: div(w_h) is constant on the cell
        const double dv = (fa.f0[g.d[0]] * g.gx[0] + fa.f0[g.d[1]] * g.gx[1] + fa.f0[g.d[2]] * g.gx[2]) +
                          (fa.f1[g.d[0]] * g.gy[0] + fa.f1[g.d[1]] * g.gy[1] + fa.f1[g.d[2]] * g.gy[2]);
        const double sm = dv * g.m12;
#pragma unroll
        for (int b = 0; b < 3; ++b) e[b] = sm * ((b == a) ? 2.0 : 1.0);
    } else if (KIND == FCT_FORM_WIND_POLY3 || KIND == FCT_FORM_WIND_POLY3_T) {
        // W_b = int w phi_b with the cubic wind evaluated at the 7 quadrature points of the degree-5 rule
        const double2 p0 = __ldg(reinterpret_cast<const double2*>(fa.f1) + g.d[0]);   // fa.f1 = dof coordinates
        const double2 p1 = __ldg(reinterpret_cast<const double2*>(fa.f1) + g.d[1]);
        const double2 p2 = __ldg(reinterpret_cast<const double2*>(fa.f1) + g.d[2]);
        double Wx[3] = {0, 0, 0}, Wy[3] = {0, 0, 0};
        const double* cw = fa.f0;
        for (int q = 0; q < 7; ++q) {
            const double ph[3] = {1.0 - c_q5[q][0] - c_q5[q][1], c_q5[q][0], c_q5[q][1]};
            const double x = ph[0] * p0.x + ph[1] * p1.x + ph[2] * p2.x;
            const double y = ph[0] * p0.y + ph[1] * p1.y + ph[2] * p2.y;
            const double mono[10] = {1.0, x, y, x * x, x * y, y * y, x * x * x, x * x * y, x * y * y, y * y * y};
            double wx = 0.0, wy = 0.0;
#pragma unroll
            for (int m = 0; m < 10; ++m) { wx += __ldg(cw + m) * mono[m]; wy += __ldg(cw + 10 + m) * mono[m]; }
            const double w = c_q5[q][2] * g.detJ;
#pragma unroll
            for (int b = 0; b < 3; ++b) { Wx[b] += w * wx * ph[b]; Wy[b] += w * wy * ph[b]; }
        }
#pragma unroll
        for (int b = 0; b < 3; ++b) {
            if (KIND == FCT_FORM_WIND_POLY3) e[b] = g.gx[a] * Wx[b] + g.gy[a] * Wy[b];
            else e[b] = g.gx[b] * Wx[a] + g.gy[b] * Wy[a];
        }
    } else if (KIND == FCT_FORM_WMASS1 || KIND == FCT_FORM_WMASS2 || KIND == FCT_FORM_WMASS3) {
        e[0] = e[1] = e[2] = 0.0;
        const double a0[3] = {fa.f0[g.d[0]], fa.f0[g.d[1]], fa.f0[g.d[2]]};
        double a1[3] = {1, 1, 1}, a2[3] = {1, 1, 1};
        if (KIND >= FCT_FORM_WMASS2) { a1[0] = fa.f1[g.d[0]]; a1[1] = fa.f1[g.d[1]]; a1[2] = fa.f1[g.d[2]]; }
        if (KIND >= FCT_FORM_WMASS3) { a2[0] = fa.f2[g.d[0]]; a2[1] = fa.f2[g.d[1]]; a2[2] = fa.f2[g.d[2]]; }
        for (int q = 0; q < 7; ++q) {
            const double ph[3] = {1.0 - c_q5[q][0] - c_q5[q][1], c_q5[q][0], c_q5[q][1]};
            double v = a0[0] * ph[0] + a0[1] * ph[1] + a0[2] * ph[2];
            if (KIND >= FCT_FORM_WMASS2) v *= a1[0] * ph[0] + a1[1] * ph[1] + a1[2] * ph[2];
            if (KIND >= FCT_FORM_WMASS3) v *= a2[0] * ph[0] + a2[1] * ph[1] + a2[2] * ph[2];
            const double w = c_q5[q][2] * v * ph[a];
#pragma unroll
            for (int b = 0; b < 3; ++b) e[b] += w * ph[b];
        }
#pragma unroll
        for (int b = 0; b < 3; ++b) e[b] *= g.detJ;
    } else if (KIND == FCT_FORM_CHTX || KIND == FCT_FORM_CHTX_EXP) {
        const double f0 = fa.f0[g.d[0]], f1 = fa.f0[g.d[1]], f2 = fa.f0[g.d[2]];
        const double gfx = f0 * g.gx[0] + f1 * g.gx[1] + f2 * g.gx[2];
        const double gfy = f0 * g.gy[0] + f1 * g.gy[1] + f2 * g.gy[2];
        const double s = gfx * g.gx[a] + gfy * g.gy[a];
        if (KIND == FCT_FORM_CHTX) {
            e[0] = e[1] = e[2] = s * (g.area / 3.0);
        } else {
            const double m0 = fa.f1[g.d[0]], m1 = fa.f1[g.d[1]], m2 = fa.f1[g.d[2]];
            double W[3] = {0, 0, 0};
            for (int q = 0; q < 6; ++q) {
                const double ph[3] = {1.0 - c_q4[q][0] - c_q4[q][1], c_q4[q][0], c_q4[q][1]};
                const double mq = m0 * ph[0] + m1 * ph[1] + m2 * ph[2];
                const double w = c_q4[q][2] * exp(-fa.s0 * mq);
#pragma unroll
                for (int b = 0; b < 3; ++b) W[b] += w * ph[b];
            }
#pragma unroll
            for (int b = 0; b < 3; ++b) e[b] = s * (W[b] * g.detJ);
        }
    } else if (KIND == FCT_FORM_CHTX_ADJ) {
        const double v0 = fa.f0[g.d[0]], v1 = fa.f0[g.d[1]], v2 = fa.f0[g.d[2]];
        const double gvx = v0 * g.gx[0] + v1 * g.gx[1] + v2 * g.gx[2];
        const double gvy = v0 * g.gy[0] + v1 * g.gy[1] + v2 * g.gy[2];
        const double u0 = fa.f1[g.d[0]], u1 = fa.f1[g.d[1]], u2 = fa.f1[g.d[2]];
        double Wa = 0.0;
        for (int q = 0; q < 7; ++q) {
            const double ph[3] = {1.0 - c_q5[q][0] - c_q5[q][1], c_q5[q][0], c_q5[q][1]};
            const double uq = u0 * ph[0] + u1 * ph[1] + u2 * ph[2];
            Wa += c_q5[q][2] * ((1.0 - fa.s0 * uq) * exp(-fa.s0 * uq)) * ph[a];
        }
        Wa *= g.detJ;
#pragma unroll
        for (int b = 0; b < 3; ++b) e[b] = Wa * (g.gx[b] * gvx + g.gy[b] * gvy);
    }
}

template <int KIND>
__global__ void __launch_bounds__(FCT_RB)
k_assemble_matrix(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                  const int32_t* __restrict__ v2c_ptr, const int32_t* __restrict__ v2c_idx,
                  const int32_t* __restrict__ cells, const double* __restrict__ xy, FormArgs fa, double scale,
                  int accumulate, double* __restrict__ out, int row_begin, int row_end, int64_t nnz, int cap) {
    double* sV = reinterpret_cast<double*>(fct_smem);
    double* sO = sV + cap;                                   // previous values when accumulating
    int32_t* sC = reinterpret_cast<int32_t*>(sO + (accumulate ? cap : 0));
    FCT_FOR_ROW_BLOCKS(blk, row_begin, row_end) {
    const RowBlock b = row_block(rowptr, row_begin, row_end, blk);
    stage_s32(sC, colidx, b, nnz);
    if (accumulate) stage_f64(sO, out, b, nnz);
    __syncthreads();
    if ((int)threadIdx.x < b.nr) {
        const int r = b.r0 + threadIdx.x;
        const int ks = rowptr[r] - b.ka, ke = rowptr[r + 1] - b.ka;
        int kd = ks;
        for (int k = ks; k < ke; ++k) {
            sV[k] = 0.0;
            if (sC[k] == r) kd = k;
        }
        const int cs = v2c_ptr[r], ce = v2c_ptr[r + 1];
        for (int ci = cs; ci < ce; ++ci) {
            const int2 nb = __ldg(reinterpret_cast<const int2*>(v2c_idx) + ci);
            const CellGeom g = cell_geom(xy, r, nb.x, nb.y);
            double e[3];
            element_row<KIND>(g, 0, fa, e);
            sV[kd] += e[0];
#pragma unroll
            for (int q = 1; q < 3; ++q) {
                const int col = g.d[q];
                for (int k = ks; k < ke; ++k)
                    if (sC[k] == col) { sV[k] += e[q]; break; }
            }
        }
        for (int k = ks; k < ke; ++k) sV[k] = accumulate ? (sO[k] + scale * sV[k]) : (scale * sV[k]);
    }
    __syncthreads();
    unstage_f64(out, sV, b);
    __syncthreads();
    }
}

// Register path for patterns whose rows have at most 8 entries (every P1 row of the structured meshes): the row's
// column indices and accumulators live in registers, matching is a predicated add per slot, and nothing is staged in
// shared memory -- the whole L1 serves the cell / coordinate / coefficient gathers, which is what bounds this kernel.
// Same summation order (ascending cell index) as the staged kernel, hence bit-identical results.
template <int KIND>
__global__ void __launch_bounds__(FCT_RB)
k_assemble_matrix_r8(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                     const int32_t* __restrict__ v2c_ptr, const int32_t* __restrict__ v2c_idx,
                     const int32_t* __restrict__ cells, const double* __restrict__ xy, FormArgs fa, double scale,
                     int accumulate, double* __restrict__ out, int row_begin, int row_end) {
    const int stride = (int)gridDim.x * FCT_RB;
    for (int r = row_begin + (int)blockIdx.x * FCT_RB + (int)threadIdx.x; r < row_end; r += stride) {
        const int ks = rowptr[r];
        const int len = rowptr[r + 1] - ks;
        int c[8];
        double acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            c[j] = (j < len) ? __ldg(colidx + ks + j) : -1;
            acc[j] = 0.0;
        }
        const int cs = v2c_ptr[r], ce = v2c_ptr[r + 1];
        for (int ci = cs; ci < ce; ++ci) {
            const int2 nb = __ldg(reinterpret_cast<const int2*>(v2c_idx) + ci);
            const CellGeom g = cell_geom(xy, r, nb.x, nb.y);
            double e[3];
            element_row<KIND>(g, 0, fa, e);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const double add = (c[j] == g.d[0]) ? e[0] : ((c[j] == g.d[1]) ? e[1] : ((c[j] == g.d[2]) ? e[2] : 0.0));
                acc[j] += add;
            }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j)
            if (j < len) out[(int64_t)ks + j] = accumulate ? (out[(int64_t)ks + j] + scale * acc[j]) : (scale * acc[j]);
    }
}

// ---- linear forms ------------------------------------------------------------------------------------
template <int KIND>
__device__ __forceinline__ double element_load(const CellGeom& g, int a, const FormArgs& fa) {
    if (KIND == FCT_LOAD_CONST) {
        return fa.s0 * (g.area / 3.0);
    } else if (KIND == FCT_LOAD_DRIFT_GRAD) {
        const double u0 = fa.f1[g.d[0]], u1 = fa.f1[g.d[1]], u2 = fa.f1[g.d[2]];
        const double gux = u0 * g.gx[0] + u1 * g.gx[1] + u2 * g.gx[2];
        const double guy = u0 * g.gy[0] + u1 * g.gy[1] + u2 * g.gy[2];
        const double s = fa.s0 * gux + fa.s1 * guy;
        const double p[3] = {fa.f0[g.d[0]], fa.f0[g.d[1]], fa.f0[g.d[2]]};
        return (s * (g.m12)) * (((p[0] + p[1]) + p[2]) + p[a]);
    } else if (KIND == FCT_LOAD_CHTX_ADJ) {
        const double p0 = fa.f0[g.d[0]], p1 = fa.f0[g.d[1]], p2 = fa.f0[g.d[2]];
        const double gpx = p0 * g.gx[0] + p1 * g.gx[1] + p2 * g.gx[2];
        const double gpy = p0 * g.gy[0] + p1 * g.gy[1] + p2 * g.gy[2];
        const double u0 = fa.f1[g.d[0]], u1 = fa.f1[g.d[1]], u2 = fa.f1[g.d[2]];
        double I = 0.0;
        for (int q = 0; q < 6; ++q) {
            const double ph[3] = {1.0 - c_q4[q][0] - c_q4[q][1], c_q4[q][0], c_q4[q][1]};
            const double uq = u0 * ph[0] + u1 * ph[1] + u2 * ph[2];
            I += c_q4[q][2] * (fa.s1 * uq * exp(-fa.s0 * uq));
        }
        return (gpx * g.gx[a] + gpy * g.gy[a]) * (I * g.detJ);
    } else if (KIND == FCT_LOAD_POLY3) {
        // int p(x, y) phi_a with a polynomial of degree <= 3 (10 monomial coefficients in fa.f0; fa.f1 = DoF coordinates),
        // 7-point degree-5 rule: exact
        const double2 p0 = __ldg(reinterpret_cast<const double2*>(fa.f1) + g.d[0]);
        const double2 p1 = __ldg(reinterpret_cast<const double2*>(fa.f1) + g.d[1]);
        const double2 p2 = __ldg(reinterpret_cast<const double2*>(fa.f1) + g.d[2]);
        double acc = 0.0;
        for (int q = 0; q < 7; ++q) {
            const double ph[3] = {1.0 - c_q5[q][0] - c_q5[q][1], c_q5[q][0], c_q5[q][1]};
            const double x = ph[0] * p0.x + ph[1] * p1.x + ph[2] * p2.x;
            const double y = ph[0] * p0.y + ph[1] * p1.y + ph[2] * p2.y;
            const double mono[10] = {1.0, x, y, x * x, x * y, y * y, x * x * x, x * x * y, x * y * y, y * y * y};
            double pv = 0.0;
#pragma unroll
            for (int m = 0; m < 10; ++m) pv += __ldg(fa.f0 + m) * mono[m];
            acc += c_q5[q][2] * pv * ph[a];
        }
        return acc * g.detJ;
    } else {   // FCT_LOAD_P1_1..4: product of 1..4 P1 fields, 7-point degree-5 rule
        const double* fs[4] = {fa.f0, fa.f1, fa.f2, fa.f3};
        const int nf = KIND - FCT_LOAD_P1_1 + 1;
        double v[4][3];
        for (int i = 0; i < nf; ++i) { v[i][0] = fs[i][g.d[0]]; v[i][1] = fs[i][g.d[1]]; v[i][2] = fs[i][g.d[2]]; }
        double acc = 0.0;
        for (int q = 0; q < 7; ++q) {
            const double ph[3] = {1.0 - c_q5[q][0] - c_q5[q][1], c_q5[q][0], c_q5[q][1]};
            double t = 1.0;
            for (int i = 0; i < nf; ++i) t *= v[i][0] * ph[0] + v[i][1] * ph[1] + v[i][2] * ph[2];
            acc += c_q5[q][2] * t * ph[a];
        }
        return acc * g.detJ;
    }
}

template <int KIND>
__global__ void __launch_bounds__(FCT_RB)
k_assemble_vector(const int32_t* __restrict__ v2c_ptr, const int32_t* __restrict__ v2c_idx,
                  const int32_t* __restrict__ cells, const double* __restrict__ xy, FormArgs fa, double scale,
                  int accumulate, double* __restrict__ out, int row_begin, int row_end) {
    const int r = row_begin + blockIdx.x * FCT_RB + threadIdx.x;
    if (r >= row_end) return;
    double acc = 0.0;
    const int cs = v2c_ptr[r], ce = v2c_ptr[r + 1];
    for (int ci = cs; ci < ce; ++ci) {
        const int2 nb = __ldg(reinterpret_cast<const int2*>(v2c_idx) + ci);
        const CellGeom g = cell_geom(xy, r, nb.x, nb.y);
        acc += element_load<KIND>(g, 0, fa);
    }
    out[r] = accumulate ? (out[r] + scale * acc) : (scale * acc);
}

// ======================================================================================================
// Geometry templates.  On meshes with repeated element shapes the per-row assembly input -- which cells touch the
// vertex, where their other two vertices sit relative to it, the cells' gradients and areas, and which row slot each
// local entry goes to -- is the same for most rows.  At fct_ctx_set_mesh the rows are hashed on the device, distinct
// signatures (<= 65535, <= 6 incident cells, rows of <= 8 entries) become templates and every row gets a 16-bit code.
// The templated kernels then read 2 B per row instead of the incidence lists, the coordinates of three vertices per
// cell and the column indices (1.8 GB per assembly at 4097^2); the cell geometry in the table is what cell_geom()
// computes, the element arithmetic and the summation order are unchanged, so results are bit-identical to the
// generic kernels.  A mesh that does not compress keeps the generic kernels.
#define GT_MAXC 6
struct __align__(16) GeomCell {
    int off1, off2;        // the cell's other two vertices, relative to the row's vertex
    int slots;             // row slots of (d0, d1, d2): bits 0-7, 8-15, 16-23
    int pad;
    double gx[3], gy[3], detJ, m12;
};
struct __align__(16) GeomTpl {
    int ncell, len, pad0, pad1;
    GeomCell c[GT_MAXC];
};

__device__ __forceinline__ unsigned long long gmix(unsigned long long h, unsigned long long v) {
    h ^= v + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
    h *= 0xFF51AFD7ED558CCDull;
    h ^= h >> 33;
    return h;
}

// the signature of row r, as a template (valid == false: more cells / entries than a template holds)
__device__ __forceinline__ bool geom_signature(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                                               const int32_t* __restrict__ v2c_ptr, const int32_t* __restrict__ v2c_idx,
                                               const double* __restrict__ xy, int r, GeomTpl& T) {
    const int ks = rowptr[r], len = rowptr[r + 1] - ks;
    const int cs = v2c_ptr[r], nc = v2c_ptr[r + 1] - cs;
    T.ncell = nc; T.len = len; T.pad0 = 0; T.pad1 = 0;
    if (nc > GT_MAXC || len > 8) return false;
    for (int q = 0; q < GT_MAXC; ++q) {
        GeomCell& c = T.c[q];
        c.off1 = 0; c.off2 = 0; c.slots = 0; c.pad = 0;
        for (int k = 0; k < 3; ++k) { c.gx[k] = 0.0; c.gy[k] = 0.0; }
        c.detJ = 0.0; c.m12 = 0.0;
        if (q >= nc) continue;
        const int2 nb = __ldg(reinterpret_cast<const int2*>(v2c_idx) + cs + q);
        const CellGeom g = cell_geom(xy, r, nb.x, nb.y);
        int sl[3] = {255, 255, 255};
        for (int j = 0; j < len; ++j) {
            const int col = colidx[ks + j];
            for (int k = 0; k < 3; ++k) if (col == g.d[k]) sl[k] = j;
        }
        if (sl[0] == 255 || sl[1] == 255 || sl[2] == 255) return false;     // a cell vertex missing from the row's pattern
        c.off1 = nb.x - r; c.off2 = nb.y - r;
        c.slots = sl[0] | (sl[1] << 8) | (sl[2] << 16);
        for (int k = 0; k < 3; ++k) { c.gx[k] = g.gx[k]; c.gy[k] = g.gy[k]; }
        c.detJ = g.detJ; c.m12 = g.m12;
    }
    return true;
}

__device__ __forceinline__ unsigned long long geom_hash_of(const GeomTpl& T) {
    unsigned long long h = gmix(0xABCDEFull, ((unsigned long long)(unsigned)T.ncell << 32) | (unsigned)T.len);
    for (int q = 0; q < GT_MAXC; ++q) {
        const GeomCell& c = T.c[q];
        h = gmix(h, ((unsigned long long)(unsigned)c.off1 << 32) | (unsigned)c.off2);
        h = gmix(h, (unsigned long long)(unsigned)c.slots);
        for (int k = 0; k < 3; ++k) {
            h = gmix(h, (unsigned long long)__double_as_longlong(c.gx[k]));
            h = gmix(h, (unsigned long long)__double_as_longlong(c.gy[k]));
        }
        h = gmix(h, (unsigned long long)__double_as_longlong(c.detJ));
    }
    return h;
}

__global__ void k_geom_hash(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                            const int32_t* __restrict__ v2c_ptr, const int32_t* __restrict__ v2c_idx,
                            const double* __restrict__ xy, int n, unsigned long long* __restrict__ hash, int* __restrict__ bad) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    GeomTpl T;
    if (!geom_signature(rowptr, colidx, v2c_ptr, v2c_idx, xy, r, T)) { atomicAdd(bad, 1); hash[r] = 0ull; return; }
    hash[r] = geom_hash_of(T);
}

__global__ void k_geom_codes(const unsigned long long* __restrict__ hash, const unsigned long long* __restrict__ uniq, int T,
                             int n, uint16_t* __restrict__ code, int* __restrict__ rep) {
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

__global__ void k_geom_fill(const int* __restrict__ rep, const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                            const int32_t* __restrict__ v2c_ptr, const int32_t* __restrict__ v2c_idx,
                            const double* __restrict__ xy, int T, GeomTpl* __restrict__ tab) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    GeomTpl G;
    geom_signature(rowptr, colidx, v2c_ptr, v2c_idx, xy, rep[t], G);
    tab[t] = G;
}

// exact check: every row must reproduce its template bit for bit
__global__ void k_geom_verify(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                              const int32_t* __restrict__ v2c_ptr, const int32_t* __restrict__ v2c_idx,
                              const double* __restrict__ xy, int n, const uint16_t* __restrict__ code,
                              const GeomTpl* __restrict__ tab, int* __restrict__ bad) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    GeomTpl G;
    bool ok = geom_signature(rowptr, colidx, v2c_ptr, v2c_idx, xy, r, G);
    const GeomTpl& T = tab[code[r]];
    ok = ok && G.ncell == T.ncell && G.len == T.len;
    for (int q = 0; q < GT_MAXC && ok; ++q) {
        const GeomCell &a = G.c[q], &b = T.c[q];
        ok = a.off1 == b.off1 && a.off2 == b.off2 && a.slots == b.slots &&
             __double_as_longlong(a.detJ) == __double_as_longlong(b.detJ);
        for (int k = 0; k < 3 && ok; ++k)
            ok = __double_as_longlong(a.gx[k]) == __double_as_longlong(b.gx[k]) &&
                 __double_as_longlong(a.gy[k]) == __double_as_longlong(b.gy[k]);
    }
    if (!ok) atomicAdd(bad, 1);
}

__device__ __forceinline__ CellGeom geom_from_tpl(const GeomCell* __restrict__ c, int r) {
    CellGeom g;
    const int4 h = __ldg(reinterpret_cast<const int4*>(c));
    g.d[0] = r; g.d[1] = r + h.x; g.d[2] = r + h.y;
    const double* gd = reinterpret_cast<const double*>(c) + 2;
    const double2 a = __ldg(reinterpret_cast<const double2*>(gd));
    const double2 b = __ldg(reinterpret_cast<const double2*>(gd) + 1);
    const double2 cc = __ldg(reinterpret_cast<const double2*>(gd) + 2);
    g.gx[0] = a.x; g.gx[1] = a.y; g.gx[2] = b.x;
    g.gy[0] = b.y; g.gy[1] = cc.x; g.gy[2] = cc.y;
    const double2 dm = __ldg(reinterpret_cast<const double2*>(gd) + 3);
    g.detJ = dm.x;
    g.area = 0.5 * g.detJ;
    g.m12 = dm.y;
    return g;
}

// Matrix assembly on geometry templates: thread per row, accumulators in registers, rows leave through shared memory
// as coalesced stores.
template <int KIND>
__global__ void __launch_bounds__(FCT_RB)
k_assemble_matrix_tpl(const int32_t* __restrict__ rowptr, const uint16_t* __restrict__ gcode, const GeomTpl* __restrict__ gtab,
                      FormArgs fa, double scale, int accumulate, double* __restrict__ out, int row_begin, int row_end,
                      int64_t nnz) {
    double* sV = reinterpret_cast<double*>(fct_smem);
    FCT_FOR_ROW_BLOCKS(blk, row_begin, row_end) {
        const RowBlock b = row_block(rowptr, row_begin, row_end, blk);
        if ((int)threadIdx.x < b.nr) {
            const int r = b.r0 + threadIdx.x;
            const int ks = rowptr[r];
            const int len = rowptr[r + 1] - ks;
            const GeomTpl* T = gtab + gcode[r];
            const int nc = __ldg(&T->ncell);
            // the row's slots live in shared memory (odd row stride: conflict-free); contributions are added in cell order
            double* sr = sV + (ks - b.ka);
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (j < len) sr[j] = 0.0;
            // Fixed trip count, unconditional body: unused cells of a template are zero-filled (offsets 0, zero geometry ->
            // zero element row added to slot 0), so the table reads and coefficient gathers of three cells are in flight
            // together instead of one dependent chain per cell.
            (void)nc;
#pragma unroll 3
            for (int q = 0; q < GT_MAXC; ++q) {
                const CellGeom g = geom_from_tpl(&T->c[q], r);
                const int sl = __ldg(&T->c[q].slots);
                double e[3];
                element_row<KIND>(g, 0, fa, e);
                sr[sl & 255] += e[0];
                sr[(sl >> 8) & 255] += e[1];
                sr[(sl >> 16) & 255] += e[2];
            }
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (j < len) sr[j] = accumulate ? (out[(int64_t)ks + j] + scale * sr[j]) : (scale * sr[j]);
        }
        __syncthreads();
        unstage_f64(out, sV, b);
        __syncthreads();
    }
}

template <int KIND>
__global__ void __launch_bounds__(FCT_RB)
k_assemble_vector_tpl(const uint16_t* __restrict__ gcode, const GeomTpl* __restrict__ gtab, FormArgs fa, double scale,
                      int accumulate, double* __restrict__ out, int row_begin, int row_end) {
    const int stride = (int)gridDim.x * FCT_RB;
    for (int r = row_begin + (int)blockIdx.x * FCT_RB + (int)threadIdx.x; r < row_end; r += stride) {
        const GeomTpl* T = gtab + gcode[r];
        const int nc = __ldg(&T->ncell);
        double acc = 0.0;
        (void)nc;
#pragma unroll 3
        for (int q = 0; q < GT_MAXC; ++q) {          // unused cells: zero geometry -> zero contribution
            const CellGeom g = geom_from_tpl(&T->c[q], r);
            acc += element_load<KIND>(g, 0, fa);
        }
        out[r] = accumulate ? (out[r] + scale * acc) : (scale * acc);
    }
}

// Fused drift-operator assembly + low-order build (the state / adjoint loops of the drift-control problem,
// advection_solidbody_FCT_PDECO_alltime.py:210-259 -> helpers.py:1769-1780): one pass produces
//   A   = ascale * [ (b.grad c) u v + (b.grad v) c u ]        (row i: a_ij, written for the later g = -A u_low + rhs)
//   D, L (row-scaled), b                                        (exactly what k_low_build makes of A)
// A thread walks the cells of its row once (geometry templates) and evaluates, per cell, the element row of its own
// vertex AND the element column of its own vertex: a_ij and a_ji come from the same two cells in the same (ascending
// cell) order in which row j itself would add them, so the transposed entries need no second pass over A, no tpos and
// no column indices.  (a_ji is evaluated in row i's rotation of the cell; on meshes where cell_geom() is not exact the
// last bit may differ from row j's own value -- inside the parity tolerance, and independent of the GPU count.)
__global__ void __launch_bounds__(FCT_RB)
k_drift_low_build(const int32_t* __restrict__ rowptr, const uint16_t* __restrict__ gcode, const GeomTpl* __restrict__ gtab,
                  const double* __restrict__ cf, double bx, double by, double ascale, double sign,
                  const double* __restrict__ ML, const double* __restrict__ un, const double* __restrict__ rhs, double dt,
                  double* __restrict__ Av, double* __restrict__ Lv, double* __restrict__ Dv, double* __restrict__ bvec,
                  double* __restrict__ dinv, unsigned long long* __restrict__ min_rowsum_key, int scale_rows, int row_begin,
                  int row_end, int64_t nnz, int cap) {
    __shared__ double sred[FCT_RB / 32];
    double* sA = reinterpret_cast<double*>(fct_smem);      // a_ij, then the scaled row of A
    double* sT = sA + cap;                                  // a_ji, then the row of L
    double* sD = sT + cap;
    double rowsum = 1e300;
    FCT_FOR_ROW_BLOCKS(blk, row_begin, row_end) {
        const RowBlock b = row_block(rowptr, row_begin, row_end, blk);
        if ((int)threadIdx.x < b.nr) {
            const int r = b.r0 + threadIdx.x;
            const int ks = rowptr[r];
            const int len = rowptr[r + 1] - ks;
            const GeomTpl* T = gtab + gcode[r];
            double* ra = sA + (ks - b.ka);
            double* rt = sT + (ks - b.ka);
            double* rd = sD + (ks - b.ka);
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (j < len) { ra[j] = 0.0; rt[j] = 0.0; }
            const double c0 = cf[r];
            const int kd = __ldg(&T->c[0].slots) & 255;          // every incident cell maps the row's vertex to the diagonal slot
#pragma unroll 3
            for (int q = 0; q < GT_MAXC; ++q) {                   // unused cells: zero geometry -> zero contributions
                const CellGeom g = geom_from_tpl(&T->c[q], r);
                const int sl = __ldg(&T->c[q].slots);
                const double c1 = cf[g.d[1]], c2 = cf[g.d[2]];
                // element tensor of FCT_FORM_DRIFT, same expressions as element_row<FCT_FORM_DRIFT>
                const double gcx = c0 * g.gx[0] + c1 * g.gx[1] + c2 * g.gx[2];
                const double gcy = c0 * g.gy[0] + c1 * g.gy[1] + c2 * g.gy[2];
                const double s = bx * gcx + by * gcy;
                const double m = g.m12;
                const double csum = (c0 + c1) + c2;
                const double cc[3] = {c0, c1, c2};
                const double bg0 = bx * g.gx[0] + by * g.gy[0];
                double er[3];
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const double mass = (s * m) * ((k == 0) ? 2.0 : 1.0);
                    const double conv = bg0 * (m * (csum + cc[k]));
                    er[k] = mass + conv;
                }
                // element column of the row's vertex: A_e[a][0], a = 1, 2 (what rows j1, j2 add to their entry towards r)
                const double convc = m * (csum + c0);
                const double ec1 = (s * m) * 1.0 + (bx * g.gx[1] + by * g.gy[1]) * convc;
                const double ec2 = (s * m) * 1.0 + (bx * g.gx[2] + by * g.gy[2]) * convc;
                const int s0 = sl & 255, s1 = (sl >> 8) & 255, s2 = (sl >> 16) & 255;
                ra[s0] += er[0]; ra[s1] += er[1]; ra[s2] += er[2];
                rt[s1] += ec1; rt[s2] += ec2;
            }
            // low-order operator of this row (k_low_build, helpers.py:1769-1780)
            const double ml = ML[r], unr = un[r];
            double dsum = 0.0, lsum = 0.0;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                if (j < len) {
                    const double av = ascale * ra[j];
                    ra[j] = av;                                   // A as fct_assemble_matrix(.., scale = ascale) stores it
                    if (j != kd) {
                        const double a = sign * av;
                        const double at = sign * (ascale * rt[j]);
                        const double d = fmax(0.0, fmax(a, at));
                        dsum += d;
                        const double l = dt * (a - d);
                        lsum += l;
                        rt[j] = l;
                        rd[j] = d;
                    }
                }
            }
            const double a = sign * ra[kd];
            const double dii = -dsum;
            const double l = ml + dt * (a - dii);
            lsum += l;
            rt[kd] = 0.0;
            const double di = 1.0 / l;
            if (dinv) dinv[r] = di;
            rd[kd] = dii;
            rowsum = fmin(rowsum, lsum);
            double br = ml * unr + (rhs ? dt * rhs[r] : 0.0);
            if (scale_rows) {
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    if (j < len) rt[j] *= di;
                br *= di;
            }
            bvec[r] = br;
        }
        __syncthreads();
        unstage_f64(Av, sA, b);
        unstage_f64(Lv, sT, b);
        unstage_f64(Dv, sD, b);
        __syncthreads();
    }
    const double mrs = block_min(rowsum, sred);
    if (threadIdx.x == 0) {
        // same sortable-key transform as k_low_build (atomicMin on signed doubles)
        unsigned long long u = (unsigned long long)__double_as_longlong(mrs);
        u = (u >> 63) ? ~u : (u | 0x8000000000000000ull);
        atomicMin(min_rowsum_key, u);
    }
}

// Host side of the fused pass on the context's current ring.  Returns 0 when the fused kernel is not applicable
// (no geometry templates, FCT_NO_FUSED_DRIFT=1): the caller assembles A and lets fct_step build the low-order system.
int fct_drift_low_build(fct_ctx* ctx, const double* c, double bx, double by, double ascale, double sign, const double* rhs,
                        const double* un, double dt, double* bvec, double* dinv_out) {
    const char* e = getenv("FCT_NO_FUSED_DRIFT");         // read per call: the tests toggle it between contexts
    if ((e && atoi(e) == 1) || ctx->gt_count <= 0 || ctx->max_row > 8) return 0;
    const size_t smem = 3 * (size_t)ctx->cap * 8;
    if (smem > (size_t)FCT_SMEM_OPTIN) return 0;
    const int nb = fct_grid(ctx, fct_nblocks(ctx));
    if (nb > 0) {
        k_drift_low_build<<<nb, FCT_RB, smem, ctx->stream>>>(ctx->rowptr, ctx->gt_code, reinterpret_cast<const GeomTpl*>(ctx->gt_tab),
                                                             c, bx, by, ascale, sign, ctx->ML, un, rhs, dt, ctx->Avals, ctx->Lvals,
                                                             ctx->Dvals, bvec, dinv_out, ctx->jstate + 7, ctx->jac_mode == 2,
                                                             ctx->cur_rb, ctx->cur_re, ctx->nnz, ctx->cap);
        ctx->launches++;
    }
    return 1;
}

// ======================================================================================================
template <int KIND>
static int launch_matrix(fct_ctx* ctx, const FormArgs& fa, double scale, int accumulate, double* out) {
    // all local rows, halo rows included: their entries (j,i) towards owned rows i are complete because every
    // cell containing an owned vertex is local, and those are the only halo-row values the FCT step reads (a_ji).
    if (ctx->gt_count > 0) {
        const int nbr = fct_grid(ctx, (ctx->n + FCT_RB - 1) / FCT_RB);
        k_assemble_matrix_tpl<KIND><<<nbr, FCT_RB, (size_t)ctx->cap * 8, ctx->stream>>>(
            ctx->rowptr, ctx->gt_code, reinterpret_cast<const GeomTpl*>(ctx->gt_tab), fa, scale, accumulate, out, 0, ctx->n,
            ctx->nnz);
        ctx->launches++;
        return fct_launch_error(ctx, "fct_assemble_matrix");
    }
    if (ctx->max_row <= 8) {
        const int nbr = fct_grid(ctx, (ctx->n + FCT_RB - 1) / FCT_RB);
        k_assemble_matrix_r8<KIND><<<nbr, FCT_RB, 0, ctx->stream>>>(ctx->rowptr, ctx->colidx, ctx->v2c_ptr, ctx->v2c_idx,
                                                                    ctx->cells, ctx->xy, fa, scale, accumulate, out, 0,
                                                                    ctx->n);
        ctx->launches++;
        return fct_launch_error(ctx, "fct_assemble_matrix");
    }
    const int nb = fct_grid(ctx, (ctx->n + FCT_RB - 1) / FCT_RB);
    const size_t smem = (size_t)ctx->cap * (8 * (accumulate ? 2 : 1) + 4);
    k_assemble_matrix<KIND><<<nb, FCT_RB, smem, ctx->stream>>>(ctx->rowptr, ctx->colidx, ctx->v2c_ptr, ctx->v2c_idx,
                                                               ctx->cells, ctx->xy, fa, scale, accumulate, out,
                                                               0, ctx->n, ctx->nnz, ctx->cap);
    ctx->launches++;
    return fct_launch_error(ctx, "fct_assemble_matrix");
}

template <int KIND>
static int configure_matrix(int bytes) {
    FCT_CUDA(cudaFuncSetAttribute(k_assemble_matrix<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    FCT_CUDA(cudaFuncSetAttribute(k_assemble_matrix_tpl<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    return 0;
}

int fct_assembly_configure(fct_ctx* ctx) {
    const int bytes = FCT_SMEM_OPTIN; (void)ctx;
    int rc = 0;
    rc |= configure_matrix<FCT_FORM_MASS>(bytes);
    rc |= configure_matrix<FCT_FORM_STIFFNESS>(bytes);
    rc |= configure_matrix<FCT_FORM_DRIFT>(bytes);
    rc |= configure_matrix<FCT_FORM_WIND_P1>(bytes);
    rc |= configure_matrix<FCT_FORM_WIND_P1_T>(bytes);
    rc |= configure_matrix<FCT_FORM_WMASS1>(bytes);
    rc |= configure_matrix<FCT_FORM_WMASS2>(bytes);
    rc |= configure_matrix<FCT_FORM_WMASS3>(bytes);
    rc |= configure_matrix<FCT_FORM_CHTX>(bytes);
    rc |= configure_matrix<FCT_FORM_CHTX_EXP>(bytes);
    rc |= configure_matrix<FCT_FORM_CHTX_ADJ>(bytes);
    rc |= configure_matrix<FCT_FORM_WIND_POLY3>(bytes);
    rc |= configure_matrix<FCT_FORM_WIND_POLY3_T>(bytes);
    rc |= configure_matrix<FCT_FORM_DRIFT_MASS>(bytes);
    rc |= configure_matrix<FCT_FORM_DRIFT_CONV>(bytes);
    rc |= configure_matrix<FCT_FORM_DIVW_MASS>(bytes);
    if (cudaFuncSetAttribute(k_drift_low_build, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes) != cudaSuccess) rc |= 1;
    return rc;
}

void fct_geom_templates_free(fct_ctx* ctx) {
    cudaFree(ctx->gt_code); cudaFree(ctx->gt_tab);
    ctx->gt_code = nullptr; ctx->gt_tab = nullptr; ctx->gt_count = 0;
}

// (Re)build the geometry templates of the context's mesh.  Never fails the caller: on any problem (too many
// templates, a vertex with more than GT_MAXC cells, a hash collision caught by the exact verification pass,
// FCT_NO_GEOM_TPL=1) the generic assembly kernels stay in use.
static int fct_geom_templates_build(fct_ctx* ctx) {
    fct_geom_templates_free(ctx);
    const char* e = getenv("FCT_NO_GEOM_TPL");
    if (e && atoi(e) == 1) return 0;
    if (ctx->max_row > 8) return 0;
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
        if (cudaMalloc((void**)&bad, sizeof(int)) != cudaSuccess) break;
        cudaMemsetAsync(bad, 0, sizeof(int), st);
        k_geom_hash<<<(n + 127) / 128, 128, 0, st>>>(ctx->rowptr, ctx->colidx, ctx->v2c_ptr, ctx->v2c_idx, ctx->xy, n, hash, bad);
        cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, hash, sorted, n, 0, 64, st);
        cub::DeviceSelect::Unique(nullptr, tb2, sorted, uniq, dT, n, st);
        if (tb2 > tmp_bytes) tmp_bytes = tb2;
        if (cudaMalloc(&tmp, tmp_bytes) != cudaSuccess) break;
        cub::DeviceRadixSort::SortKeys(tmp, tmp_bytes, hash, sorted, n, 0, 64, st);
        cub::DeviceSelect::Unique(tmp, tmp_bytes, sorted, uniq, dT, n, st);
        if (cudaMemcpyAsync(&T, dT, sizeof(int), cudaMemcpyDeviceToHost, st) != cudaSuccess) break;
        if (cudaMemcpyAsync(&hbad, bad, sizeof(int), cudaMemcpyDeviceToHost, st) != cudaSuccess) break;
        if (cudaStreamSynchronize(st) != cudaSuccess) break;
        if (hbad != 0 || T < 1 || T > 65535) break;
        if (cudaMalloc((void**)&rep, sizeof(int) * (size_t)T) != cudaSuccess) break;
        if (cudaMalloc((void**)&ctx->gt_code, sizeof(uint16_t) * ((size_t)n + 8)) != cudaSuccess) break;
        if (cudaMalloc((void**)&ctx->gt_tab, sizeof(GeomTpl) * (size_t)T) != cudaSuccess) break;
        cudaMemsetAsync(rep, 0x7f, sizeof(int) * (size_t)T, st);
        k_geom_codes<<<(n + 255) / 256, 256, 0, st>>>(hash, uniq, T, n, ctx->gt_code, rep);
        k_geom_fill<<<(T + 127) / 128, 128, 0, st>>>(rep, ctx->rowptr, ctx->colidx, ctx->v2c_ptr, ctx->v2c_idx, ctx->xy, T,
                                                     reinterpret_cast<GeomTpl*>(ctx->gt_tab));
        k_geom_verify<<<(n + 127) / 128, 128, 0, st>>>(ctx->rowptr, ctx->colidx, ctx->v2c_ptr, ctx->v2c_idx, ctx->xy, n,
                                                       ctx->gt_code, reinterpret_cast<const GeomTpl*>(ctx->gt_tab), bad);
        ctx->launches += 4;
        if (cudaMemcpyAsync(&hbad, bad, sizeof(int), cudaMemcpyDeviceToHost, st) != cudaSuccess) break;
        if (cudaStreamSynchronize(st) != cudaSuccess) break;
        ok = (hbad == 0);
    } while (0);
    cudaFree(hash); cudaFree(sorted); cudaFree(uniq); cudaFree(dT); cudaFree(rep); cudaFree(bad); cudaFree(tmp);
    cudaGetLastError();
    if (ok) ctx->gt_count = T;
    else fct_geom_templates_free(ctx);
    return 0;
}

extern "C" int fct_ctx_set_mesh(fct_ctx* ctx, int64_t ncells, const int32_t* cell_dofs, const double* dof_xy) {
    FCT_CHECK(ctx && cell_dofs && dof_xy && ncells >= 1, "fct_ctx_set_mesh: bad argument");
    FCT_CHECK(3 * ncells < 2147483647LL, "fct_ctx_set_mesh: too many cells for int32 incidence");
    const int n = ctx->n;
    // vertex -> incident cells, ascending cell index (counting sort)
    std::vector<int32_t> ptr((size_t)n + 1, 0);
    for (int64_t i = 0; i < 3 * ncells; ++i) {
        const int32_t d = cell_dofs[i];
        FCT_CHECK(d >= 0 && d < n, "fct_ctx_set_mesh: cell %lld references DoF %d outside [0,%d)", (long long)(i / 3), d, n);
        ptr[(size_t)d + 1]++;
    }
    for (int i = 0; i < n; ++i) ptr[(size_t)i + 1] += ptr[i];
    // per incidence: the cell's two other vertices in cyclic order (ascending cell index per vertex)
    std::vector<int32_t> idx((size_t)(6 * ncells));
    {
        std::vector<int32_t> fill(ptr.begin(), ptr.end() - 1);
        for (int64_t c = 0; c < ncells; ++c)
            for (int q = 0; q < 3; ++q) {
                const size_t slot = (size_t)fill[cell_dofs[3 * c + q]]++;
                idx[2 * slot] = cell_dofs[3 * c + (q + 1) % 3];
                idx[2 * slot + 1] = cell_dofs[3 * c + (q + 2) % 3];
            }
    }
    cudaFree(ctx->cells); cudaFree(ctx->xy); cudaFree(ctx->v2c_ptr); cudaFree(ctx->v2c_idx);
    ctx->cells = nullptr; ctx->xy = nullptr; ctx->v2c_ptr = nullptr; ctx->v2c_idx = nullptr;
    FCT_CUDA(cudaMalloc((void**)&ctx->cells, sizeof(int32_t) * 3 * (size_t)ncells));
    FCT_CUDA(cudaMalloc((void**)&ctx->xy, sizeof(double) * 2 * (size_t)n));
    FCT_CUDA(cudaMalloc((void**)&ctx->v2c_ptr, sizeof(int32_t) * ((size_t)n + 1)));
    FCT_CUDA(cudaMalloc((void**)&ctx->v2c_idx, sizeof(int32_t) * 6 * (size_t)ncells));
    FCT_CUDA(cudaMemcpy(ctx->cells, cell_dofs, sizeof(int32_t) * 3 * (size_t)ncells, cudaMemcpyHostToDevice));
    FCT_CUDA(cudaMemcpy(ctx->xy, dof_xy, sizeof(double) * 2 * (size_t)n, cudaMemcpyHostToDevice));
    FCT_CUDA(cudaMemcpy(ctx->v2c_ptr, ptr.data(), sizeof(int32_t) * ((size_t)n + 1), cudaMemcpyHostToDevice));
    FCT_CUDA(cudaMemcpy(ctx->v2c_idx, idx.data(), sizeof(int32_t) * 6 * (size_t)ncells, cudaMemcpyHostToDevice));
    ctx->ncells = ncells;
    return fct_geom_templates_build(ctx);
}

extern "C" int fct_assemble_matrix(fct_ctx* ctx, int32_t kind, const double* c0, const double* c1, const double* c2,
                                   double s0, double s1, double scale, int32_t accumulate, double* out) {
    FCT_CHECK(ctx && out, "fct_assemble_matrix: null argument");
    FCT_CHECK(ctx->cells, "fct_assemble_matrix: no mesh set (fct_ctx_set_mesh)");
    FormArgs fa{c0, c1, c2, nullptr, s0, s1};
    const int acc = accumulate ? 1 : 0;
    switch (kind) {
        case FCT_FORM_MASS: return launch_matrix<FCT_FORM_MASS>(ctx, fa, scale, acc, out);
        case FCT_FORM_STIFFNESS: return launch_matrix<FCT_FORM_STIFFNESS>(ctx, fa, scale, acc, out);
        case FCT_FORM_DRIFT:
            FCT_CHECK(c0, "fct_assemble_matrix(DRIFT): coef0 (control) required");
            return launch_matrix<FCT_FORM_DRIFT>(ctx, fa, scale, acc, out);
        case FCT_FORM_DRIFT_MASS:
            FCT_CHECK(c0, "fct_assemble_matrix(DRIFT_MASS): coef0 (control) required");
            return launch_matrix<FCT_FORM_DRIFT_MASS>(ctx, fa, scale, acc, out);
        case FCT_FORM_DRIFT_CONV:
            FCT_CHECK(c0, "fct_assemble_matrix(DRIFT_CONV): coef0 (control) required");
            return launch_matrix<FCT_FORM_DRIFT_CONV>(ctx, fa, scale, acc, out);
        case FCT_FORM_WIND_P1:
            FCT_CHECK(c0 && c1, "fct_assemble_matrix(WIND_P1): coef0/coef1 (wind components) required");
            return launch_matrix<FCT_FORM_WIND_P1>(ctx, fa, scale, acc, out);
        case FCT_FORM_WIND_P1_T:
            FCT_CHECK(c0 && c1, "fct_assemble_matrix(WIND_P1_T): coef0/coef1 (wind components) required");
            return launch_matrix<FCT_FORM_WIND_P1_T>(ctx, fa, scale, acc, out);
        case FCT_FORM_DIVW_MASS:
            FCT_CHECK(c0 && c1, "fct_assemble_matrix(DIVW_MASS): coef0/coef1 (wind components) required");
            return launch_matrix<FCT_FORM_DIVW_MASS>(ctx, fa, scale, acc, out);
        case FCT_FORM_WMASS1:
            FCT_CHECK(c0, "fct_assemble_matrix(WMASS1): coef0 required");
            return launch_matrix<FCT_FORM_WMASS1>(ctx, fa, scale, acc, out);
        case FCT_FORM_WMASS2:
            FCT_CHECK(c0 && c1, "fct_assemble_matrix(WMASS2): coef0, coef1 required");
            return launch_matrix<FCT_FORM_WMASS2>(ctx, fa, scale, acc, out);
        case FCT_FORM_WMASS3:
            FCT_CHECK(c0 && c1 && c2, "fct_assemble_matrix(WMASS3): coef0..2 required");
            return launch_matrix<FCT_FORM_WMASS3>(ctx, fa, scale, acc, out);
        case FCT_FORM_CHTX:
            FCT_CHECK(c0, "fct_assemble_matrix(CHTX): coef0 required");
            return launch_matrix<FCT_FORM_CHTX>(ctx, fa, scale, acc, out);
        case FCT_FORM_CHTX_EXP:
            FCT_CHECK(c0 && c1, "fct_assemble_matrix(CHTX_EXP): coef0, coef1 required");
            return launch_matrix<FCT_FORM_CHTX_EXP>(ctx, fa, scale, acc, out);
        case FCT_FORM_CHTX_ADJ:
            FCT_CHECK(c0 && c1, "fct_assemble_matrix(CHTX_ADJ): coef0, coef1 required");
            return launch_matrix<FCT_FORM_CHTX_ADJ>(ctx, fa, scale, acc, out);
        case FCT_FORM_WIND_POLY3:
        case FCT_FORM_WIND_POLY3_T: {
            FCT_CHECK(c0, "fct_assemble_matrix(WIND_POLY3): coef0 (20 polynomial coefficients) required");
            FormArgs fw{c0, ctx->xy, nullptr, nullptr, s0, s1};      // f1 carries the coordinates
            return kind == FCT_FORM_WIND_POLY3 ? launch_matrix<FCT_FORM_WIND_POLY3>(ctx, fw, scale, acc, out)
                                               : launch_matrix<FCT_FORM_WIND_POLY3_T>(ctx, fw, scale, acc, out);
        }
        default: break;
    }
    fct_set_error("fct_assemble_matrix: unknown form kind %d", kind);
    return 2;
}

template <int KIND>
static int launch_vector(fct_ctx* ctx, const FormArgs& fa, double scale, int accumulate, double* out) {
    const int nb = fct_nblocks(ctx);
    if (nb <= 0) return 0;
    if (ctx->gt_count > 0) {
        k_assemble_vector_tpl<KIND><<<fct_grid(ctx, nb), FCT_RB, 0, ctx->stream>>>(
            ctx->gt_code, reinterpret_cast<const GeomTpl*>(ctx->gt_tab), fa, scale, accumulate, out, ctx->cur_rb, ctx->cur_re);
        ctx->launches++;
        return fct_launch_error(ctx, "fct_assemble_vector");
    }
    k_assemble_vector<KIND><<<nb, FCT_RB, 0, ctx->stream>>>(ctx->v2c_ptr, ctx->v2c_idx, ctx->cells, ctx->xy, fa, scale,
                                                            accumulate, out, ctx->cur_rb, ctx->cur_re);
    ctx->launches++;
    return fct_launch_error(ctx, "fct_assemble_vector");
}

extern "C" int fct_assemble_vector(fct_ctx* ctx, int32_t kind, const double* c0, const double* c1, const double* c2,
                                   const double* c3, double s0, double s1, double scale, int32_t accumulate,
                                   double* out) {
    FCT_CHECK(ctx && out, "fct_assemble_vector: null argument");
    FCT_CHECK(ctx->cells, "fct_assemble_vector: no mesh set (fct_ctx_set_mesh)");
    FormArgs fa{c0, c1, c2, c3, s0, s1};
    const int acc = accumulate ? 1 : 0;
    switch (kind) {
        case FCT_LOAD_P1_1:
            FCT_CHECK(c0, "fct_assemble_vector(P1_1): coef0 required");
            return launch_vector<FCT_LOAD_P1_1>(ctx, fa, scale, acc, out);
        case FCT_LOAD_P1_2:
            FCT_CHECK(c0 && c1, "fct_assemble_vector(P1_2): coef0, coef1 required");
            return launch_vector<FCT_LOAD_P1_2>(ctx, fa, scale, acc, out);
        case FCT_LOAD_P1_3:
            FCT_CHECK(c0 && c1 && c2, "fct_assemble_vector(P1_3): coef0..2 required");
            return launch_vector<FCT_LOAD_P1_3>(ctx, fa, scale, acc, out);
        case FCT_LOAD_P1_4:
            FCT_CHECK(c0 && c1 && c2 && c3, "fct_assemble_vector(P1_4): coef0..3 required");
            return launch_vector<FCT_LOAD_P1_4>(ctx, fa, scale, acc, out);
        case FCT_LOAD_CONST: return launch_vector<FCT_LOAD_CONST>(ctx, fa, scale, acc, out);
        case FCT_LOAD_DRIFT_GRAD:
            FCT_CHECK(c0 && c1, "fct_assemble_vector(DRIFT_GRAD): coef0 (p), coef1 (u) required");
            return launch_vector<FCT_LOAD_DRIFT_GRAD>(ctx, fa, scale, acc, out);
        case FCT_LOAD_CHTX_ADJ:
            FCT_CHECK(c0 && c1, "fct_assemble_vector(CHTX_ADJ): coef0 (p), coef1 (u) required");
            return launch_vector<FCT_LOAD_CHTX_ADJ>(ctx, fa, scale, acc, out);
        case FCT_LOAD_POLY3: {
            FCT_CHECK(c0, "fct_assemble_vector(POLY3): coef0 (10 polynomial coefficients) required");
            FormArgs fp{c0, ctx->xy, nullptr, nullptr, s0, s1};      // f1 carries the coordinates
            return launch_vector<FCT_LOAD_POLY3>(ctx, fp, scale, acc, out);
        }
        default: break;
    }
    fct_set_error("fct_assemble_vector: unknown form kind %d", kind);
    return 2;
}

int fct_row_lump_diag(fct_ctx* ctx, const double* mat, double* out, double* diag);
int fct_halo_exchange_if(fct_ctx* ctx, double* vec);
int fct_templates_build(fct_ctx* ctx);

extern "C" int fct_assemble_static(fct_ctx* ctx) {
    FCT_CHECK(ctx, "fct_assemble_static: null context");
    FCT_CHECK(ctx->cells, "fct_assemble_static: no mesh set (fct_ctx_set_mesh)");
    FormArgs fa{nullptr, nullptr, nullptr, nullptr, 0.0, 0.0};
    if (launch_matrix<FCT_FORM_MASS>(ctx, fa, 1.0, 0, ctx->M)) return 1;
    if (launch_matrix<FCT_FORM_STIFFNESS>(ctx, fa, 1.0, 0, ctx->K)) return 1;
    if (fct_row_lump_diag(ctx, ctx->M, ctx->ML, ctx->Mdiag)) return 1;
    if (fct_halo_exchange_if(ctx, ctx->ML)) return 1;
    if (fct_halo_exchange_if(ctx, ctx->Mdiag)) return 1;
    ctx->mass_set = true;
    return fct_templates_build(ctx);
}
