/* Oracle, C/OpenMP twin.  TEST INFRASTRUCTURE (see oracle/__init__.py): only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may build, load or call this file.  It is never linked into the
 * product library.
 *
 * A plain-C restatement of the same algorithm oracle/fct_numpy.py and oracle/p1assembly.py restate in numpy, so that
 * the CPU baseline of bench.py can run the *named* workload (BASELINE config 5, 4097^2 DoF) at full size on all host
 * cores instead of a scaled sample:
 *   mesh / DoF numbering / CSR pattern     SURVEY.md App. B  (RectangleMesh "right", CG1 anti-diagonal numbering;
 *                                          advection_solidbody_FCT.py:48-50,82; helpers.py:87-104)
 *   mass matrix, drift operator            advection_solidbody_FCT_PDECO_alltime.py:143-147,222-226 (u v dx;
 *                                          (b.grad c) u v + (b.grad v) c u, P1 control c)
 *   one FCT step                           helpers.py:1715-1872 (FCT_alg_ref); the low-order system is solved by Jacobi
 *                                          sweeps (the CPU twin of the GPU solver; the reference uses SuperLU, :1782,
 *                                          which is impractical at this size -- BASELINE.md)
 *   ChebSI                                 helpers.py:143-185
 *   state loop                             advection_solidbody_FCT_PDECO_alltime.py:210-228
 * Pinned by tests/test_oracle_c.py against the numpy oracle (which is pinned on the reference's goldens): mesh arrays
 * bit for bit, trajectories to 1e-12.
 *
 * gcc -O3 -fopenmp -shared -fPIC oracle/fct_c.c -o oracle/_fct_c.so   (oracle/build_c.py)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

int fctc_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* bench.py's CPU arm sets the thread count itself: a torchrun launch exports OMP_NUM_THREADS=1 to every rank, which would
 * silently turn the "all host threads" baseline into a one-thread run */
void fctc_set_threads(int32_t t) {
#ifdef _OPENMP
    if (t >= 1) omp_set_num_threads(t);
#else
    (void)t;
#endif
}

/* ---- mesh, DoF numbering, pattern ------------------------------------------------------------ */
static int64_t diag_len(int64_t d, int64_t n) { return (d < 2 * n - d ? d : 2 * n - d) + 1; }

/* v2d[(n+1)^2], cells[2 n^2][3] (DoF-indexed), xy[(n+1)^2][2] (DoF order), rowptr[nodes+1], colidx[nnz] */
void fctc_mesh(int32_t n, double a1, double a2, int32_t* v2d, int32_t* cells, double* xy, int32_t* rowptr, int32_t* colidx) {
    const int64_t N = (int64_t)n + 1, nodes = N * N;
    int64_t* start = (int64_t*)malloc(sizeof(int64_t) * (size_t)(2 * n + 2));
    start[0] = 0;
    for (int64_t d = 0; d <= 2 * (int64_t)n; ++d) start[d + 1] = start[d] + diag_len(d, n);
    const double h = (a2 - a1) / n;
#pragma omp parallel for schedule(static)
    for (int64_t iy = 0; iy < N; ++iy)
        for (int64_t ix = 0; ix < N; ++ix) {
            const int64_t d = ix - iy + n;
            const int64_t dof = start[d] + (d <= n ? ix : iy);
            v2d[iy * N + ix] = (int32_t)dof;
            xy[2 * dof] = a1 + h * (double)ix;
            xy[2 * dof + 1] = a1 + h * (double)iy;
        }
#pragma omp parallel for schedule(static)
    for (int64_t iy = 0; iy < n; ++iy)
        for (int64_t ix = 0; ix < n; ++ix) {
            const int64_t v0 = iy * N + ix, v1 = v0 + 1, v2 = v0 + N, v3 = v2 + 1;
            int32_t* c = cells + 6 * (iy * n + ix);
            c[0] = v2d[v0]; c[1] = v2d[v1]; c[2] = v2d[v3];          /* (v0, v1, v3) */
            c[3] = v2d[v0]; c[4] = v2d[v2]; c[5] = v2d[v3];          /* (v0, v2, v3) */
        }
    /* pattern: vertex (ix,iy) couples to itself, (ix+-1,iy), (ix,iy+-1), (ix+1,iy+1), (ix-1,iy-1); columns ascending */
    int32_t* cnt = (int32_t*)calloc((size_t)nodes + 1, sizeof(int32_t));
    static const int dx[7] = {0, 1, -1, 0, 0, 1, -1}, dy[7] = {0, 0, 0, 1, -1, 1, -1};
#pragma omp parallel for schedule(static)
    for (int64_t iy = 0; iy < N; ++iy)
        for (int64_t ix = 0; ix < N; ++ix) {
            int c = 0;
            for (int q = 0; q < 7; ++q) {
                const int64_t jx = ix + dx[q], jy = iy + dy[q];
                if (jx >= 0 && jx < N && jy >= 0 && jy < N) ++c;
            }
            cnt[v2d[iy * N + ix] + 1] = c;
        }
    rowptr[0] = 0;
    for (int64_t i = 0; i < nodes; ++i) rowptr[i + 1] = rowptr[i] + cnt[i + 1];
    free(cnt);
#pragma omp parallel for schedule(static)
    for (int64_t iy = 0; iy < N; ++iy)
        for (int64_t ix = 0; ix < N; ++ix) {
            int32_t col[7];
            int c = 0;
            for (int q = 0; q < 7; ++q) {
                const int64_t jx = ix + dx[q], jy = iy + dy[q];
                if (jx >= 0 && jx < N && jy >= 0 && jy < N) col[c++] = v2d[jy * N + jx];
            }
            for (int a = 1; a < c; ++a) {                              /* insertion sort, <= 7 entries */
                const int32_t v = col[a];
                int b = a - 1;
                while (b >= 0 && col[b] > v) { col[b + 1] = col[b]; --b; }
                col[b + 1] = v;
            }
            memcpy(colidx + rowptr[v2d[iy * N + ix]], col, sizeof(int32_t) * (size_t)c);
        }
    free(start);
}

/* transposed-entry positions and diagonal positions of a structurally symmetric pattern */
int fctc_tpos(int32_t n, const int32_t* rowptr, const int32_t* colidx, int32_t* tpos, int32_t* diagpos) {
    int bad = 0;
#pragma omp parallel for schedule(static) reduction(| : bad)
    for (int32_t i = 0; i < n; ++i) {
        int dg = -1;
        for (int32_t k = rowptr[i]; k < rowptr[i + 1]; ++k) {
            const int32_t j = colidx[k];
            if (j == i) dg = k;
            int32_t p = -1;
            for (int32_t q = rowptr[j]; q < rowptr[j + 1]; ++q)
                if (colidx[q] == i) { p = q; break; }
            if (p < 0) bad |= 1;
            tpos[k] = p;
        }
        if (dg < 0) bad |= 2;
        diagpos[i] = dg;
    }
    return bad;
}

/* vertex -> incident cells (ascending cell index), storing the two other vertices in the cell's cyclic order */
void fctc_incidence(int32_t nodes, int64_t ncells, const int32_t* cells, int32_t* ptr, int32_t* idx) {
    memset(ptr, 0, sizeof(int32_t) * ((size_t)nodes + 1));
    for (int64_t i = 0; i < 3 * ncells; ++i) ptr[cells[i] + 1]++;
    for (int32_t i = 0; i < nodes; ++i) ptr[i + 1] += ptr[i];
    int32_t* fill = (int32_t*)malloc(sizeof(int32_t) * (size_t)nodes);
    memcpy(fill, ptr, sizeof(int32_t) * (size_t)nodes);
    for (int64_t c = 0; c < ncells; ++c)
        for (int q = 0; q < 3; ++q) {
            const int32_t s = fill[cells[3 * c + q]]++;
            idx[2 * (int64_t)s] = cells[3 * c + (q + 1) % 3];
            idx[2 * (int64_t)s + 1] = cells[3 * c + (q + 2) % 3];
        }
    free(fill);
}

/* ---- P1 assembly as a row gather (cell order per row = dolfin's cell loop order) ------------------------------ */
typedef struct { double gx[3], gy[3], area; } geom_t;

static inline geom_t cell_geom(const double* xy, int32_t r, int32_t j1, int32_t j2) {
    geom_t g;
    const double x0 = xy[2 * (int64_t)r], y0 = xy[2 * (int64_t)r + 1];
    const double x1 = xy[2 * (int64_t)j1], y1 = xy[2 * (int64_t)j1 + 1];
    const double x2 = xy[2 * (int64_t)j2], y2 = xy[2 * (int64_t)j2 + 1];
    const double det = (x1 - x0) * (y2 - y0) - (x2 - x0) * (y1 - y0);
    const double inv = 1.0 / det;
    g.gx[1] = (y2 - y0) * inv;  g.gy[1] = -(x2 - x0) * inv;
    g.gx[2] = -(y1 - y0) * inv; g.gy[2] = (x1 - x0) * inv;
    g.gx[0] = -(g.gx[1] + g.gx[2]); g.gy[0] = -(g.gy[1] + g.gy[2]);
    g.area = 0.5 * fabs(det);
    return g;
}

static inline void row_add(const int32_t* colidx, int32_t k0, int32_t k1, double* vals, int32_t col, double v) {
    for (int32_t k = k0; k < k1; ++k)
        if (colidx[k] == col) { vals[k] += v; return; }
}

/* kind 0: mass matrix u v dx;  kind 1: scale * [(b.grad c) u v + (b.grad v) c u]  (row index = test function v) */
void fctc_assemble(int32_t kind, int32_t nodes, const int32_t* rowptr, const int32_t* colidx, const int32_t* inc_ptr,
                   const int32_t* inc_idx, const double* xy, const double* c, double bx, double by, double scale,
                   double* out) {
#pragma omp parallel for schedule(static)
    for (int32_t r = 0; r < nodes; ++r) {
        const int32_t k0 = rowptr[r], k1 = rowptr[r + 1];
        for (int32_t k = k0; k < k1; ++k) out[k] = 0.0;
        for (int32_t ci = inc_ptr[r]; ci < inc_ptr[r + 1]; ++ci) {
            const int32_t j1 = inc_idx[2 * (int64_t)ci], j2 = inc_idx[2 * (int64_t)ci + 1];
            const geom_t g = cell_geom(xy, r, j1, j2);
            const double m = g.area / 12.0;
            double e[3];
            if (kind == 0) {
                e[0] = 2.0 * m; e[1] = m; e[2] = m;
            } else {
                const double c0 = c[r], c1 = c[j1], c2 = c[j2];
                const double gcx = c0 * g.gx[0] + c1 * g.gx[1] + c2 * g.gx[2];
                const double gcy = c0 * g.gy[0] + c1 * g.gy[1] + c2 * g.gy[2];
                const double s = bx * gcx + by * gcy;                 /* b . grad c */
                const double bg = bx * g.gx[0] + by * g.gy[0];        /* b . grad phi_r */
                const double csum = (c0 + c1) + c2;
                e[0] = (s * m) * 2.0 + bg * (m * (csum + c0));
                e[1] = (s * m) + bg * (m * (csum + c1));
                e[2] = (s * m) + bg * (m * (csum + c2));
            }
            row_add(colidx, k0, k1, out, r, e[0]);
            row_add(colidx, k0, k1, out, j1, e[1]);
            row_add(colidx, k0, k1, out, j2, e[2]);
        }
        if (scale != 1.0)
            for (int32_t k = k0; k < k1; ++k) out[k] *= scale;
    }
}

void fctc_row_lump(int32_t n, const int32_t* rowptr, const int32_t* diagpos, const double* M, double* ML, double* Md) {
#pragma omp parallel for schedule(static)
    for (int32_t i = 0; i < n; ++i) {
        double s = 0.0;
        for (int32_t k = rowptr[i]; k < rowptr[i + 1]; ++k) s += M[k];
        ML[i] = s;
        Md[i] = M[diagpos[i]];
    }
}

/* ---- one FCT step (helpers.py:1715-1872), FCT_alg_ref sign convention -------------------------------------------- */
/* work: L[nnz], D[nnz]; vectors b, ulow, tmp, g, y0, y1, y2 (udot ends up in the returned pointer), Rp, Rn: 10 x n */
int fctc_step(int32_t n, const int32_t* rowptr, const int32_t* colidx, const int32_t* tpos, const int32_t* diagpos,
              const double* A, const double* rhs, const double* un, double dt, const double* M, const double* ML,
              const double* Md, double* L, double* D, double* vec, double rtol, int32_t maxit, double* out) {
    double* b = vec;
    double* ulow = vec + (size_t)n;
    double* tmp = vec + 2 * (size_t)n;
    double* g = vec + 3 * (size_t)n;
    double* yb[3] = {vec + 4 * (size_t)n, vec + 5 * (size_t)n, vec + 6 * (size_t)n};
    double* Rp = vec + 7 * (size_t)n;
    double* Rn = vec + 8 * (size_t)n;
    double* dinv = vec + 9 * (size_t)n;
    /* 1-2. D = artificial_diffusion_mat(-A) (:1769), L = M_L + dt (A - D) (:1775), b = M_L u_n + dt rhs (:1780) */
#pragma omp parallel for schedule(static)
    for (int32_t i = 0; i < n; ++i) {
        double dsum = 0.0;
        const int32_t kd = diagpos[i];
        for (int32_t k = rowptr[i]; k < rowptr[i + 1]; ++k) {
            if (k == kd) continue;
            const double a = A[k], at = A[tpos[k]];
            const double d = fmax(0.0, fmax(a, at));
            dsum += d;
            D[k] = d;
            L[k] = dt * (a - d);
        }
        D[kd] = -dsum;
        const double l = ML[i] + dt * (A[kd] + dsum);
        L[kd] = l;
        dinv[i] = 1.0 / l;
        b[i] = ML[i] * un[i] + (rhs ? dt * rhs[i] : 0.0);
        ulow[i] = un[i];
    }
    /* Jacobi sweeps x <- x + D^-1 (b - L x) until ||dx||_inf <= rtol ||x||_inf (twin of the GPU solver) */
    int its = 0;
    double* x = ulow;
    double* xn = tmp;
    for (its = 1; its <= maxit; ++its) {
        double dmax = 0.0, xmax = 0.0;
#pragma omp parallel for schedule(static) reduction(max : dmax, xmax)
        for (int32_t i = 0; i < n; ++i) {
            double acc = 0.0;
            for (int32_t k = rowptr[i]; k < rowptr[i + 1]; ++k) acc += L[k] * x[colidx[k]];
            const double dx = dinv[i] * (b[i] - acc);
            const double v = x[i] + dx;
            xn[i] = v;
            dmax = fmax(dmax, fabs(dx));
            xmax = fmax(xmax, fabs(v));
        }
        double* t = x; x = xn; xn = t;
        if (dmax <= rtol * xmax) break;
    }
    if (x != ulow) memcpy(ulow, x, sizeof(double) * (size_t)n);
    /* 4. g = -A u_low + rhs (:1814); udot = ChebSI(g, M, diag M, 20, 0.5, 2) (:143-185) */
#pragma omp parallel for schedule(static)
    for (int32_t i = 0; i < n; ++i) {
        double acc = 0.0;
        for (int32_t k = rowptr[i]; k < rowptr[i + 1]; ++k) acc += A[k] * ulow[colidx[k]];
        g[i] = -acc + (rhs ? rhs[i] : 0.0);
    }
    const double lmin = 0.5, lmax = 2.0, rho = (lmax - lmin) / (lmax + lmin), dscale = (lmin + lmax) / 2;
    double omega = 0.0;
    double *ymid = NULL, *yold = NULL, *ynew = NULL;
    for (int k = 1; k <= 20; ++k) {
        omega = (k == 2) ? 1 / (1 - rho * rho / 2) : 1 / (1 - (omega * rho * rho) / 4);
        ynew = yb[k % 3];
#pragma omp parallel for schedule(static)
        for (int32_t i = 0; i < n; ++i) {
            double acc = 0.0;
            if (ymid)
                for (int32_t q = rowptr[i]; q < rowptr[i + 1]; ++q) acc += M[q] * ymid[colidx[q]];
            const double z = (g[i] - acc) / (dscale * Md[i]);
            const double ym = ymid ? ymid[i] : 0.0, yo = yold ? yold[i] : 0.0;
            ynew[i] = omega * (z + ym - yo) + yo;
        }
        yold = ymid;
        ymid = ynew;
    }
    const double* udot = ynew;
    /* 5-7. fluxes, P+-, Q+-, R+- (:1818-1851) */
#pragma omp parallel for schedule(static)
    for (int32_t i = 0; i < n; ++i) {
        double pp = 0.0, pn = 0.0, umax = ulow[i], umin = ulow[i];
        for (int32_t k = rowptr[i]; k < rowptr[i + 1]; ++k) {
            const int32_t j = colidx[k];
            if (j == i) continue;
            const double f = M[k] * (udot[i] - udot[j]) + D[k] * (ulow[i] - ulow[j]);
            pp += fmax(f, 0.0);
            pn += fmin(f, 0.0);
            umax = fmax(umax, ulow[j]);
            umin = fmin(umin, ulow[j]);
        }
        const double qp = umax - ulow[i], qn = umin - ulow[i];
        Rp[i] = (pp != 0.0) ? fmin(1.0, ML[i] * qp / (dt * pp)) : 1.0;
        Rn[i] = (pn != 0.0) ? fmin(1.0, ML[i] * qn / (dt * pn)) : 1.0;
    }
    /* 8-9. limited fluxes and the explicit correction (:1860-1870) */
#pragma omp parallel for schedule(static)
    for (int32_t i = 0; i < n; ++i) {
        double fbar = 0.0;
        for (int32_t k = rowptr[i]; k < rowptr[i + 1]; ++k) {
            const int32_t j = colidx[k];
            if (j == i) continue;
            const double f = M[k] * (udot[i] - udot[j]) + D[k] * (ulow[i] - ulow[j]);
            const double alpha = (f > 0.0) ? fmin(Rp[i], Rn[j]) : fmin(Rn[i], Rp[j]);
            fbar += alpha * f;
        }
        out[i] = ulow[i] + dt * fbar / ML[i];
    }
    return its;
}

/* (x - t)^T M (x - t), t may be NULL: one term of L2_norm_sq_Q / L2_norm_sq_Omega (helpers.py:330-381), used by bench.py's
 * full-size parity check of the cost functional */
double fctc_norm_sq_M(int32_t n, const int32_t* rowptr, const int32_t* colidx, const double* M, const double* x,
                      const double* t) {
    double s = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : s)
    for (int32_t i = 0; i < n; ++i) {
        double acc = 0.0;
        for (int32_t k = rowptr[i]; k < rowptr[i + 1]; ++k) {
            const int32_t j = colidx[k];
            acc += M[k] * (t ? x[j] - t[j] : x[j]);
        }
        s += (t ? x[i] - t[i] : x[i]) * acc;
    }
    return s;
}
