// Context, memory and error plumbing of libfctpdeco; closed-form structured-mesh bookkeeping.
#include "fct_common.cuh"
#include <cuda_profiler_api.h>
#include "../../include/fctpdeco.h"

#include <stdarg.h>
#include <stdlib.h>
#include <vector>

static thread_local char g_err[1024] = "";

void fct_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char* fct_last_error(void) { return g_err; }
extern "C" int fct_version(void) { return 100; }
extern "C" int fct_device_count(void) {
    int c = 0;
    if (cudaGetDeviceCount(&c) != cudaSuccess) { cudaGetLastError(); return 0; }
    return c;
}

// ---- structured mesh, closed form --------------------------------------------
// dolfin RectangleMesh "right": vertex id = iy*(n+1)+ix; square (ix,iy) -> triangles (v0,v1,v3),(v0,v2,v3).
// CG1 vertex_to_dof_map: anti-diagonal numbering d = ix - iy + n, dof = start(d) + (d <= n ? ix : iy).
namespace {
struct DofMap {
    int64_t n;
    std::vector<int64_t> start;
    explicit DofMap(int64_t n_) : n(n_), start(2 * n_ + 2, 0) {
        for (int64_t d = 0; d <= 2 * n; ++d) start[d + 1] = start[d] + (d <= n ? d : 2 * n - d) + 1;
    }
    inline int64_t dof(int64_t ix, int64_t iy) const {
        const int64_t d = ix - iy + n;
        return start[d] + (d <= n ? ix : iy);
    }
};
}  // namespace

extern "C" int fct_mesh_rect_sizes(int32_t n, int64_t* nodes, int64_t* cells, int64_t* nnz) {
    FCT_CHECK(n >= 1, "fct_mesh_rect_sizes: n must be >= 1");
    const int64_t N = (int64_t)n + 1;
    const int64_t edges = 2 * N * (N - 1) + (N - 1) * (N - 1);
    if (nodes) *nodes = N * N;
    if (cells) *cells = 2 * (int64_t)n * n;
    if (nnz) *nnz = N * N + 2 * edges;
    return 0;
}

extern "C" int fct_mesh_rect_build(int32_t n, double a1, double a2, int32_t* v2d, int32_t* cell_dofs, double* dof_xy,
                                   int32_t* rowptr, int32_t* colidx) {
    FCT_CHECK(n >= 1, "fct_mesh_rect_build: n must be >= 1");
    int64_t nodes, cells, nnz;
    fct_mesh_rect_sizes(n, &nodes, &cells, &nnz);
    FCT_CHECK(nnz < 2147483647LL, "fct_mesh_rect_build: nnz does not fit int32");
    const int64_t N = (int64_t)n + 1;
    const DofMap dm(n);
    const double h = (a2 - a1) / n;
    if (v2d || dof_xy) {
        for (int64_t iy = 0; iy < N; ++iy)
            for (int64_t ix = 0; ix < N; ++ix) {
                const int64_t d = dm.dof(ix, iy);
                if (v2d) v2d[iy * N + ix] = (int32_t)d;
                if (dof_xy) {
                    dof_xy[2 * d] = a1 + h * (double)ix;
                    dof_xy[2 * d + 1] = a1 + h * (double)iy;
                }
            }
    }
    if (cell_dofs) {
        for (int64_t iy = 0; iy < n; ++iy)
            for (int64_t ix = 0; ix < n; ++ix) {
                const int32_t d0 = (int32_t)dm.dof(ix, iy), d1 = (int32_t)dm.dof(ix + 1, iy);
                const int32_t d2 = (int32_t)dm.dof(ix, iy + 1), d3 = (int32_t)dm.dof(ix + 1, iy + 1);
                int32_t* c = cell_dofs + 6 * (iy * n + ix);
                c[0] = d0; c[1] = d1; c[2] = d3;
                c[3] = d0; c[4] = d2; c[5] = d3;
            }
    }
    if (rowptr && colidx) {
        // neighbours of (ix,iy) in ascending DoF order: diagonal d-1: (ix-1,iy),(ix,iy+1); diagonal d:
        // (ix-1,iy-1), self, (ix+1,iy+1); diagonal d+1: (ix,iy-1),(ix+1,iy).
        static const int ox[7] = {-1, 0, -1, 0, 1, 0, 1};
        static const int oy[7] = {0, 1, -1, 0, 1, -1, 0};
        // rows are in DoF order: walk diagonals
        int64_t k = 0;
        for (int64_t d = 0; d <= 2 * n; ++d) {
            const int64_t len = (d <= n ? d : 2 * n - d) + 1;
            for (int64_t p = 0; p < len; ++p) {
                int64_t ix, iy;
                if (d <= n) { ix = p; iy = ix - (d - n); } else { iy = p; ix = iy + (d - n); }
                const int64_t row = dm.start[d] + p;
                rowptr[row] = (int32_t)k;
                for (int q = 0; q < 7; ++q) {
                    const int64_t jx = ix + ox[q], jy = iy + oy[q];
                    if (jx < 0 || jy < 0 || jx > n || jy > n) continue;
                    colidx[k++] = (int32_t)dm.dof(jx, jy);
                }
            }
        }
        rowptr[nodes] = (int32_t)k;
        FCT_CHECK(k == nnz, "fct_mesh_rect_build: internal nnz mismatch");
    } else {
        FCT_CHECK(!rowptr && !colidx, "fct_mesh_rect_build: rowptr and colidx must be given together");
    }
    return 0;
}

// ---- context ---------------------------------------------------------------------------------------
int fct_kernels_configure(fct_ctx* ctx);
int fct_build_tpos(fct_ctx* ctx);
int fct_assembly_configure(fct_ctx* ctx);
int fct_drivers_configure(fct_ctx* ctx);
void fct_comm_destroy(fct_ctx* ctx);
void fct_p2p_destroy(fct_ctx* ctx);
void fct_templates_free(fct_ctx* ctx);
int fct_templates_build(fct_ctx* ctx);
void fct_geom_templates_free(fct_ctx* ctx);
void fct_tiles_free(fct_ctx* ctx);
int fct_tiles_prepare(fct_ctx* ctx);

template <typename T>
static int dev_alloc(T** p, size_t count) {
    // a few elements of slack so that 16-byte staging loads of the last block stay inside the allocation
    FCT_CUDA(cudaMalloc((void**)p, sizeof(T) * (count + 8)));
    FCT_CUDA(cudaMemset(*p, 0, sizeof(T) * (count + 8)));
    return 0;
}

extern "C" int fct_ctx_create(fct_ctx** out, int device, int32_t n, const int32_t* rowptr, const int32_t* colidx,
                              int32_t row_begin, int32_t row_end) {
    FCT_CHECK(out && rowptr && colidx, "fct_ctx_create: null argument");
    FCT_CHECK(n >= 1 && 0 <= row_begin && row_begin <= row_end && row_end <= n, "fct_ctx_create: bad row range");
    int ndev = fct_device_count();
    FCT_CHECK(ndev > 0, "fct_ctx_create: no CUDA device available (this library has no CPU fallback)");
    FCT_CHECK(device >= 0 && device < ndev, "fct_ctx_create: device %d out of range (%d devices)", device, ndev);
    FCT_CUDA(cudaSetDevice(device));
    fct_ctx* c = new fct_ctx();
    c->device = device;
    c->n = n;
    c->nnz = rowptr[n];
    c->row_begin = row_begin;
    c->row_end = row_end;
    c->cur_rb = row_begin;
    c->cur_re = row_end;
    c->depth = 1;
    c->ring_lo[0] = row_begin; c->ring_hi[0] = row_end;
    for (int j = 1; j < 9; ++j) { c->ring_lo[j] = 0; c->ring_hi[j] = n; }
    // staging capacity: any window of FCT_RB consecutive rows (row-block launches start at arbitrary ring boundaries)
    int cap = 0, maxrow = 0;
    for (int r0 = 0; r0 < n; ++r0) {
        const int r1 = (r0 + FCT_RB < n) ? r0 + FCT_RB : n;
        const int cnt = rowptr[r1] - (rowptr[r0] & ~(FCT_ALIGN - 1));
        if (cnt > cap) cap = cnt;
    }
    for (int r = 0; r < n; ++r) {
        const int len = rowptr[r + 1] - rowptr[r];
        if (len > maxrow) maxrow = len;
        if (len < 1) { delete c; FCT_CHECK(false, "fct_ctx_create: row %d is empty", r); }
    }
    c->cap = ((cap + 3) & ~3) + 4;
    { const char* e = getenv("FCT_NO_GRAPH"); c->use_graph = !(e && atoi(e) == 1); }
    { const char* e = getenv("FCT_TILE_ADAPT"); c->tile_adapt = !(e && atoi(e) == 0); }
    { const char* e = getenv("FCT_PDL"); c->use_pdl = (e && atoi(e) == 1); }   // measured: no gain with persistent grids
    { const char* e = getenv("FCT_TILE_KJ"); if (e && atoi(e) >= 2 && atoi(e) <= 4) c->tile_kj = atoi(e); }
    { const char* e = getenv("FCT_TILE_GRID"); if (e && atoi(e) >= 1) c->tile_grid_cap = atoi(e); }
    { const char* e = getenv("FCT_TILE_KC"); if (e && (atoi(e) == 0 || (atoi(e) >= 2 && atoi(e) <= 5))) c->tile_kc = atoi(e); }
    {
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, device) == cudaSuccess && prop.multiProcessorCount > 0)
            c->grid_cap = prop.multiProcessorCount * 8;
    }
    c->max_row = maxrow;
    int rc = 0;
    rc |= dev_alloc(&c->rowptr, (size_t)n + 1);
    rc |= dev_alloc(&c->colidx, (size_t)c->nnz);
    rc |= dev_alloc(&c->tpos, (size_t)c->nnz);
    rc |= dev_alloc(&c->M, (size_t)c->nnz);
    rc |= dev_alloc(&c->K, (size_t)c->nnz);
    rc |= dev_alloc(&c->Lvals, (size_t)c->nnz);
    rc |= dev_alloc(&c->Dvals, (size_t)c->nnz);
    rc |= dev_alloc(&c->Avals, (size_t)c->nnz);
    rc |= dev_alloc(&c->Svals, (size_t)c->nnz);
    rc |= dev_alloc(&c->ML, (size_t)n);
    rc |= dev_alloc(&c->Mdiag, (size_t)n);
    for (int i = 0; i < 12; ++i) rc |= dev_alloc(&c->w[i], (size_t)n);
    rc |= dev_alloc(&c->red, 64);
    rc |= dev_alloc(&c->jstate, 32);
    if (rc) { fct_ctx_destroy(c); return 1; }
    if (cudaMallocHost((void**)&c->pinned, 64 * sizeof(double)) != cudaSuccess) {
        fct_set_error("fct_ctx_create: cudaMallocHost failed");
        fct_ctx_destroy(c);
        return 1;
    }
    if (cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking) != cudaSuccess) {
        fct_set_error("fct_ctx_create: cudaStreamCreate failed");
        fct_ctx_destroy(c);
        return 1;
    }
    if (cudaMemcpy(c->rowptr, rowptr, sizeof(int32_t) * ((size_t)n + 1), cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(c->colidx, colidx, sizeof(int32_t) * (size_t)c->nnz, cudaMemcpyHostToDevice) != cudaSuccess) {
        fct_set_error("fct_ctx_create: pattern upload failed");
        fct_ctx_destroy(c);
        return 1;
    }
    if (fct_kernels_configure(c) || fct_assembly_configure(c) || fct_drivers_configure(c) || fct_build_tpos(c)) {
        fct_ctx_destroy(c);
        return 1;
    }
    *out = c;
    return 0;
}

extern "C" int fct_ctx_destroy(fct_ctx* c) {
    if (!c) return 0;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    if (c->jgraph.exec) cudaGraphExecDestroy((cudaGraphExec_t)c->jgraph.exec);
    if (c->jgraph.graph) cudaGraphDestroy((cudaGraph_t)c->jgraph.graph);
    if (c->hs.ready) {
        for (int i = 0; i < 2; ++i) { cudaFree(c->hs.cbuf[i]); cudaEventDestroy(c->hs.c_ready[i]); cudaEventDestroy(c->hs.c_free[i]); }
        for (int i = 0; i < 3; ++i) { cudaFree(c->hs.ubuf[i]); cudaEventDestroy(c->hs.u_ready[i]); cudaEventDestroy(c->hs.u_free[i]); }
        cudaStreamDestroy(c->hs.d2h_stream);
    }
    fct_templates_free(c);
    fct_tiles_free(c);
    fct_geom_templates_free(c);
    fct_p2p_destroy(c);
    fct_comm_destroy(c);
    cudaFree(c->rowptr); cudaFree(c->colidx); cudaFree(c->tpos);
    cudaFree(c->cells); cudaFree(c->xy); cudaFree(c->v2c_ptr); cudaFree(c->v2c_idx);
    cudaFree(c->M); cudaFree(c->ML); cudaFree(c->Mdiag); cudaFree(c->K);
    cudaFree(c->Lvals); cudaFree(c->Dvals); cudaFree(c->Avals); cudaFree(c->Svals);
    for (int i = 0; i < 12; ++i) { cudaFree(c->w[i]); cudaFree(c->fb_w[i]); }
    for (int i = 0; i < 6; ++i) cudaFree(c->sys_m[i]);
    for (int i = 0; i < 4; ++i) cudaFree(c->sys_v[i]);
    cudaFree(c->sys_wind);
    cudaFree(c->red); cudaFree(c->jstate);
    if (c->pinned) cudaFreeHost(c->pinned);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    cudaGetLastError();
    delete c;
    return 0;
}

extern "C" int fct_ctx_set_stream(fct_ctx* ctx, void* s) {
    FCT_CHECK(ctx, "fct_ctx_set_stream: null context");
    ctx->stream = (cudaStream_t)s;
    return 0;
}

extern "C" int fct_ctx_sync(fct_ctx* ctx) {
    FCT_CHECK(ctx, "fct_ctx_sync: null context");
    FCT_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}

extern "C" int fct_ctx_sizes(fct_ctx* ctx, int32_t* n, int64_t* nnz, int32_t* rb, int32_t* re) {
    FCT_CHECK(ctx, "fct_ctx_sizes: null context");
    if (n) *n = ctx->n;
    if (nnz) *nnz = ctx->nnz;
    if (rb) *rb = ctx->row_begin;
    if (re) *re = ctx->row_end;
    return 0;
}

extern "C" int fct_ctx_pattern_dev(fct_ctx* ctx, const int32_t** rp, const int32_t** ci, const int32_t** tp) {
    FCT_CHECK(ctx, "fct_ctx_pattern_dev: null context");
    if (rp) *rp = ctx->rowptr;
    if (ci) *ci = ctx->colidx;
    if (tp) *tp = ctx->tpos;
    return 0;
}

int fct_row_lump_diag(fct_ctx* ctx, const double* mat, double* out, double* diag);

// jstate[10] = the sweep count before which the Jacobi stopping test is not attempted, learnt from the previous solve;
// a new operator family, solver option or halo depth starts without it
static int fct_reset_check_from(fct_ctx* ctx) {
    FCT_CUDA(cudaMemsetAsync(ctx->jstate + 10, 0, sizeof(unsigned long long), ctx->stream));
    FCT_CUDA(cudaMemsetAsync(ctx->jstate + 13, 0, 8 * sizeof(unsigned long long), ctx->stream));      // sweep schedule (fct_kernels.cu)
    return 0;
}
int fct_halo_exchange_if(fct_ctx* ctx, double* vec);

extern "C" int fct_ctx_set_mass(fct_ctx* ctx, const double* M_dev) {
    FCT_CHECK(ctx && M_dev, "fct_ctx_set_mass: null argument");
    if (M_dev != ctx->M)
        FCT_CUDA(cudaMemcpyAsync(ctx->M, M_dev, sizeof(double) * (size_t)ctx->nnz, cudaMemcpyDeviceToDevice, ctx->stream));
    if (fct_row_lump_diag(ctx, ctx->M, ctx->ML, ctx->Mdiag)) return 1;
    if (fct_halo_exchange_if(ctx, ctx->ML)) return 1;
    if (fct_halo_exchange_if(ctx, ctx->Mdiag)) return 1;
    ctx->mass_set = true;
    if (fct_reset_check_from(ctx)) return 1;
    return fct_templates_build(ctx);
}

extern "C" int fct_ctx_static_dev(fct_ctx* ctx, const double** M, const double** ML, const double** Md, const double** K) {
    FCT_CHECK(ctx, "fct_ctx_static_dev: null context");
    if (M) *M = ctx->M;
    if (ML) *ML = ctx->ML;
    if (Md) *Md = ctx->Mdiag;
    if (K) *K = ctx->K;
    return 0;
}

extern "C" int fct_ctx_set_solver(fct_ctx* ctx, double rtol, int32_t max_sweeps) {
    FCT_CHECK(ctx && rtol > 0 && max_sweeps >= 2, "fct_ctx_set_solver: bad argument");
    ctx->rtol = rtol;
    ctx->max_sweeps = max_sweeps;
    return fct_reset_check_from(ctx);
}

extern "C" int fct_malloc(fct_ctx* ctx, void** p, int64_t bytes) {
    FCT_CHECK(ctx && p && bytes >= 0, "fct_malloc: bad argument");
    FCT_CUDA(cudaSetDevice(ctx->device));
    FCT_CUDA(cudaMalloc(p, (size_t)bytes + 64));
    return 0;
}
extern "C" int fct_free(fct_ctx* ctx, void* p) {
    FCT_CHECK(ctx, "fct_free: null context");
    FCT_CUDA(cudaFree(p));
    return 0;
}
extern "C" int fct_h2d(fct_ctx* ctx, void* dst, const void* src, int64_t bytes) {
    FCT_CHECK(ctx && dst && src && bytes >= 0, "fct_h2d: bad argument");
    FCT_CUDA(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyHostToDevice, ctx->stream));
    return 0;
}
extern "C" int fct_d2h(fct_ctx* ctx, void* dst, const void* src, int64_t bytes) {
    FCT_CHECK(ctx && dst && src && bytes >= 0, "fct_d2h: bad argument");
    FCT_CUDA(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDeviceToHost, ctx->stream));
    FCT_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}

extern "C" int fct_launch_count(fct_ctx* ctx, int64_t* count) {
    FCT_CHECK(ctx && count, "fct_launch_count: null argument");
    *count = ctx->launches;
    return 0;
}

int fct_p2p_exchange_count(fct_ctx* ctx, int64_t* count);     // fct_p2p.cu

extern "C" int fct_exchange_count(fct_ctx* ctx, int64_t* count) {
    FCT_CHECK(ctx && count, "fct_exchange_count: null argument");
    *count = ctx->exchanges;
    if (ctx->p2p) return fct_p2p_exchange_count(ctx, count);
    return 0;
}

extern "C" int fct_geom_template_count(fct_ctx* ctx, int32_t* count) {
    FCT_CHECK(ctx && count, "fct_geom_template_count: null argument");
    *count = ctx->gt_count;
    return 0;
}

// cudaProfilerStart/Stop of the runtime this library is linked against (ncu --profile-from-start off)
extern "C" int fct_profiler_range(int32_t start) {
    if (start) cudaProfilerStart(); else cudaProfilerStop();
    return 0;
}

extern "C" int fct_ctx_set_lumped(fct_ctx* ctx, const double* ML_dev) {
    FCT_CHECK(ctx && ML_dev, "fct_ctx_set_lumped: null argument");
    FCT_CUDA(cudaMemcpyAsync(ctx->ML, ML_dev, sizeof(double) * (size_t)ctx->n, cudaMemcpyDeviceToDevice, ctx->stream));
    return 0;
}

extern "C" int fct_event_create(fct_ctx* ctx, void** ev) {
    FCT_CHECK(ctx && ev, "fct_event_create: null argument");
    cudaEvent_t e;
    FCT_CUDA(cudaEventCreate(&e));
    *ev = (void*)e;
    return 0;
}
extern "C" int fct_event_record(fct_ctx* ctx, void* ev) {
    FCT_CHECK(ctx && ev, "fct_event_record: null argument");
    FCT_CUDA(cudaEventRecord((cudaEvent_t)ev, ctx->stream));
    return 0;
}
extern "C" int fct_event_elapsed_ms(fct_ctx* ctx, void* e0, void* e1, float* ms) {
    FCT_CHECK(ctx && e0 && e1 && ms, "fct_event_elapsed_ms: null argument");
    FCT_CUDA(cudaEventSynchronize((cudaEvent_t)e1));
    FCT_CUDA(cudaEventElapsedTime(ms, (cudaEvent_t)e0, (cudaEvent_t)e1));
    return 0;
}
extern "C" int fct_event_destroy(fct_ctx* ctx, void* ev) {
    FCT_CHECK(ctx, "fct_event_destroy: null context");
    if (ev) FCT_CUDA(cudaEventDestroy((cudaEvent_t)ev));
    return 0;
}

extern "C" int fct_host_alloc(fct_ctx* ctx, void** p, int64_t bytes) {
    FCT_CHECK(ctx && p && bytes >= 0, "fct_host_alloc: bad argument");
    FCT_CUDA(cudaMallocHost(p, (size_t)bytes + 64));
    return 0;
}
extern "C" int fct_host_free(fct_ctx* ctx, void* p) {
    FCT_CHECK(ctx, "fct_host_free: null context");
    if (p) FCT_CUDA(cudaFreeHost(p));
    return 0;
}

extern "C" int fct_ctx_set_rings(fct_ctx* ctx, int32_t depth, const int32_t* ring_lo, const int32_t* ring_hi) {
    FCT_CHECK(ctx && ring_lo && ring_hi, "fct_ctx_set_rings: null argument");
    FCT_CHECK(depth >= 1 && depth <= 8, "fct_ctx_set_rings: depth must be in 1..8");
    FCT_CHECK(ring_lo[0] == ctx->row_begin && ring_hi[0] == ctx->row_end, "fct_ctx_set_rings: ring 0 must be the owned rows");
    FCT_CHECK(ring_lo[depth] == 0 && ring_hi[depth] == ctx->n, "fct_ctx_set_rings: ring `depth` must be all local rows");
    for (int j = 1; j <= depth; ++j)
        FCT_CHECK(ring_lo[j] <= ring_lo[j - 1] && ring_hi[j] >= ring_hi[j - 1], "fct_ctx_set_rings: rings must be nested");
    ctx->depth = depth;
    for (int j = 0; j <= depth; ++j) { ctx->ring_lo[j] = ring_lo[j]; ctx->ring_hi[j] = ring_hi[j]; }
    for (int j = depth + 1; j < 9; ++j) { ctx->ring_lo[j] = 0; ctx->ring_hi[j] = ctx->n; }
    if (fct_tiles_prepare(ctx)) return 1;      // the tile lists depend on the rings
    return fct_reset_check_from(ctx);
}

extern "C" int fct_template_count(fct_ctx* ctx, int32_t* count) {
    FCT_CHECK(ctx && count, "fct_template_count: null argument");
    *count = ctx->tpl_count;
    return 0;
}
