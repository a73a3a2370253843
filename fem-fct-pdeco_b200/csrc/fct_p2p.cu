// Halo exchange and the Jacobi stopping test over NVLink peer memory (one process per GPU, CUDA IPC).
//
// NCCL send/recv costs ~30 us per (tiny) exchange and an FCT step needs ~45 of them, which caps strong scaling.
// Here every rank exports one small "mailbox" region; a neighbour's kernel writes halo data straight into it through
// a peer mapping (NVLink stores), publishes a sequence number with a system-scope release, and the owner's kernel
// acquires it and unpacks into the halo entries of the vector.  One kernel per exchange (push + wait + unpack), no
// host involvement, capturable in CUDA graphs (the sequence counter lives in device memory).  The Jacobi stopping
// test all-reduces two words the same way (every rank writes to every rank; world <= 8).
//
// Progress: a push never waits; a wait only depends on the neighbour's push of the same sequence number, which that
// rank issues before its own wait.  A rank can run at most one exchange ahead of a neighbour, so FCT_P2P_SLOTS = 4
// mailbox slots cannot be lapped.  Waits time out (~2 s) into an error flag instead of hanging the GPU.
#define FCT_GUARD_IMPL      // the IPC region must be a plain allocation (cudaIpcGetMemHandle wants its base address)
#include "fct_common.cuh"
#include "../../include/fctpdeco.h"

#define FCT_P2P_SLOTS 4
#define FCT_P2P_MAXWORLD 8
#define FCT_P2P_MAXVEC 2

struct P2PHeader {
    unsigned long long flag_lo[FCT_P2P_SLOTS];      // written by rank-1: sequence number of the data in mail_lo[slot]
    unsigned long long flag_hi[FCT_P2P_SLOTS];      // written by rank+1
    unsigned long long red_flag[FCT_P2P_MAXWORLD][FCT_P2P_SLOTS];
    unsigned long long red_val[FCT_P2P_MAXWORLD][FCT_P2P_SLOTS][2];
    unsigned long long xseq;                        // exchanges done so far (owner only)
    unsigned long long rseq;                        // reductions done so far (owner only)
    unsigned long long error;                       // set when a wait timed out
    unsigned long long done_cnt;                    // blocks of k_halo_xchg finished so far (owner only)
    unsigned long long push_cnt[2];                 // blocks that finished their push, per direction (owner only)
    unsigned long long pad[2];
};

struct fct_p2p {
    unsigned char* region = nullptr;                // my region: header + mailboxes
    size_t bytes = 0;
    int max_halo = 0;
    int rank = 0, world = 1;
    unsigned char* peer[FCT_P2P_MAXWORLD] = {nullptr};   // peer-mapped regions (peer[rank] = region)
};

__device__ __forceinline__ double* mail_ptr(unsigned char* region, int dir /*0 = from lo, 1 = from hi*/, int slot,
                                            int max_halo) {
    double* base = reinterpret_cast<double*>(region + sizeof(P2PHeader));
    return base + ((size_t)dir * FCT_P2P_SLOTS + slot) * (size_t)(FCT_P2P_MAXVEC * max_halo);
}

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
// spin until *p == want; false on timeout
__device__ __forceinline__ bool wait_flag(const unsigned long long* p, unsigned long long want) {
    const long long t0 = clock64();
    while (ld_acquire_sys(p) != want) {
        __nanosleep(64);
        if (clock64() - t0 > 4000000000LL) return false;
    }
    return true;
}

// Blocks [0, NB) exchange with rank-1, blocks [NB, 2 NB) with rank+1; nvec vectors travel in one message.  A message is
// split over the NB blocks of its direction (one block alone moves ~5 GB/s over NVLink, which made a 130 KB halo cost 26 us):
// every block pushes its slice, the last one to finish publishes the sequence number, every block waits for the incoming
// number and unpacks its slice.  All 2 NB blocks are co-resident (the launch follows the producing kernel in stream order), and
// a wait depends only on the neighbour's pushes, which never wait.
#define FCT_P2P_NB 16
#define FCT_P2P_NT 256
__global__ void __launch_bounds__(FCT_P2P_NT)
k_halo_xchg(unsigned char* mine, unsigned char* peer_lo, unsigned char* peer_hi, double* v0, double* v1, int nvec,
            int send_lo0, int send_lo1, int send_hi0, int send_hi1, int row_begin, int row_end, int n, int max_halo,
            const unsigned long long* cond) {
    // conditional exchange: *cond is a rank-uniform word (a decision taken on all-reduced values), so that either every rank
    // takes part or none does
    if (cond && *cond == 0ull) return;
    P2PHeader* H = reinterpret_cast<P2PHeader*>(mine);
    const unsigned long long seq = H->xseq + 1;          // same on every rank: all ranks run the same sequence
    const int slot = (int)(seq % FCT_P2P_SLOTS);
    const int dir = blockIdx.x / FCT_P2P_NB;              // 0: lo neighbour, 1: hi neighbour
    const int part = blockIdx.x % FCT_P2P_NB;
    unsigned char* peer = dir == 0 ? peer_lo : peer_hi;
    double* vecs[FCT_P2P_MAXVEC] = {v0, v1};
    if (peer) {
        // push: my boundary rows -> the neighbour's mailbox (I am its hi neighbour when it is my lo neighbour)
        const int s0 = dir == 0 ? send_lo0 : send_hi0, s1 = dir == 0 ? send_lo1 : send_hi1;
        double* dst = mail_ptr(peer, dir == 0 ? 1 : 0, slot, max_halo);
        for (int q = 0; q < nvec; ++q)
            for (int i = part * FCT_P2P_NT + threadIdx.x; i < s1 - s0; i += FCT_P2P_NB * FCT_P2P_NT)
                dst[(size_t)q * max_halo + i] = vecs[q][s0 + i];
        __syncthreads();
        __shared__ int ok;
        if (threadIdx.x == 0) {
            __threadfence_system();
            const unsigned long long arrived = atomicAdd(&H->push_cnt[dir], 1ull);
            if ((arrived + 1ull) % FCT_P2P_NB == 0ull) {
                __threadfence_system();                   // the other blocks' slices (fenced before their arrival) come first
                P2PHeader* PH = reinterpret_cast<P2PHeader*>(peer);
                st_release_sys(dir == 0 ? &PH->flag_hi[slot] : &PH->flag_lo[slot], seq);
            }
            // wait for the neighbour's message; once a wait has timed out the context is in error (the host fails the call at
            // its next synchronisation point, fct_p2p_check) and later exchanges do not spend another time-out each
            ok = (*reinterpret_cast<volatile unsigned long long*>(&H->error) == 0ull &&
                  wait_flag(dir == 0 ? &H->flag_lo[slot] : &H->flag_hi[slot], seq)) ? 1 : 0;
        }
        __syncthreads();
        if (ok) {
            const double* src = mail_ptr(mine, dir, slot, max_halo);
            const int h0 = dir == 0 ? 0 : row_end, h1 = dir == 0 ? row_begin : n;
            for (int q = 0; q < nvec; ++q)
                for (int i = part * FCT_P2P_NT + threadIdx.x; i < h1 - h0; i += FCT_P2P_NB * FCT_P2P_NT)
                    vecs[q][h0 + i] = __ldcg(src + (size_t)q * max_halo + i);
        } else if (threadIdx.x == 0) {
            H->error = 1ull;
        }
    }
    // the exchange counter advances once every block is done: the last one to finish bumps it
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned long long t = atomicAdd(&H->done_cnt, 1ull);
        if ((t + 1ull) % (2 * FCT_P2P_NB) == 0ull) H->xseq = seq;
    }
}

// all ranks: jstate[0..1] <- max over ranks (bit patterns of nonnegative doubles), then the Jacobi decision; `handle`
// is the graph WHILE condition when use_handle != 0
__global__ void __launch_bounds__(32)
k_p2p_max2_decide(unsigned char* mine, unsigned char* p0, unsigned char* p1, unsigned char* p2, unsigned char* p3,
                  unsigned char* p4, unsigned char* p5, unsigned char* p6, unsigned char* p7, int rank, int world,
                  unsigned long long* jstate, double rtol, unsigned long long max_sweeps, int use_handle,
                  cudaGraphConditionalHandle handle, unsigned long long which, int dry_force, unsigned long long kmax) {
    unsigned char* peers[FCT_P2P_MAXWORLD] = {p0, p1, p2, p3, p4, p5, p6, p7};
    P2PHeader* H = reinterpret_cast<P2PHeader*>(mine);
    __shared__ unsigned long long m0[32], m1[32];
    const unsigned long long seq = H->rseq + 1;
    const int slot = (int)(seq % FCT_P2P_SLOTS);
    const int t = threadIdx.x;
    // nothing to exchange when already converged, or before the sweep count at which the previous solve converged
    // (jstate[10], identical on every rank) -- the all-to-all is only paid for the last one or two tests of a solve
    const bool skip = jstate[3] != 0ull || (jstate[4] < jstate[10] && jstate[4] < max_sweeps);
    unsigned long long a = 0ull, b = 0ull;
    bool ok = *reinterpret_cast<volatile unsigned long long*>(&H->error) == 0ull;
    __shared__ int failed;
    if (t == 0) failed = 0;
    __syncthreads();
    if (!skip && t < world && ok) {
        if (t != rank) {
            P2PHeader* PH = reinterpret_cast<P2PHeader*>(peers[t]);
            PH->red_val[rank][slot][0] = jstate[0];
            PH->red_val[rank][slot][1] = jstate[1];
            __threadfence_system();
            st_release_sys(&PH->red_flag[rank][slot], seq);
            ok = wait_flag(&H->red_flag[t][slot], seq);
            a = __ldcg(&H->red_val[t][slot][0]);      // L1 may hold a stale copy of a slot written by a peer
            b = __ldcg(&H->red_val[t][slot][1]);
        } else {
            a = jstate[0];
            b = jstate[1];
        }
    }
    m0[t] = a; m1[t] = b;
    if (!ok) { failed = 1; H->error = 1ull; }
    __syncthreads();
    if (t == 0) {
        if (failed) {
            // a peer's words never arrived: do not decide on partial data -- leave the solve unconverged, end the graph
            // loop (every further test would time out again) and let the host fail the call (fct_p2p_check)
            if (use_handle) cudaGraphSetConditional(handle, 0u);
            return;
        }
        if (!skip) {
            for (int i = 1; i < world; ++i) { a = a > m0[i] ? a : m0[i]; b = b > m1[i] ? b : m1[i]; }
            const double delta = __longlong_as_double((long long)a);
            const double xm = __longlong_as_double((long long)b);
            jstate[5] = a;
            jstate[6] = b;
            if (delta <= rtol * xm || (dry_force && jstate[4] >= (unsigned long long)dry_force)) {
                jstate[3] = 1ull;
                jstate[12] = which;          // fused tile sweeps: 1 = the converged iterate is in the scratch vector
                const unsigned long long s = jstate[4], back = jstate[11] ? 2ull : 4ull;
                jstate[10] = s > back ? s - back : 0ull;
                if (kmax) tile_schedule_converged(jstate, rtol * xm, delta);
            } else {
                jstate[11] += 1ull;
                if (kmax) tile_schedule_failed(jstate, delta, kmax);
            }
            H->rseq = seq;
        } else if (kmax && jstate[3] == 0ull) {
            jstate[13] = tile_next_k(jstate[4], jstate[16], kmax);
        }
        if (jstate[3] == 0ull) { jstate[0] = 0ull; jstate[1] = 0ull; }    // restart the running maxima
        if (use_handle) cudaGraphSetConditional(handle, (jstate[3] != 0ull || jstate[4] >= max_sweeps) ? 0u : 1u);
    }
}

// ------------------------------------------------------------------------------------------------------
static size_t region_bytes(int max_halo) {
    return sizeof(P2PHeader) + (size_t)2 * FCT_P2P_SLOTS * FCT_P2P_MAXVEC * (size_t)max_halo * sizeof(double);
}

extern "C" int fct_p2p_create(fct_ctx* ctx, int32_t rank, int32_t world, int32_t max_halo, void* handle_out_64bytes) {
    FCT_CHECK(ctx && handle_out_64bytes, "fct_p2p_create: null argument");
    FCT_CHECK(world >= 1 && world <= FCT_P2P_MAXWORLD && rank >= 0 && rank < world, "fct_p2p_create: bad rank/world (max %d)",
              FCT_P2P_MAXWORLD);
    FCT_CHECK(!ctx->p2p, "fct_p2p_create: already created");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is expected to be 64 bytes");
    FCT_CUDA(cudaSetDevice(ctx->device));
    fct_p2p* p = new fct_p2p();
    p->rank = rank; p->world = world; p->max_halo = max_halo > 0 ? max_halo : 1;
    p->bytes = region_bytes(p->max_halo);
    if (cudaMalloc((void**)&p->region, p->bytes) != cudaSuccess) { delete p; fct_set_error("fct_p2p_create: cudaMalloc failed"); return 1; }
    cudaMemset(p->region, 0, p->bytes);
    cudaIpcMemHandle_t h;
    if (cudaIpcGetMemHandle(&h, p->region) != cudaSuccess) {
        fct_set_error("fct_p2p_create: cudaIpcGetMemHandle failed: %s", cudaGetErrorString(cudaGetLastError()));
        cudaFree(p->region); delete p; return 1;
    }
    memcpy(handle_out_64bytes, &h, sizeof(h));
    p->peer[rank] = p->region;
    ctx->p2p = p;
    return 0;
}

extern "C" int fct_p2p_connect(fct_ctx* ctx, const void* all_handles /* world x 64 bytes, rank order */) {
    FCT_CHECK(ctx && ctx->p2p && all_handles, "fct_p2p_connect: create the region first");
    fct_p2p* p = ctx->p2p;
    FCT_CUDA(cudaSetDevice(ctx->device));
    for (int r = 0; r < p->world; ++r) {
        if (r == p->rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, (const unsigned char*)all_handles + 64 * (size_t)r, sizeof(h));
        void* ptr = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            fct_set_error("fct_p2p_connect: cudaIpcOpenMemHandle(rank %d) failed: %s", r, cudaGetErrorString(e));
            cudaGetLastError();
            return 1;
        }
        p->peer[r] = (unsigned char*)ptr;
    }
    return 0;
}

void fct_p2p_destroy(fct_ctx* ctx) {
    fct_p2p* p = ctx->p2p;
    if (!p) return;
    for (int r = 0; r < p->world; ++r)
        if (r != p->rank && p->peer[r]) cudaIpcCloseMemHandle(p->peer[r]);
    cudaFree(p->region);
    delete p;
    ctx->p2p = nullptr;
}

// FCT_P2P_DRY (diagnostic, timing only -- results are WRONG): 1 = the exchange / stopping-test kernels are launched but neither
// push, wait nor unpack; 2 = the exchange kernels are not launched at all.  tools/mgpu_diag.py uses it to split the cost of
// a multi-GPU step into launch overhead and peer waiting.
static int p2p_dry() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("FCT_P2P_DRY"); v = e ? atoi(e) : 0; }
    return v;
}

bool fct_p2p_ready(const fct_ctx* ctx) {
    const fct_p2p* p = ctx->p2p;
    if (!p) return false;
    for (int r = 0; r < p->world; ++r)
        if (!p->peer[r]) return false;
    return true;
}

int fct_p2p_exchange(fct_ctx* ctx, double* v0, double* v1, const unsigned long long* cond) {
    fct_p2p* p = ctx->p2p;
    const int halo_lo = ctx->row_begin, halo_hi = ctx->n - ctx->row_end;
    const int slo = ctx->send_lo[1] - ctx->send_lo[0], shi = ctx->send_hi[1] - ctx->send_hi[0];
    FCT_CHECK(halo_lo <= p->max_halo && halo_hi <= p->max_halo && slo <= p->max_halo && shi <= p->max_halo,
              "fct_p2p_exchange: halo larger than the mailbox");
    unsigned char* lo = p->rank > 0 ? p->peer[p->rank - 1] : nullptr;
    unsigned char* hi = p->rank + 1 < p->world ? p->peer[p->rank + 1] : nullptr;
    if (p2p_dry() == 2) return 0;
    if (p2p_dry() == 1) lo = hi = nullptr;
    k_halo_xchg<<<2 * FCT_P2P_NB, FCT_P2P_NT, 0, ctx->stream>>>(p->region, lo, hi, v0, v1, v1 ? 2 : 1, ctx->send_lo[0], ctx->send_lo[1],
                                             ctx->send_hi[0], ctx->send_hi[1], ctx->row_begin, ctx->row_end, ctx->n,
                                             p->max_halo, cond);
    ctx->launches++;
    return 0;
}

int fct_p2p_max2_decide(fct_ctx* ctx, double rtol, int max_sweeps, int use_handle, cudaGraphConditionalHandle handle,
                        int which, int kmax) {
    fct_p2p* p = ctx->p2p;
    const bool dry = p2p_dry() != 0;
    k_p2p_max2_decide<<<1, 32, 0, ctx->stream>>>(p->region, p->peer[0], p->peer[1], p->peer[2], p->peer[3], p->peer[4],
                                                 p->peer[5], p->peer[6], p->peer[7], dry ? 0 : p->rank, dry ? 1 : p->world,
                                                 ctx->jstate, rtol,
                                                 (unsigned long long)max_sweeps, use_handle, handle,
                                                 (unsigned long long)which, dry ? 16 : 0, (unsigned long long)kmax);
    ctx->launches++;
    return 0;
}

// Called wherever the host already synchronises with the stream (fct_read_step_info, the sweep-count read-back of the
// time loops, fct_norm_sq_Q ...): a timed-out peer wait fails the call instead of leaving stale halo data behind.
int fct_p2p_check(fct_ctx* ctx, const char* what) {
    if (!ctx->p2p) return 0;
    unsigned long long e = 0;
    FCT_CUDA(cudaMemcpyAsync(&e, ctx->p2p->region + offsetof(P2PHeader, error), sizeof(e), cudaMemcpyDeviceToHost, ctx->stream));
    FCT_CUDA(cudaStreamSynchronize(ctx->stream));
    FCT_CHECK(e == 0, "%s: a peer-memory halo exchange / stopping test timed out (a neighbouring rank did not arrive within "
              "~2 s): results are invalid", what);
    return 0;
}

// halo exchanges executed so far on this context (device counter: exchanges inside CUDA-graph bodies are counted too)
int fct_p2p_exchange_count(fct_ctx* ctx, int64_t* count) {
    unsigned long long x = 0;
    FCT_CUDA(cudaMemcpyAsync(&x, ctx->p2p->region + offsetof(P2PHeader, xseq), sizeof(x), cudaMemcpyDeviceToHost, ctx->stream));
    FCT_CUDA(cudaStreamSynchronize(ctx->stream));
    *count = (int64_t)x;
    return 0;
}

extern "C" int fct_p2p_error(fct_ctx* ctx, int32_t* error_host) {
    FCT_CHECK(ctx && error_host, "fct_p2p_error: null argument");
    *error_host = 0;
    if (!ctx->p2p) return 0;
    unsigned long long e = 0;
    FCT_CUDA(cudaMemcpyAsync(&e, ctx->p2p->region + offsetof(P2PHeader, error), sizeof(e), cudaMemcpyDeviceToHost, ctx->stream));
    FCT_CUDA(cudaStreamSynchronize(ctx->stream));
    *error_host = (int32_t)e;
    return 0;
}
