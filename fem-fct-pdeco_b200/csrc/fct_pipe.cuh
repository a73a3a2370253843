// TMA-fed shared-memory ring for the persistent row-block kernels (sm_90+/sm_100a).
//
// A CTA walks its row blocks (blk = blockIdx.x + i * gridDim.x).  For block i one elected thread issues 1-D bulk
// copies (cp.async.bulk global -> shared, completion counted in bytes on an mbarrier) of the block's contiguous
// CSR ranges -- NF fp64 value arrays and NI int32 index arrays -- into ring stage i % NST, up to NST-1 blocks
// ahead of the block being consumed.  The copy engine keeps tens of KB per SM in flight with no registers and no
// LSU instructions, which is what a latency-bound streaming kernel needs to approach HBM bandwidth; the 256
// threads of the CTA only do the thread-per-row arithmetic and the (L1/L2-served) neighbour gathers.
//
// Bulk copies need 16-byte aligned addresses and sizes: ranges start at ka = k0 & ~3 and are rounded up to a
// multiple of 4 elements.  A range that would run past the end of the arrays (only the last row block can) is
// clamped for the copy engine and its last < 4 elements are fetched by ordinary loads (pipe_tail).
#pragma once
#include "fct_common.cuh"

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0;
    const uint32_t addr = smem_u32(bar);
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
    }
}

// Programmatic dependent launch (PDL): a kernel launched with the programmatic-stream-serialization attribute may
// start while its predecessor is still draining; everything before pdl_wait() (barrier init, index arithmetic)
// overlaps the predecessor's tail, everything after it sees the predecessor's memory writes.  Both are no-ops for
// ordinary launches.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <int NF, int NI, int NST>
struct RowPipe {
    unsigned char* base;     // dynamic shared memory (16-byte aligned)
    uint64_t* bars;          // NST mbarriers (static shared)
    int cap;                 // elements per staged array
    int row_begin, row_end;
    int64_t nnz;
    const int32_t* rowptr;
    const double* gf[NF > 0 ? NF : 1];
    const int32_t* gi[NI > 0 ? NI : 1];

    __device__ __forceinline__ size_t stage_bytes() const { return (size_t)cap * (8 * NF + 4 * NI); }
    __device__ __forceinline__ double* f64(int stage, int j) const {
        return reinterpret_cast<double*>(base + stage * stage_bytes()) + (size_t)j * cap;
    }
    __device__ __forceinline__ int32_t* s32(int stage, int j) const {
        return reinterpret_cast<int32_t*>(base + stage * stage_bytes() + (size_t)NF * cap * 8) + (size_t)j * cap;
    }
    __device__ __forceinline__ int my_blocks() const {
        const int nblk = (row_end - row_begin + FCT_RB - 1) / FCT_RB;
        return ((int)blockIdx.x < nblk) ? (nblk - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    }
    __device__ __forceinline__ RowBlock block(int i) const {
        return row_block(rowptr, row_begin, row_end, (int)blockIdx.x + i * (int)gridDim.x);
    }
    // elements the copy engine fetches for block b (multiple of 4, clamped to the arrays)
    __device__ __forceinline__ int tma_count(const RowBlock& b) const {
        int cnt = ((b.k1 - b.ka) + 3) & ~3;
        const int64_t room = (nnz - (int64_t)b.ka) & ~(int64_t)3;
        if ((int64_t)cnt > room) cnt = (int)room;
        return cnt;
    }
    __device__ __forceinline__ void init() {
        pdl_trigger();
        if (threadIdx.x == 0) {
            for (int s = 0; s < NST; ++s) mbar_init(&bars[s], 1);
            mbar_fence_init();
        }
        __syncthreads();
        pdl_wait();            // nothing above touches global memory
    }
    // issuing thread only: CSR range of the next block to issue, loaded one iteration ahead so that the rowptr
    // round trip to DRAM is off the critical path
    int c_blk = -1, c_k0 = 0, c_k1 = 0;
    __device__ __forceinline__ void cache_range(int i, int nmine) {
        if (i < nmine) {
            const int blk = (int)blockIdx.x + i * (int)gridDim.x;
            const int r0 = row_begin + blk * FCT_RB;
            const int nr = min(FCT_RB, row_end - r0);
            c_k0 = rowptr[r0];
            c_k1 = rowptr[r0 + nr];
            c_blk = i;
        }
    }
    // one thread: start the copies of block i into stage i % NST
    __device__ __forceinline__ void issue(int i) {
        RowBlock b;
        if (c_blk == i) { b.k0 = c_k0; b.k1 = c_k1; b.ka = c_k0 & ~(FCT_ALIGN - 1); }
        else b = block(i);
        const int st = i % NST;
        const int cnt = tma_count(b);
        uint64_t* bar = &bars[st];
        mbar_expect_tx(bar, (uint32_t)cnt * (8 * NF + 4 * NI));
        if (cnt > 0) {
#pragma unroll
            for (int j = 0; j < NF; ++j) tma_load_1d(f64(st, j), gf[j] + b.ka, (uint32_t)cnt * 8, bar);
#pragma unroll
            for (int j = 0; j < NI; ++j) tma_load_1d(s32(st, j), gi[j] + b.ka, (uint32_t)cnt * 4, bar);
        }
    }
    __device__ __forceinline__ void prologue(int nmine) {
        if (threadIdx.x == 0) {
            for (int i = 0; i < NST - 1 && i < nmine; ++i) issue(i);
            cache_range(NST - 1, nmine);
        }
    }
    // top of iteration i: refill the stage that iteration i-1 released
    __device__ __forceinline__ void prefetch(int i, int nmine) {
        if (threadIdx.x == 0 && i + NST - 1 < nmine) {
            issue(i + NST - 1);
            cache_range(i + NST, nmine);
        }
    }
    // all threads: wait for block i's data; fetch the (rare) clamped tail with ordinary loads
    __device__ __forceinline__ void wait(int i, const RowBlock& b) const {
        const int st = i % NST;
        mbar_wait(&bars[st], (uint32_t)((i / NST) & 1));
        const int cnt = tma_count(b);
        const int need = b.k1 - b.ka;
        if (cnt < need) {      // block-uniform
            for (int e = cnt + (int)threadIdx.x; e < need; e += FCT_RB) {
#pragma unroll
                for (int j = 0; j < NF; ++j) f64(st, j)[e] = gf[j][(int64_t)b.ka + e];
#pragma unroll
                for (int j = 0; j < NI; ++j) s32(st, j)[e] = gi[j][(int64_t)b.ka + e];
            }
            __syncthreads();
        }
    }
};
