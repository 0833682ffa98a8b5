// peer_reduce.cu -- the small all-reduces of an iteration over NVLink peer memory.
//
// A particle-sharded iteration sums six small vectors across the ranks:
// the barrier word behind the peer push, the pilot histogram, the sweep counters, the band
// counters, the histogram of the exact band keys (stein/utilities/compute_median.py:4-16 as a
// distributed radix select) and sum(phi^2) for the clip (abstract_stein_sampler.py:125).  At
// 8 B ... 128 KB they are pure latency.  When the engines of a node have exchanged the CUDA-IPC
// handles of their particle buffers (engine.cu), each buffer carries a MAILBOX behind the
// particles, and one kernel does the whole all-reduce:
//   1. every rank stores its vector into slot [rank] of every rank's mailbox (NVLink stores),
//   2. fences, then raises a flag (the epoch number) in every mailbox,
//   3. waits until all flags of its own mailbox show the epoch,
//   4. adds the slots in rank order -- the same order on every rank, so the f64 sum has the
//      same bits everywhere (the integer sums trivially).
// Two parities of slots alternate: a rank can start all-reduce k+1 (other parity) while a slow
// peer still reads k, but nobody can reach k+2 before every rank has raised its k+1 flags,
// i.e. has finished reading k.  The vector is cut into up to MB_BLOCKS slices, one CTA and one
// flag per slice, so the slices travel independently.
//
// A wait that lasts longer than MB_TIMEOUT_NS gives up and raises a host-mapped error word
// (reported by the next call): a dead peer costs an error, not a hung GPU.
#include "common.cuh"
#include "peer_reduce.cuh"

namespace stein {

constexpr unsigned long long MB_TIMEOUT_NS = 20ull * 1000ull * 1000ull * 1000ull;

struct MboxPtrs {
    unsigned long long *base[MB_RANKS];
};

struct PeerReduce {
    int rank = 0, world = 1;
    MboxPtrs mb{};
    unsigned long long epoch = 0;     // all-reduces issued so far (identical on every rank)
    int *h_error = nullptr;           // mapped pinned word, set by a kernel that timed out
    int *d_error = nullptr;
};

__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

template <typename T>
__global__ void __launch_bounds__(512)
mbox_allreduce_kernel(T *__restrict__ buf, int count, const MboxPtrs mb, int rank, int world,
                      unsigned long long epoch, int *__restrict__ err) {
    const int g = blockIdx.x, G = gridDim.x;
    const int i0 = (int)((long long)count * g / G), i1 = (int)((long long)count * (g + 1) / G);
    __shared__ int s_dead;          // this block gave up waiting (the host-mapped error word is only ever WRITTEN here:
    if (threadIdx.x == 0) s_dead = 0;   // reading it would cost every thread a PCIe round trip)
    const size_t pbase = (size_t)(epoch & 1ull) * MB_PARITY_WORDS;
    // 1. this rank's slice into slot [rank] of every mailbox (own included)
    for (int r = 0; r < world; ++r) {
        T *dst = reinterpret_cast<T *>(mb.base[r] + pbase + MB_FLAGS + (size_t)rank * MB_CAP);
        for (int i = i0 + (int)threadIdx.x; i < i1; i += (int)blockDim.x) dst[i] = buf[i];
    }
    __threadfence_system();
    __syncthreads();
    // 2. + 3. thread r raises this rank's flag at rank r, then waits for rank r's flag here
    if ((int)threadIdx.x < world) {
        __threadfence_system();
        st_release_sys(mb.base[threadIdx.x] + pbase + (size_t)rank * MB_BLOCKS + g, epoch);
        const unsigned long long *f = mb.base[rank] + pbase + (size_t)threadIdx.x * MB_BLOCKS + g;
        const unsigned long long t0 = global_timer_ns();
        while (ld_acquire_sys(f) < epoch) {
            if (global_timer_ns() - t0 > MB_TIMEOUT_NS) {
                *err = 1;
                s_dead = 1;
                break;
            }
        }
    }
    __syncthreads();
    // 4. the slots in rank order (L2 loads: the lines were written by other GPUs).  After a timeout
    // the slots are not trustworthy: the result is poisoned (all bits set: NaN / a count no bracket
    // check accepts) and the host reports the error word at its next look.
    const bool dead = s_dead != 0;
    const T *src = reinterpret_cast<const T *>(mb.base[rank] + pbase + MB_FLAGS);
    for (int i = i0 + (int)threadIdx.x; i < i1; i += (int)blockDim.x) {
        T s = __ldcg(src + i);
        for (int r = 1; r < world; ++r) s += __ldcg(src + (size_t)r * MB_CAP + i);
        if (dead) memset(&s, 0xff, sizeof(T));
        buf[i] = s;
    }
}

unsigned long long peer_reduce_epoch(const PeerReduce *pr) { return pr ? pr->epoch : 0ull; }

int peer_reduce_create(stein_ctx *ctx, int rank, int world, void *const *mailboxes, unsigned long long epoch0,
                       PeerReduce **out) {
    STEIN_REQUIRE(ctx, world >= 2 && world <= MB_RANKS && rank >= 0 && rank < world, "peer reduce: bad rank/world");
    PeerReduce *pr = new PeerReduce();
    pr->rank = rank;
    pr->world = world;
    pr->epoch = epoch0;
    for (int r = 0; r < world; ++r) pr->mb.base[r] = static_cast<unsigned long long *>(mailboxes[r]);
    cudaError_t err = cudaHostAlloc((void **)&pr->h_error, sizeof(int), cudaHostAllocMapped);
    if (err == cudaSuccess) {
        *pr->h_error = 0;
        err = cudaHostGetDevicePointer((void **)&pr->d_error, pr->h_error, 0);
    }
    if (err != cudaSuccess) {
        if (pr->h_error) cudaFreeHost(pr->h_error);
        delete pr;
        return fail(ctx, STEIN_ERR_CUDA, "peer reduce: %s", cudaGetErrorString(err));
    }
    *out = pr;
    return STEIN_OK;
}

void peer_reduce_destroy(PeerReduce *pr) {
    if (!pr) return;
    if (pr->h_error) cudaFreeHost(pr->h_error);
    delete pr;
}

static int peer_reduce_run(stein_ctx *ctx, void *buf, int64_t count, bool f64) {
    PeerReduce *pr = ctx->peer_reduce;
    if (*reinterpret_cast<volatile int *>(pr->h_error))
        return fail(ctx, STEIN_ERR_COMM, "peer-memory all-reduce timed out waiting for another rank");
    if (count <= 0) return STEIN_OK;
    pr->epoch += 1;
    const int blocks = (int)std::min<int64_t>(MB_BLOCKS, (count + 1023) / 1024);
    const int threads = count >= 512 ? 512 : std::max(32, (int)round_up(std::max<int64_t>(count, pr->world), 32));
    if (f64)
        mbox_allreduce_kernel<double><<<blocks, threads, 0, ctx->stream>>>(static_cast<double *>(buf), (int)count, pr->mb,
                                                                          pr->rank, pr->world, pr->epoch, pr->d_error);
    else
        mbox_allreduce_kernel<unsigned long long><<<blocks, threads, 0, ctx->stream>>>(
            static_cast<unsigned long long *>(buf), (int)count, pr->mb, pr->rank, pr->world, pr->epoch, pr->d_error);
    STEIN_CHECK_LAUNCH(ctx);
    return STEIN_OK;
}

// The all-reduces of the library: over peer memory when an engine has installed its mailboxes
// and the vector fits a slot, else through the hook of stein_comm.
int allreduce_u64(stein_ctx *ctx, void *buf_dev, int64_t count) {
    RegionTimer timer(ctx, STEIN_REGION_COLL);
    if (ctx->peer_reduce && count <= MB_CAP) {
        trace_mark(ctx, "(before all-reduce)");
        const int rc = peer_reduce_run(ctx, buf_dev, count, false);
        trace_mark(ctx, count == 1 ? "all-reduce:1 word" : (count > 8192 ? "all-reduce:band histogram" : "all-reduce:sweep counters"));
        return rc;
    }
    if (ctx->comm.allreduce_sum_u64(ctx->comm.user, buf_dev, count) != 0)
        return fail(ctx, STEIN_ERR_COMM, "allreduce_sum_u64 hook failed");
    return STEIN_OK;
}

int allreduce_f64(stein_ctx *ctx, void *buf_dev, int64_t count) {
    RegionTimer timer(ctx, STEIN_REGION_COLL);
    if (ctx->peer_reduce && count <= MB_CAP) {
        trace_mark(ctx, "(before all-reduce)");
        const int rc = peer_reduce_run(ctx, buf_dev, count, true);
        trace_mark(ctx, "all-reduce:sumsq");
        return rc;
    }
    if (ctx->comm.allreduce_sum_f64(ctx->comm.user, buf_dev, count) != 0)
        return fail(ctx, STEIN_ERR_COMM, "allreduce_sum_f64 hook failed");
    return STEIN_OK;
}

}  // namespace stein
