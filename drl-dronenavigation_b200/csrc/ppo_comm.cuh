// ppo_comm.cuh -- the gradient all-reduce of the data-parallel PPO update over NVLink peer memory (include/dnppo.h:
// dn_ppo_comm_create / dn_ppo_comm_connect / dn_ppo_allreduce).
//
// The reference's learner is single-process (sb3_ppo.py:288-294: backward, clip, optimizer.step); with one process per GPU the
// only exchange is the sum of the flat gradient bucket (n_params + 1 floats, the extra one is the KL early-stop vote) before the
// clip.  3.2 MB on 8 GPUs is latency-bound: ncclAllReduce costs ~45 us per optimiser step against a 0.6 ms step.  Here every rank
// owns one cudaMalloc'ed region that all peers map through CUDA IPC:
//
//     flags [2][MAX_RANKS] | data [n_pad] (this rank's gradient, published) | result [n_pad] (the sum, pushed by the chunk owners)
//
// and one all-reduce is ONE kernel of three phases (at most one CTA per SM, all resident), capturable in the minibatch CUDA graph:
//   publish : grads -> own data; the last CTA, after a system-scope fence, writes the call's epoch into flags[0][rank] of EVERY peer
//   reduce  : waits until flags[0][*] of its own region carry the epoch (every peer has published), sums ITS chunk of all peers'
//             data in rank order (so the result is bit-identical on every rank and run to run) and stores the sum into the result
//             buffer of every peer; last CTA: fence, flags[1][rank] of every peer
//   collect : waits for flags[1][*] (every chunk owner has pushed), result -> grads, advances the epoch
// Two-shot (reduce-scatter + all-gather): each GPU moves 2 x (W-1)/W x 3.2 MB over NVLink instead of (W-1) x 3.2 MB.
// Every wait is bounded and traps (the host sees an error; a rank that was never called cannot hang its peers' GPUs).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace dncomm {

constexpr int MAX_RANKS = 16;
constexpr int FLAG_BYTES = 1024;                 // [2][MAX_RANKS] uint32, padded
constexpr int AR_THREADS = 512;                  // per CTA: enough threads that the reduce phase is one or two NVLink round trips per thread
constexpr uint32_t SPIN_CAP = 1u << 27;          // polls of local memory (~0.5 us each, about a minute) before a wait gives up

struct CommArgs {
    float* grads;                                // the learner's flat bucket (n floats incl. the vote)
    long long n;                                 // elements of the bucket
    long long n_pad;                             // multiple of 4 * world
    int rank, world;
    uint32_t* epoch;                             // device word: epoch of the last completed all-reduce
    unsigned int* counter;                       // [3] device words: CTAs that finished phase k of the running call (zero between calls)
    char* peer[MAX_RANKS];                       // base of every rank's region (own region included)
};

__device__ __forceinline__ uint32_t* flags_of(char* base, int phase) { return reinterpret_cast<uint32_t*>(base) + phase * MAX_RANKS; }
__device__ __forceinline__ float* data_of(char* base) { return reinterpret_cast<float*>(base + FLAG_BYTES); }
__device__ __forceinline__ float* result_of(char* base, long long n_pad) { return reinterpret_cast<float*>(base + FLAG_BYTES) + n_pad; }

__device__ __forceinline__ uint32_t ld_volatile_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// peer memory is read past every cache (another GPU rewrites it between two calls)
__device__ __forceinline__ float4 ld_volatile_f4(const float4* p) {
    float4 v;
    asm volatile("ld.volatile.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}

// all CTAs: wait until every rank's flag of `phase` in OUR region has reached `e`.  Only the polling threads fence (acquire: the
// peers' data stores happened before their flag stores); the CTA barrier orders everybody else's loads behind them.
__device__ __forceinline__ void wait_all(const CommArgs& a, int phase, uint32_t e) {
    if (threadIdx.x < a.world) {
        const uint32_t* f = flags_of(a.peer[a.rank], phase) + threadIdx.x;
        uint32_t spins = 0;
        while (static_cast<int32_t>(ld_volatile_u32(f) - e) < 0) {
            if (++spins > SPIN_CAP) __trap();
        }
        __threadfence_system();
    }
    __syncthreads();
}

// The last CTA to get here tells every peer that this rank has finished `phase` of epoch `e`.  One system-scope fence per CTA
// (thread 0, after the CTA barrier: the fence is cumulative over the stores the barrier ordered before it) -- a fence per thread
// made each of the three kernels cost ~15 us.
__device__ __forceinline__ void signal_all(const CommArgs& a, int phase, uint32_t e, int counter_idx) {
    __syncthreads();
    __shared__ int last;
    if (threadIdx.x == 0) {
        __threadfence_system();                  // release: this CTA's stores (peer stores included) before it is counted
        const unsigned int prev = atomicAdd(a.counter + counter_idx, 1u);
        last = (prev == gridDim.x - 1) ? 1 : 0;
    }
    __syncthreads();
    if (last && threadIdx.x < a.world) {
        __threadfence_system();                  // the other CTAs' counted stores before the flag
        *reinterpret_cast<volatile uint32_t*>(flags_of(a.peer[threadIdx.x], phase) + a.rank) = e;
    }
    __syncthreads();                             // `last` is reused by the next phase
}

// One kernel, three phases (see the header of this file).  All CTAs must be resident at once (a CTA that waits for the peers'
// flags keeps its SM slot while the rank's own flag needs every CTA of this grid to have published): the grid is at most one
// CTA per SM.  `counter[k]` counts the CTAs that finished phase k; the last CTA of the last phase clears all three.
__global__ void __launch_bounds__(AR_THREADS) allreduce_kernel(const CommArgs a) {
    const uint32_t e = *a.epoch + 1u;
    const long long n4 = a.n / 4, np4 = a.n_pad / 4;
    const long long t0 = blockIdx.x * static_cast<long long>(AR_THREADS) + threadIdx.x, stride = gridDim.x * static_cast<long long>(AR_THREADS);
    // ---- publish: grads -> own data
    {
        float4* dst = reinterpret_cast<float4*>(data_of(a.peer[a.rank]));
        const float4* src = reinterpret_cast<const float4*>(a.grads);
        for (long long i = t0; i < np4; i += stride) {
            float4 v;
            if (i < n4) v = src[i];
            else {
                const long long b = 4 * i;
                v.x = (b < a.n) ? a.grads[b] : 0.0f; v.y = (b + 1 < a.n) ? a.grads[b + 1] : 0.0f;
                v.z = (b + 2 < a.n) ? a.grads[b + 2] : 0.0f; v.w = (b + 3 < a.n) ? a.grads[b + 3] : 0.0f;
            }
            dst[i] = v;
        }
    }
    signal_all(a, 0, e, 0);
    // ---- reduce-scatter + push: this rank's chunk of every peer's data, summed in rank order, into every peer's result
    wait_all(a, 0, e);
    {
        const long long chunk4 = np4 / a.world;
        const long long c0 = chunk4 * a.rank;
        for (long long i = t0; i < chunk4; i += stride) {
            float4 v[MAX_RANKS];
#pragma unroll
            for (int p = 0; p < MAX_RANKS; ++p)                           // all peers' loads in flight together (one NVLink round trip)
                if (p < a.world) v[p] = ld_volatile_f4(reinterpret_cast<const float4*>(data_of(a.peer[p])) + c0 + i);
            float4 s = v[0];
#pragma unroll
            for (int p = 1; p < MAX_RANKS; ++p)                           // rank order: the same sum on every rank, every run
                if (p < a.world) { s.x += v[p].x; s.y += v[p].y; s.z += v[p].z; s.w += v[p].w; }
#pragma unroll
            for (int p = 0; p < MAX_RANKS; ++p)
                if (p < a.world) reinterpret_cast<float4*>(result_of(a.peer[p], a.n_pad))[c0 + i] = s;
        }
    }
    signal_all(a, 1, e, 1);
    // ---- collect: result -> grads
    wait_all(a, 1, e);
    {
        const float4* src = reinterpret_cast<const float4*>(result_of(a.peer[a.rank], a.n_pad));
        float4* dst = reinterpret_cast<float4*>(a.grads);
        for (long long i = t0; i < np4; i += stride) {
            const float4 v = __ldcg(src + i);                             // local memory, written by the peers: from L2, never from L1
            if (i < n4) dst[i] = v;
            else {
                const long long b = 4 * i;
                if (b < a.n) a.grads[b] = v.x;
                if (b + 1 < a.n) a.grads[b + 1] = v.y;
                if (b + 2 < a.n) a.grads[b + 2] = v.z;
                if (b + 3 < a.n) a.grads[b + 3] = v.w;
            }
        }
    }
    // the epoch advances and the counters are cleared when every CTA has finished (nobody reads *a.epoch after its first statement)
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int prev = atomicAdd(a.counter + 2, 1u);
        if (prev == gridDim.x - 1) { a.counter[0] = 0u; a.counter[1] = 0u; a.counter[2] = 0u; *a.epoch = e; }
    }
}

}  // namespace dncomm
