// ppo_update.cu -- the PPO minibatch update (include/dnppo.h) on sm_100a: tensor-map construction, launchers of the
// tcgen05 contraction kernel (dn_umma.cuh) and the CUDA-core kernels around it.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/dnppo.h"
#include "../../include/dronenav.h"
#include "dn_umma.cuh"

int dn_internal_fail(int code, const std::string& msg);   // dronenav.cu: sets the thread-local dn_last_error message

#define PPO_CUDA(expr)                                                                              \
    do {                                                                                            \
        cudaError_t _e = (expr);                                                                    \
        if (_e != cudaSuccess)                                                                      \
            return dn_internal_fail(DN_ECUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));  \
    } while (0)

namespace {

using namespace dnmma;

// ---- cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time dependency on libcuda) ----------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// BF16 row-major [rows, cols] tensor, box = box_rows x 64 columns (one 128-byte swizzle row per box row)
int make_map(CUtensorMap* m, const void* base, long long rows, long long cols, int box_rows) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return dn_internal_fail(DN_ECUDA, "cuTensorMapEncodeTiled is not available from this driver");
    cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
    cuuint64_t strides[1] = {static_cast<cuuint64_t>(cols) * 2};
    cuuint32_t box[2] = {64, static_cast<cuuint32_t>(box_rows)};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return dn_internal_fail(DN_ECUDA, "cuTensorMapEncodeTiled failed with CUresult " + std::to_string(static_cast<int>(r)));
    return DN_OK;
}

int pick_bn(int n) { return (n % 256 == 0) ? 256 : (n % 128 == 0) ? 128 : 64; }

struct GemmPlan {       // one contraction, ready to launch
    int kind = 0, bn = 0, grid = 0;
    CUtensorMap ma, mb;
    GemmArgs args;
};

int g_num_sms = 0;
int num_sms() {
    if (!g_num_sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
        if (g_num_sms <= 0) g_num_sms = 148;
    }
    return g_num_sms;
}

template <int KIND, int BN>
int launch_one(const GemmPlan& p, cudaStream_t st) {
    auto kern = umma_gemm<KIND, BN>;
    static bool attr_set = false;
    if (!attr_set) {
        PPO_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(Cfg<BN>::SMEM_BYTES)));
        attr_set = true;
    }
    kern<<<p.grid, NUM_THREADS, Cfg<BN>::SMEM_BYTES, st>>>(p.ma, p.mb, p.args);
    PPO_CUDA(cudaGetLastError());
    return DN_OK;
}

int launch_gemm(const GemmPlan& p, cudaStream_t st) {
#define DN_CASE(K, B) \
    if (p.kind == K && p.bn == B) return launch_one<K, B>(p, st);
    DN_CASE(K_FWD, 256) DN_CASE(K_FWD, 128) DN_CASE(K_FWD, 64)
    DN_CASE(K_DGRAD, 256) DN_CASE(K_DGRAD, 128) DN_CASE(K_DGRAD, 64)
    DN_CASE(K_WGRAD, 256) DN_CASE(K_WGRAD, 128) DN_CASE(K_WGRAD, 64)
#undef DN_CASE
    return dn_internal_fail(DN_EINVAL, "launch_gemm: no such kernel instantiation");
}

// Plans.  `*_planes` point at the hi plane; the lo plane starts `rows * cols` elements later.
// forward: out[M,N] = act(A[M,K] W[N,K]^T + bias)
int plan_fwd(GemmPlan* p, int passes, int M, int N, int K, const void* a, const void* w, const float* bias, int act, void* out) {
    if (M % BM || N % 64 || K % 64) return dn_internal_fail(DN_EINVAL, "mlp forward: M % 128, N % 64, K % 64 must be 0");
    p->kind = K_FWD;
    p->bn = pick_bn(N);
    int rc;
    if ((rc = make_map(&p->ma, a, 2LL * M, K, BM)) || (rc = make_map(&p->mb, w, 2LL * N, K, p->bn))) return rc;
    GemmArgs& g = p->args;
    memset(&g, 0, sizeof(g));
    g.m_tiles = M / BM; g.n_tiles = N / p->bn; g.slices = 1; g.k_blocks = K / BK; g.passes = passes;
    g.a_lo_row = M; g.b_lo_row = N; g.act = act; g.write_lo = passes > 1; g.bias = bias;
    g.out_hi = static_cast<__nv_bfloat16*>(out); g.out_lo = g.out_hi + static_cast<long long>(M) * N; g.ld_out = N;
    p->grid = std::min(g.m_tiles * g.n_tiles, num_sms());
    return DN_OK;
}
// dgrad: out[M,N] = (A[M,K] W[K,N]) * (1 - H[M,N]^2)
int plan_dgrad(GemmPlan* p, int passes, int M, int N, int K, const void* a, const void* w, const void* h, void* out) {
    if (M % BM || N % 64 || K % 64) return dn_internal_fail(DN_EINVAL, "mlp dgrad: M % 128, N % 64, K % 64 must be 0");
    p->kind = K_DGRAD;
    p->bn = pick_bn(N);
    int rc;
    if ((rc = make_map(&p->ma, a, 2LL * M, K, BM)) || (rc = make_map(&p->mb, w, 2LL * K, N, 64))) return rc;
    GemmArgs& g = p->args;
    memset(&g, 0, sizeof(g));
    g.m_tiles = M / BM; g.n_tiles = N / p->bn; g.slices = 1; g.k_blocks = K / BK; g.passes = passes;
    g.a_lo_row = M; g.b_lo_row = K; g.write_lo = passes > 1;
    g.out_hi = static_cast<__nv_bfloat16*>(out); g.out_lo = g.out_hi + static_cast<long long>(M) * N; g.ld_out = N;
    g.h_hi = static_cast<const __nv_bfloat16*>(h); g.h_lo = passes > 1 ? g.h_hi + static_cast<long long>(M) * N : nullptr;
    p->grid = std::min(g.m_tiles * g.n_tiles, num_sms());
    return DN_OK;
}
// wgrad: partial[s][Mo][No] = A[rows_s, Mo]^T B[rows_s, No], rows split into `slices`
int plan_wgrad(GemmPlan* p, int passes, int Mo, int No, int rows, int slices, const void* a, const void* b, float* partial) {
    if (Mo % BM || No % 64 || slices < 1 || rows % (64 * slices)) return dn_internal_fail(DN_EINVAL, "mlp wgrad: Mo % 128, No % 64, rows % (64 slices) must be 0");
    p->kind = K_WGRAD;
    p->bn = pick_bn(No);
    int rc;
    if ((rc = make_map(&p->ma, a, 2LL * rows, Mo, 64)) || (rc = make_map(&p->mb, b, 2LL * rows, No, 64))) return rc;
    GemmArgs& g = p->args;
    memset(&g, 0, sizeof(g));
    g.m_tiles = Mo / BM; g.n_tiles = No / p->bn; g.slices = slices; g.k_blocks = rows / slices / BK; g.passes = passes;
    g.a_lo_row = rows; g.b_lo_row = rows;
    g.partial = partial; g.ld_partial = No; g.slice_stride = static_cast<long long>(Mo) * No;
    p->grid = std::min(g.m_tiles * g.n_tiles * slices, num_sms());
    return DN_OK;
}

}  // namespace

extern "C" {

int dn_mlp_gemm(int kind, int passes, int M, int N, int K, int slices, const void* a, const void* b, const float* bias, int act,
                const void* h, void* out, float* partial, void* stream) {
    if (passes != 1 && passes != 3) return dn_internal_fail(DN_EINVAL, "dn_mlp_gemm: passes must be 1 or 3");
    if (!a || !b) return dn_internal_fail(DN_EINVAL, "dn_mlp_gemm: null operand");
    GemmPlan p;
    int rc;
    if (kind == 0) {
        if (!bias || !out) return dn_internal_fail(DN_EINVAL, "dn_mlp_gemm: forward needs bias and out");
        rc = plan_fwd(&p, passes, M, N, K, a, b, bias, act, out);
    } else if (kind == 1) {
        if (!h || !out) return dn_internal_fail(DN_EINVAL, "dn_mlp_gemm: dgrad needs h and out");
        rc = plan_dgrad(&p, passes, M, N, K, a, b, h, out);
    } else if (kind == 2) {
        if (!partial) return dn_internal_fail(DN_EINVAL, "dn_mlp_gemm: wgrad needs partial");
        rc = plan_wgrad(&p, passes, M, N, K, slices, a, b, partial);
    } else {
        return dn_internal_fail(DN_EINVAL, "dn_mlp_gemm: kind must be 0, 1 or 2");
    }
    if (rc) return rc;
    return launch_gemm(p, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
