/*
 * dnppo.h -- C ABI of the PPO minibatch update of libdronenav.so (B200, sm_100a).
 *
 * Replaces, for the policy the reference trains (Sol/Model/PBDroneSimulator.py:251-258: separate pi / vf MLPs
 * [512, 512, 256] with Tanh, SB3 ActorCriticPolicy, diagonal Gaussian head), the body of the minibatch loop of
 *   PPO.train                               Sol/Model/Algorithms/sb3_ppo.py:190-316
 * i.e. evaluate_actions (:222), advantage normalisation (:233-234), clipped surrogate (:237-243), clipped value
 * loss (:251-262), entropy bonus (:265-271), approx-KL (:279-282), backward (:291), clip_grad_norm_ (:293) and the Adam step
 * (:294), as hand-written kernels: TMA + tcgen05.mma (BF16 planes, FP32 accumulation in TMEM) for the three
 * contractions of every layer, CUDA-core kernels for the heads / losses / reductions / Adam.
 *
 * Conventions are those of dronenav.h: 0 / negative DN_E* codes, dn_last_error(), caller-owned DEVICE buffers,
 * everything enqueued on the caller's stream, CUDA-graph capturable, no host synchronisation, no CPU fallback.
 */
#ifndef DNPPO_H_
#define DNPPO_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* arithmetic of the contractions */
#define DN_MLP_BF16X3  0   /* FP32-faithful: x = hi + lo BF16 planes, A B ~= Ah Bh + Ah Bl + Al Bh (default) */
#define DN_MLP_BF16    1   /* plain BF16 inputs, FP32 accumulation (labelled option) */

/* Numerics-test hook: one contraction of the update on caller-owned planes.  A "planes" buffer holds the hi plane
 * followed by the lo plane (BF16, row-major); with passes = 1 only the hi plane is read / written.
 *   kind 0 (forward)   out[M,N] planes = act(A[M,K] B[N,K]^T + bias[N])          act: 1 tanh, 0 identity
 *   kind 1 (dgrad)     out[M,N] planes = (A[M,K] B[K,N]) * (1 - H[M,N]^2)
 *   kind 2 (wgrad)     partial[slices][M][N] f32 = A[rows_s, M]^T B[rows_s, N]    rows = K, split into `slices`
 * M, N, K: multiples of 128 / 64 / 64 (kind 2: K a multiple of 64 * slices). */
int dn_mlp_gemm(int kind, int passes, int M, int N, int K, int slices, const void* a_planes, const void* b_planes,
                const float* bias, int act, const void* h_planes, void* out_planes, float* partial, void* stream);

#ifdef __cplusplus
}
#endif
#endif
