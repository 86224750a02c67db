// dronenav.cu -- fused control-step kernel for sm_100a + the C ABI of include/dronenav.h.
//
// One launch of step_kernel = PBDroneEnv.step for every environment of the shard:
//   rescale_action -> _preprocessAction -> S x (_dynamics + _integrateQ) -> Euler ->
//   _computeObs -> _computeReward (waypoint state machine) -> _computeTerminated ->
//   _computeTruncated -> _update_state_post_step -> Monitor bookkeeping ->
//   SubprocVecEnv auto-reset [-> NormalizeObservation]
// (Sol/Model/Environments/PBDroneEnv.py:171-223,296-398,434-607,609-665,678-786,872-971;
//  Sol/PyBullet/BaseAviary.py:324-453,899-973; Sol/Model/Environments/normalize.py:10-97).
//
// Data movement per env-step: 1 x LDG.128 action, 7 x LDG.128 + 7 x STG.128 state planes,
// reward / done / found_targets scalars, and the [BLOCK, obs_dim] observation tile which
// is staged in shared memory and written with ONE 1-D TMA bulk store
// (cp.async.bulk.global.shared::cta) per CTA.  No tensor cores: the step is a streaming
// map, HBM-bound at S = 1 and issue/HBM balanced at S = 8.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
#include <cmath>
#include <string>
#include <vector>
#include <new>

#include "../../include/dronenav.h"
#include "dn_params.h"
#include "dn_device.cuh"

namespace dn {

constexpr int kBlock = 128;
constexpr int kMaxObs = 13;

// ---------------------------------------------------------------------------
// small PTX wrappers (TMA 1-D bulk store of the observation tile)
// ---------------------------------------------------------------------------
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void bulk_store_g2s_commit(void* gdst, const void* ssrc, uint32_t bytes) {
    const uint32_t saddr = static_cast<uint32_t>(__cvta_generic_to_shared(ssrc));
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 :: "l"(gdst), "r"(saddr), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_store_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

// ---------------------------------------------------------------------------
// normalize.RunningMeanStd with a batch of one (normalize.py:19-47) followed by
// NormalizeObservation.normalize (:94-97).  mean/var/count are per env, FP32 planes.
// ---------------------------------------------------------------------------
__device__ __forceinline__ float rms_update_normalize(float x, float& mean, float& var, float count) {
    const float tot = count + 1.0f;
    const float delta = x - mean;
    mean = mean + delta / tot;
    const float m2 = var * count + (delta * delta) * count / tot;
    var = m2 / tot;
    return (x - mean) / sqrtf(var + 1e-8f);
}

struct StepResult {
    float reward;
    uint8_t done;
    int found;
    bool finished;          // done (terminated or truncated)
    float ep_ret;
    int ep_len;
    bool success, crash;
};

// One control step for one environment.  `obs_row` receives the observation the VecEnv
// returns (the reset observation when the episode ended), `term_row` (may alias nothing)
// the terminal observation.  Returns bookkeeping for outputs and statistics.
template <int PHYS>
__device__ __forceinline__ StepResult env_step(const Params& P, EnvState& s, const float4 act,
                                               float& last_rpm_sum, float* obs_row, float* term_row) {
    StepResult out;
    const int T = P.num_targets;
    int idx = static_cast<int>(s.bits >> kIdxShift);
    int steps = static_cast<int>(s.bits & kStepsMask);
    bool just_found = (s.bits & kJustFoundBit) != 0;

    // state at step entry: PBDroneEnv.current_vel / current_ang_v and _current_position
    const float evx = s.vx, evy = s.vy, evz = s.vz;
    const float eax = s.ax, eay = s.ay, eaz = s.az;
    const float epx = s.px, epy = s.py, epz = s.pz;

    // ---- action -> rpm (PBDroneEnv.py:173-176,872-895) ----------------------
    float rpm[4];
    rpm[0] = action_to_rpm(P, act.x);
    if (P.act_type == 2) { rpm[1] = rpm[2] = rpm[3] = rpm[0]; }
    else { rpm[1] = action_to_rpm(P, act.y); rpm[2] = action_to_rpm(P, act.z); rpm[3] = action_to_rpm(P, act.w); }

    // ---- physics (BaseAviary.py:410-444) -------------------------------------
    integrate<PHYS>(P, s, rpm, last_rpm_sum);

    // ---- observation (PBDroneEnv.py:296-336): new pose, STALE distance -------
    kinematic_obs(P, s, term_row);
    if (P.obs_dim == 13) term_row[12] = s.dist / P.max_target_dist;
#pragma unroll
    for (int k = 0; k < kMaxObs; ++k) if (k < P.obs_dim) term_row[k] = clip_f32_range(term_row[k]);

    // ---- reward + waypoint state machine (PBDroneEnv.py:475-571) -------------
    float fx, fy, fz;
    forward_vector(s.qx, s.qy, s.qz, s.qw, fx, fy, fz);
    const RewardParams& W = P.rw;
    bool terminated;
    bool is_done = false;
    float reward;
    out.crash = false;
    if (collided(P, s.px, s.py, s.pz, idx)) {
        reward = W.crash;                      // -10.0, not divided (:489-490)
        terminated = true;
        out.crash = true;
    } else {
        if (s.dist <= P.threshold) {           // stale distance (:539)
            idx += 1;
            if (idx == T) {
                reward = W.final_bonus / W.divisor;
                is_done = true;
            } else {
                const float4 tg = __ldg(&P.targets[idx]);
                reward = (W.capture_bonus + W.capture_orient_w * orientation_term(fx, fy, fz, s.px, s.py, s.pz, tg)) / W.divisor;
                just_found = true;
            }
        } else {
            const float4 tg = __ldg(&P.targets[idx]);
            float r = W.exp_w * expf(-W.exp_k * s.dist);
            r += just_found ? 0.0f : (s.prev_dist - s.dist) * W.progress_w;
            r += W.orient_w * orientation_term(fx, fy, fz, s.px, s.py, s.pz, tg);
            if (W.smooth_w != 0.0f) {          // smoothness_reward (:599-607), one-step-stale velocities
                const float lx = evx - s.pvx, ly = evy - s.pvy, lz = evz - s.pvz;
                const float gx = eax - s.pax, gy = eay - s.pay, gz = eaz - s.paz;
                const float lin = sqrtf(lx * lx + ly * ly + lz * lz);
                const float ang = sqrtf(gx * gx + gy * gy + gz * gz);
                r += W.smooth_w * ((lin > W.smooth_lin_thr ? -lin : 0.0f) + (ang > W.smooth_ang_thr ? -ang : 0.0f));
            }
            reward = r / W.divisor;
            just_found = false;
        }
        s.prev_dist = s.dist;                  // :568
        // _computeTerminated after the reward (:448,:456-473): index possibly advanced
        terminated = is_done || collided(P, s.px, s.py, s.pz, idx);
    }
    const bool truncated = (P.max_steps <= steps);   // before this step's increment (:444-454)
    out.found = idx;                                 // :434-442

    // ---- _update_state_post_step (PBDroneEnv.py:196-223), skipped when terminated
    if (!terminated) {
        steps += 1;
        s.pvx = evx; s.pvy = evy; s.pvz = evz;
        s.pax = eax; s.pay = eay; s.paz = eaz;
        const float4 tg = __ldg(&P.targets[idx]);
        const float dx = tg.x - s.px, dy = tg.y - s.py, dz = tg.z - s.pz;
        s.dist = sqrtf(dx * dx + dy * dy + dz * dz);
    }

    // ---- Monitor (SB3) --------------------------------------------------------
    s.ep_ret += reward;
    s.ep_len += 1;
    out.reward = reward;
    out.done = static_cast<uint8_t>((terminated ? DN_DONE_TERMINATED : 0) | (truncated ? DN_DONE_TRUNCATED : 0));
    out.finished = terminated || truncated;
    out.ep_ret = s.ep_ret;
    out.ep_len = s.ep_len;
    out.success = is_done;

    if (!out.finished) {
#pragma unroll
        for (int k = 0; k < kMaxObs; ++k) if (k < P.obs_dim) obs_row[k] = term_row[k];
    } else {
        // ---- auto-reset: BaseAviary.reset (:276-320) then PBDroneEnv.reset (:609-665).
        // The reset observation is taken BEFORE the distances are reset (:318 vs :651), and the
        // new distance uses the stale _current_position: the position of the last non-terminal
        // post-step (entry position if this step terminated, the new position if it was only
        // truncated, unchanged if no post-step has run since the previous reset).
        const float stale_dist = s.dist;
        float D;
        if (steps == 0) {
            D = s.dist;
        } else {
            const float4 t0 = __ldg(&P.targets[0]);
            const float cx = terminated ? epx : s.px, cy = terminated ? epy : s.py, cz = terminated ? epz : s.pz;
            const float dx = cx - t0.x, dy = cy - t0.y, dz = cz - t0.z;
            D = sqrtf(dx * dx + dy * dy + dz * dz);
        }
        s.px = P.init_pos[0]; s.py = P.init_pos[1]; s.pz = P.init_pos[2];
        s.qx = P.init_quat[0]; s.qy = P.init_quat[1]; s.qz = P.init_quat[2]; s.qw = P.init_quat[3];
        s.vx = s.vy = s.vz = 0.0f; s.wx = s.wy = s.wz = 0.0f; s.ax = s.ay = s.az = 0.0f;
        s.pvx = s.pvy = s.pvz = 0.0f; s.pax = s.pay = s.paz = 0.0f;
        s.dist = D; s.prev_dist = D;
        idx = 0; steps = 0; just_found = false;
        s.ep_ret = 0.0f; s.ep_len = 0; s.ep_count += 1u;
        last_rpm_sum = 0.0f;                   // _housekeeping: last_clipped_action = 0 (BaseAviary.py:545)
#pragma unroll
        for (int k = 0; k < 12; ++k) obs_row[k] = P.init_obs[k];
        if (P.obs_dim == 13) obs_row[12] = clip_f32_range(stale_dist / P.max_target_dist);
    }
    s.bits = (static_cast<uint32_t>(idx) << kIdxShift) | (just_found ? kJustFoundBit : 0u) |
             (static_cast<uint32_t>(steps) & kStepsMask);
    return out;
}

// ---------------------------------------------------------------------------
// the fused step kernel
// ---------------------------------------------------------------------------
template <int PHYS, bool NORM>
__global__ void __launch_bounds__(kBlock)
step_kernel(const __grid_constant__ Params P, const __grid_constant__ StepIO io, int num_steps, int per_step) {
    __shared__ __align__(128) float tile[kBlock * kMaxObs];
    const int tid = threadIdx.x;
    const int base = blockIdx.x * kBlock;
    const int i = base + tid;
    const bool active = i < P.n;
    const int D = P.obs_dim;
    const int n_here = min(kBlock, P.n - base);

    EnvState s;
    float last_rpm_sum = 0.0f;
    if (active) {
        load_state(P, i, s);
        if (PHYS & 1) last_rpm_sum = P.last_rpm_sum[i];
    }

    for (int t = 0; t < num_steps; ++t) {
        const bool write_out = per_step || (t == num_steps - 1);
        const size_t orow = per_step ? static_cast<size_t>(t) * P.n : 0;   // output row offset (in envs)
        float* obs_row = tile + tid * D;
        StepResult r;
        r.finished = false; r.success = false; r.crash = false; r.done = 0; r.ep_ret = 0.f; r.ep_len = 0; r.found = 0; r.reward = 0.f;
        if (active) {
            const float4 act = __ldg(io.actions + static_cast<size_t>(t) * P.n + i);
            float term_row[kMaxObs];
            r = env_step<PHYS>(P, s, act, last_rpm_sum, obs_row, term_row);
            if (NORM) {
                // NormalizeObservation sits inside Monitor and the worker's auto-reset
                // (PBDroneSimulator.py:181): the terminal observation updates the running
                // statistics in .step, the reset observation again in .reset (normalize.py:74-92).
                const size_t N = P.n;
                float* cnt_p = P.obs_rms + static_cast<size_t>(2 * D) * N + i;
                float cnt = *cnt_p;
                for (int k = 0; k < D; ++k) {
                    float* mp = P.obs_rms + static_cast<size_t>(k) * N + i;
                    float* vp = P.obs_rms + static_cast<size_t>(D + k) * N + i;
                    float m = *mp, v = *vp;
                    const float tn = rms_update_normalize(term_row[k], m, v, cnt);
                    term_row[k] = tn;
                    if (r.finished) obs_row[k] = rms_update_normalize(obs_row[k], m, v, cnt + 1.0f);
                    else obs_row[k] = tn;
                    *mp = m; *vp = v;
                }
                *cnt_p = cnt + (r.finished ? 2.0f : 1.0f);
            }
            if (write_out) {
                io.reward[orow + i] = r.reward;
                io.done[orow + i] = r.done;
                if (io.found_targets) io.found_targets[orow + i] = r.found;
                if (r.finished) {
                    if (io.terminal_obs) {
                        float* to = io.terminal_obs + (orow + i) * D;
                        for (int k = 0; k < D; ++k) to[k] = term_row[k];
                    }
                    if (io.episode_return) io.episode_return[orow + i] = r.ep_ret;
                    if (io.episode_length) io.episode_length[orow + i] = r.ep_len;
                }
            }
        }
        // ---- Monitor statistics: warp-aggregated, one atomic per counter per warp ----
        const unsigned fin = __ballot_sync(0xffffffffu, r.finished);
        if (fin != 0u) {
            float ret = r.finished ? r.ep_ret : 0.0f;
            int len = r.finished ? r.ep_len : 0;
            int fnd = r.finished ? r.found : 0;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                ret += __shfl_xor_sync(0xffffffffu, ret, o);
                len += __shfl_xor_sync(0xffffffffu, len, o);
                fnd += __shfl_xor_sync(0xffffffffu, fnd, o);
            }
            const unsigned suc = __ballot_sync(0xffffffffu, r.finished && r.success);
            const unsigned cra = __ballot_sync(0xffffffffu, r.finished && r.crash);
            const unsigned tru = __ballot_sync(0xffffffffu, r.finished && r.done == DN_DONE_TRUNCATED);
            if ((tid & 31) == 0) {
                atomicAdd(&P.stats->return_sum, static_cast<double>(ret));
                atomicAdd(&P.stats->length_sum, static_cast<unsigned long long>(len));
                atomicAdd(&P.stats->episodes, static_cast<unsigned long long>(__popc(fin)));
                atomicAdd(&P.stats->found_targets, static_cast<unsigned long long>(fnd));
                if (suc) atomicAdd(&P.stats->successes, static_cast<unsigned long long>(__popc(suc)));
                if (cra) atomicAdd(&P.stats->crashes, static_cast<unsigned long long>(__popc(cra)));
                if (tru) atomicAdd(&P.stats->truncations, static_cast<unsigned long long>(__popc(tru)));
            }
        }
        // ---- observation tile: shared memory -> one TMA bulk store per CTA ----------
        if (write_out) {
            float* gdst = io.obs + (orow + base) * D;
            const uint32_t bytes = static_cast<uint32_t>(n_here) * D * 4u;
            const bool bulk_ok = ((reinterpret_cast<uintptr_t>(gdst) & 15u) == 0) && ((bytes & 15u) == 0);
            if (bulk_ok) {
                fence_proxy_async_smem();
                __syncthreads();
                if (tid == 0) {
                    bulk_store_g2s_commit(gdst, tile, bytes);
                    if (t + 1 < num_steps) bulk_store_wait_read();   // tile is rewritten by the next step
                }
                if (t + 1 < num_steps) __syncthreads();
            } else {
                __syncthreads();
                for (int j = tid; j < n_here * D; j += kBlock) gdst[j] = tile[j];
                if (t + 1 < num_steps) __syncthreads();
            }
        }
    }
    if (active) {
        store_state(P, i, s);
        if (PHYS & 1) P.last_rpm_sum[i] = last_rpm_sum;
    }
    if (tid == 0) bulk_store_wait_read();   // shared memory must outlive the bulk read
}

// ---------------------------------------------------------------------------
// construction-time state, explicit reset, state (un)packing
// ---------------------------------------------------------------------------
__global__ void init_kernel(const __grid_constant__ Params P, float d0) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.n) return;
    EnvState s;
    s.px = P.init_pos[0]; s.py = P.init_pos[1]; s.pz = P.init_pos[2]; s.dist = d0;
    s.qx = P.init_quat[0]; s.qy = P.init_quat[1]; s.qz = P.init_quat[2]; s.qw = P.init_quat[3];
    s.vx = s.vy = s.vz = 0.f; s.prev_dist = d0;
    s.wx = s.wy = s.wz = 0.f; s.ep_ret = 0.f;
    s.ax = s.ay = s.az = 0.f; s.bits = 0u;
    s.pvx = s.pvy = s.pvz = 0.f; s.ep_len = 0;
    s.pax = s.pay = s.paz = 0.f; s.ep_count = 0u;
    store_state(P, i, s);
    if (P.last_rpm_sum) P.last_rpm_sum[i] = 0.f;
    if (P.obs_rms) {
        const size_t N = P.n; const int D = P.obs_dim;
        for (int k = 0; k < D; ++k) { P.obs_rms[k * N + i] = 0.f; P.obs_rms[(D + k) * N + i] = 1.f; }
        P.obs_rms[2 * D * N + i] = 1e-4f;                   // RunningMeanStd(epsilon=1e-4), normalize.py:13-17
    }
}

// PBDroneEnv.reset called explicitly (VecEnv.reset / evaluate_policy), PBDroneEnv.py:609-665
template <bool NORM>
__global__ void reset_kernel(const __grid_constant__ Params P, const uint8_t* mask, float* obs_out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.n) return;
    if (mask && !mask[i]) return;
    EnvState s;
    load_state(P, i, s);
    const int steps = static_cast<int>(s.bits & kStepsMask);
    const float stale_dist = s.dist;
    float D0 = s.dist;
    if (steps != 0) {                           // _current_position == pos of the last post-step
        const float4 t0 = __ldg(&P.targets[0]);
        const float dx = s.px - t0.x, dy = s.py - t0.y, dz = s.pz - t0.z;
        D0 = sqrtf(dx * dx + dy * dy + dz * dz);
    }
    s.px = P.init_pos[0]; s.py = P.init_pos[1]; s.pz = P.init_pos[2];
    s.qx = P.init_quat[0]; s.qy = P.init_quat[1]; s.qz = P.init_quat[2]; s.qw = P.init_quat[3];
    s.vx = s.vy = s.vz = 0.f; s.wx = s.wy = s.wz = 0.f; s.ax = s.ay = s.az = 0.f;
    s.pvx = s.pvy = s.pvz = 0.f; s.pax = s.pay = s.paz = 0.f;
    s.dist = D0; s.prev_dist = D0; s.bits = 0u;
    s.ep_ret = 0.f; s.ep_len = 0;               // Monitor.reset
    store_state(P, i, s);
    if (P.last_rpm_sum) P.last_rpm_sum[i] = 0.f;
    const int D = P.obs_dim;
    float o[kMaxObs];
#pragma unroll
    for (int k = 0; k < 12; ++k) o[k] = P.init_obs[k];
    o[12] = clip_f32_range(stale_dist / P.max_target_dist);
    if (NORM) {
        const size_t N = P.n;
        float* cnt_p = P.obs_rms + static_cast<size_t>(2 * D) * N + i;
        const float cnt = *cnt_p;
        for (int k = 0; k < D; ++k) {
            float* mp = P.obs_rms + static_cast<size_t>(k) * N + i;
            float* vp = P.obs_rms + static_cast<size_t>(D + k) * N + i;
            float m = *mp, v = *vp;
            o[k] = rms_update_normalize(o[k], m, v, cnt);
            *mp = m; *vp = v;
        }
        *cnt_p = cnt + 1.0f;
    }
    if (obs_out) for (int k = 0; k < D; ++k) obs_out[static_cast<size_t>(i) * D + k] = o[k];
}

struct StateView {   // device mirror of dn_state_view
    float *pos, *quat, *vel, *rpy_rates, *ang_v, *prev_vel, *prev_ang_v, *dist, *prev_dist;
    int32_t *target_idx, *steps; uint8_t* just_found; float* ep_return; int32_t* ep_length;
    uint32_t* episode_count; float* last_rpm_sum; float* obs_rms;
};

template <bool SET>
__global__ void state_xfer_kernel(const __grid_constant__ Params P, const __grid_constant__ StateView V) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.n) return;
    EnvState s;
    load_state(P, i, s);
    int idx = static_cast<int>(s.bits >> kIdxShift);
    int steps = static_cast<int>(s.bits & kStepsMask);
    int jf = (s.bits & kJustFoundBit) ? 1 : 0;
#define DN_V3(ptr, a, b, c) if (V.ptr) { if (SET) { a = V.ptr[3*i]; b = V.ptr[3*i+1]; c = V.ptr[3*i+2]; } else { V.ptr[3*i] = a; V.ptr[3*i+1] = b; V.ptr[3*i+2] = c; } }
#define DN_V1(ptr, a, T) if (V.ptr) { if (SET) { a = static_cast<decltype(a)>(V.ptr[i]); } else { V.ptr[i] = static_cast<T>(a); } }
    DN_V3(pos, s.px, s.py, s.pz)
    if (V.quat) {
        if (SET) { s.qx = V.quat[4*i]; s.qy = V.quat[4*i+1]; s.qz = V.quat[4*i+2]; s.qw = V.quat[4*i+3]; }
        else { V.quat[4*i] = s.qx; V.quat[4*i+1] = s.qy; V.quat[4*i+2] = s.qz; V.quat[4*i+3] = s.qw; }
    }
    DN_V3(vel, s.vx, s.vy, s.vz)
    DN_V3(rpy_rates, s.wx, s.wy, s.wz)
    DN_V3(ang_v, s.ax, s.ay, s.az)
    DN_V3(prev_vel, s.pvx, s.pvy, s.pvz)
    DN_V3(prev_ang_v, s.pax, s.pay, s.paz)
    DN_V1(dist, s.dist, float)
    DN_V1(prev_dist, s.prev_dist, float)
    DN_V1(target_idx, idx, int32_t)
    DN_V1(steps, steps, int32_t)
    DN_V1(just_found, jf, uint8_t)
    DN_V1(ep_return, s.ep_ret, float)
    DN_V1(ep_length, s.ep_len, int32_t)
    DN_V1(episode_count, s.ep_count, uint32_t)
#undef DN_V3
#undef DN_V1
    if (V.last_rpm_sum && P.last_rpm_sum) {
        if (SET) P.last_rpm_sum[i] = V.last_rpm_sum[i]; else V.last_rpm_sum[i] = P.last_rpm_sum[i];
    }
    if (V.obs_rms && P.obs_rms) {
        const int W = 2 * P.obs_dim + 1; const size_t N = P.n;
        for (int k = 0; k < W; ++k) {
            if (SET) P.obs_rms[k * N + i] = V.obs_rms[static_cast<size_t>(i) * W + k];
            else V.obs_rms[static_cast<size_t>(i) * W + k] = P.obs_rms[k * N + i];
        }
    }
    if (SET) {
        s.bits = (static_cast<uint32_t>(idx) << kIdxShift) | (jf ? kJustFoundBit : 0u) | (static_cast<uint32_t>(steps) & kStepsMask);
        store_state(P, i, s);
    }
}

}  // namespace dn

// ===========================================================================
// host side: the C ABI
// ===========================================================================
using dn::Params;

struct dn_env {
    Params P;
    int device;
    int normalize_obs;
    void* state_mem;
    float4* d_targets;
    float4* d_segs;
    dn::Stats* d_stats;
    int64_t launches;
    float d0;
};

static thread_local std::string g_err;

static int fail(int code, const std::string& msg) { g_err = msg; return code; }

#define DN_CUDA(expr)                                                                    \
    do {                                                                                 \
        cudaError_t _e = (expr);                                                         \
        if (_e != cudaSuccess)                                                           \
            return fail(DN_ECUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));   \
    } while (0)

namespace {
struct DeviceGuard {
    int prev = -1; bool ok = false;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) == cudaSuccess && cudaSetDevice(dev) == cudaSuccess) ok = true;
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

// CF2X constants: Sol/resources/safegym/cf2x.urdf:5,11-12,34 ; derived BaseAviary.py:76,163-176
struct CF2X {
    static constexpr double M = 0.027, L = 0.0397, T2W = 2.25;
    static constexpr double IXX = 1.4e-5, IYY = 1.4e-5, IZZ = 2.17e-5;
    static constexpr double KF = 3.16e-10, KM = 7.94e-12;
    static constexpr double COLLISION_H = 0.025;
    static constexpr double GND_EFF_COEFF = 11.36859, PROP_RADIUS = 2.31348e-2;
    static constexpr double DRAG_XY = 9.1785e-7, DRAG_Z = 10.311e-7;
    static constexpr double PWM2RPM_SCALE = 0.2685, PWM2RPM_CONST = 4070.3, MIN_PWM = 20000.0, MAX_PWM = 65535.0;
    static constexpr double G = 9.8;
};

void quat_from_euler(const double rpy[3], double q[4]) {   // p.getQuaternionFromEuler (BaseAviary.py:567)
    const double r = rpy[0] * 0.5, p = rpy[1] * 0.5, y = rpy[2] * 0.5;
    const double cr = std::cos(r), sr = std::sin(r), cp = std::cos(p), sp = std::sin(p), cy = std::cos(y), sy = std::sin(y);
    q[0] = sr * cp * cy - cr * sp * sy;
    q[1] = cr * sp * cy + sr * cp * sy;
    q[2] = cr * cp * sy - sr * sp * cy;
    q[3] = cr * cp * cy + sr * sp * sy;
    const double n = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    for (int k = 0; k < 4; ++k) q[k] /= n;
}

void euler_from_quat(const double q[4], double rpy[3]) {    // p.getEulerFromQuaternion (BaseAviary.py:597)
    const double x = q[0], y = q[1], z = q[2], w = q[3];
    const double sarg = -2.0 * (x * z - w * y);
    const double pi = 3.14159265358979323846;
    if (sarg <= -0.99999) { rpy[0] = 0; rpy[1] = -0.5 * pi; rpy[2] = 2 * std::atan2(x, -y); }
    else if (sarg >= 0.99999) { rpy[0] = 0; rpy[1] = 0.5 * pi; rpy[2] = 2 * std::atan2(-x, y); }
    else {
        rpy[0] = std::atan2(2 * (y * z + w * x), w * w - x * x - y * y + z * z);
        rpy[1] = std::asin(sarg);
        rpy[2] = std::atan2(2 * (x * y + w * z), w * w + x * x - y * y - z * z);
    }
}

bool reward_table(int id, dn::RewardParams& w) {
    switch (id) {
        case DN_REWARD_DEFAULT:    // PBDroneEnv.py:475-607
            w = {-10.f, 200.f, 75.f, 5.f, 3.f, 2.f, 3000.f, 3.f, 0.7f, 0.3f, 1.f, 25.f}; return true;
        case DN_REWARD_DUMMY:      // dummy_env.py:446-550,587-598 (smoothness thresholds 0.1 / 0.1)
            w = {-10.f, 200.f, 75.f, 5.f, 3.f, 2.f, 3000.f, 3.f, 0.1f, 0.1f, 1.f, 25.f}; return true;
        case DN_REWARD_THRUSTENV:  // ThrustEnv.py:368-463 (-4 crash, +25 / +1000, 20 x progress, no orientation / smoothness)
            w = {-4.f, 1000.f, 25.f, 0.f, 3.f, 2.f, 20.f, 0.f, 0.f, 0.f, 0.f, 25.f}; return true;
        default: return false;
    }
}
}  // namespace

extern "C" {

int dn_abi_version(void) { return DN_ABI_VERSION; }
const char* dn_last_error(void) { return g_err.c_str(); }

int dn_create(const dn_config* cfg, int device, dn_env** out) {
    if (!cfg || !out) return fail(DN_EINVAL, "dn_create: null argument");
    *out = nullptr;
    if (cfg->abi_version != DN_ABI_VERSION) return fail(DN_EINVAL, "dn_create: abi_version mismatch");
    if (cfg->num_envs <= 0) return fail(DN_EINVAL, "dn_create: num_envs must be > 0");
    if (cfg->pyb_freq <= 0 || cfg->ctrl_freq <= 0 || cfg->pyb_freq % cfg->ctrl_freq != 0)
        return fail(DN_EINVAL, "dn_create: pyb_freq is not divisible by ctrl_freq");   // BaseAviary.py:81-82
    if (cfg->num_targets <= 0 || cfg->num_targets >= 2047 || !cfg->targets)
        return fail(DN_EINVAL, "dn_create: need 1..2046 targets");
    if (cfg->act_type < 0 || cfg->act_type > DN_ACT_ONE_D_RPM) return fail(DN_EINVAL, "dn_create: unsupported act_type");
    if (cfg->physics & ~7) return fail(DN_EINVAL, "dn_create: unknown physics flags");
    if (cfg->spawn_mode != DN_SPAWN_FIXED) return fail(DN_EINVAL, "dn_create: spawn_mode not implemented");
    if (cfg->max_steps < 0 || cfg->max_steps > (int)dn::kStepsMask - 1) return fail(DN_EINVAL, "dn_create: max_steps out of range");
    dn::RewardParams rw;
    if (!reward_table(cfg->reward_id, rw)) return fail(DN_EINVAL, "dn_create: reward_id not implemented");

    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(DN_ECUDA, "dn_create: no CUDA device (libdronenav has no CPU fallback)");
    if (device < 0 || device >= ndev) return fail(DN_EINVAL, "dn_create: bad device index");
    DeviceGuard guard(device);
    if (!guard.ok) return fail(DN_ECUDA, "dn_create: cudaSetDevice failed");

    dn_env* e = new (std::nothrow) dn_env();
    if (!e) return fail(DN_ENOMEM, "dn_create: out of host memory");
    std::memset(e, 0, sizeof(*e));
    e->device = device;
    e->normalize_obs = cfg->normalize_obs ? 1 : 0;
    Params& P = e->P;
    const int N = cfg->num_envs, T = cfg->num_targets;
    P.n = N;
    P.substeps = cfg->pyb_freq / cfg->ctrl_freq;
    P.act_type = cfg->act_type;
    P.normalize_actions = cfg->normalize_actions ? 1 : 0;
    P.physics = cfg->physics;
    P.obs_dim = cfg->include_distance ? 13 : 12;
    P.cylinder = cfg->cylinder ? 1 : 0;
    P.circle = cfg->circle ? 1 : 0;
    P.max_steps = cfg->max_steps;
    P.num_targets = T;
    P.spawn_mode = cfg->spawn_mode;
    P.reward_id = cfg->reward_id;
    P.dt = static_cast<float>(1.0 / cfg->pyb_freq);
    P.threshold = static_cast<float>(cfg->threshold);
    const double* ad = cfg->aviary_dim;
    P.x_low = (float)ad[0]; P.y_low = (float)ad[1]; P.z_low = (float)ad[2];
    P.x_high = (float)ad[3]; P.y_high = (float)ad[4]; P.z_high = (float)ad[5];
    const double mtd = std::fmax(std::fmax(std::fabs(ad[0]) + ad[3], std::fabs(ad[1]) + ad[4]), ad[5]);   // PBDroneEnv.py:91
    P.max_target_dist = static_cast<float>(mtd);
    double q0[4];
    quat_from_euler(cfg->init_rpy, q0);
    for (int k = 0; k < 3; ++k) { P.init_pos[k] = (float)cfg->init_xyz[k]; P.init_seg_base[k] = (float)cfg->init_xyz[k]; }
    for (int k = 0; k < 4; ++k) P.init_quat[k] = (float)q0[k];
    {   // observation of the spawn pose, entries 0..11 (PBDroneEnv.py:338-398), in double
        double rpy[3];
        euler_from_quat(q0, rpy);
        const double pi = 3.14159265358979323846;
        double o[12] = {cfg->init_xyz[0] / ad[3], cfg->init_xyz[1] / ad[4], cfg->init_xyz[2] / ad[5],
                        std::fmin(std::fmax(rpy[0], -pi), pi) / pi, std::fmin(std::fmax(rpy[1], -pi), pi) / pi, rpy[2] / pi,
                        0, 0, 0, 0, 0, 0};
        for (int k = 0; k < 12; ++k) P.init_obs[k] = (float)o[k];
    }
    // action map constants: float32 like the reference (PBDroneEnv.py:113-116)
    const double a_low = CF2X::KF * std::pow(CF2X::PWM2RPM_SCALE * CF2X::MIN_PWM + CF2X::PWM2RPM_CONST, 2);
    const double a_high = CF2X::KF * std::pow(CF2X::PWM2RPM_SCALE * CF2X::MAX_PWM + CF2X::PWM2RPM_CONST, 2);
    P.a_low = (float)a_low; P.a_high = (float)a_high;
    P.kf = (float)CF2X::KF; P.km = (float)CF2X::KM;
    P.pwm_scale = (float)CF2X::PWM2RPM_SCALE; P.pwm_const = (float)CF2X::PWM2RPM_CONST;
    P.pwm_min = (float)CF2X::MIN_PWM; P.pwm_max = (float)CF2X::MAX_PWM;
    const double gravity = CF2X::G * CF2X::M;
    const double hover_rpm = std::sqrt(gravity / (4 * CF2X::KF));
    const double max_rpm = std::sqrt((CF2X::T2W * gravity) / (4 * CF2X::KF));
    const double max_thrust = 4 * CF2X::KF * max_rpm * max_rpm;
    P.hover_rpm = (float)hover_rpm;
    P.gravity = (float)gravity; P.inv_m = (float)(1.0 / CF2X::M);
    P.arm_over_sqrt2 = (float)(CF2X::L / std::sqrt(2.0));
    P.ixx = (float)CF2X::IXX; P.iyy = (float)CF2X::IYY; P.izz = (float)CF2X::IZZ;
    P.inv_ixx = (float)(1.0 / CF2X::IXX); P.inv_iyy = (float)(1.0 / CF2X::IYY); P.inv_izz = (float)(1.0 / CF2X::IZZ);
    P.drag_xy = (float)CF2X::DRAG_XY; P.drag_z = (float)CF2X::DRAG_Z;
    P.gnd_coeff = (float)CF2X::GND_EFF_COEFF; P.prop_radius = (float)CF2X::PROP_RADIUS;
    P.gnd_h_clip = (float)(0.25 * CF2X::PROP_RADIUS * std::sqrt((15 * max_rpm * max_rpm * CF2X::KF * CF2X::GND_EFF_COEFF) / max_thrust));
    P.collision_half_h = (float)(CF2X::COLLISION_H / 2);
    const double px[4] = {0.028, -0.028, -0.028, 0.028}, py[4] = {0.028, 0.028, -0.028, -0.028};   // safegym/cf2x.urdf:42,54,66,78
    for (int k = 0; k < 4; ++k) { P.prop_x[k] = (float)px[k]; P.prop_y[k] = (float)py[k]; }
    P.rw = rw;
    P.seed = cfg->seed;
    P.env_id_offset = cfg->env_id_offset;

    // target table + segment table for the non-circle cylinder (PBDroneEnv.py:746-786), in double
    std::vector<float4> h_t(T), h_s(2 * T);
    for (int k = 0; k < T; ++k) {
        const double* tk = cfg->targets + 3 * k;
        h_t[k] = make_float4((float)tk[0], (float)tk[1], (float)tk[2], 0.f);
        const double* b1 = (k == 0) ? cfg->init_xyz : cfg->targets + 3 * (k - 1);
        double lv[3] = {tk[0] - b1[0], tk[1] - b1[1], tk[2] - b1[2]};
        const double len = std::sqrt(lv[0] * lv[0] + lv[1] * lv[1] + lv[2] * lv[2]);
        if (len == 0.0) {
            h_s[2 * k] = make_float4((float)b1[0], (float)b1[1], (float)b1[2], 0.f);
            h_s[2 * k + 1] = make_float4(0.f, 0.f, 0.f, 0.f);
        } else {
            const double u[3] = {lv[0] / len, lv[1] / len, lv[2] / len};
            const double e1[3] = {b1[0] - 0.2 * u[0], b1[1] - 0.2 * u[1], b1[2] - 0.2 * u[2]};
            const double e2[3] = {tk[0] + 0.2 * u[0], tk[1] + 0.2 * u[1], tk[2] + 0.2 * u[2]};
            const double el = std::sqrt((e2[0] - e1[0]) * (e2[0] - e1[0]) + (e2[1] - e1[1]) * (e2[1] - e1[1]) + (e2[2] - e1[2]) * (e2[2] - e1[2]));
            h_s[2 * k] = make_float4((float)e1[0], (float)e1[1], (float)e1[2], (float)el);
            h_s[2 * k + 1] = make_float4((float)u[0], (float)u[1], (float)u[2], (float)len);
        }
    }
    // constructor distance: ||INIT_XYZS[0] - target[0]|| (PBDroneEnv.py:137-138)
    {
        const double* t0 = cfg->targets;
        const double dx = cfg->init_xyz[0] - t0[0], dy = cfg->init_xyz[1] - t0[1], dz = cfg->init_xyz[2] - t0[2];
        e->d0 = (float)std::sqrt(dx * dx + dy * dy + dz * dz);
    }

    auto cleanup = [&](int code, const std::string& msg) {
        if (e->state_mem) cudaFree(e->state_mem);
        if (e->d_targets) cudaFree(e->d_targets);
        if (e->d_segs) cudaFree(e->d_segs);
        if (e->d_stats) cudaFree(e->d_stats);
        delete e;
        return fail(code, msg);
    };
    // persistent state: 7 float4 planes (+ optional drag / obs-RMS planes), one allocation
    const size_t plane = ((static_cast<size_t>(N) * sizeof(float4) + 255) / 256) * 256;
    const size_t fplane = ((static_cast<size_t>(N) * sizeof(float) + 255) / 256) * 256;
    size_t bytes = dn::kPlanes * plane;
    const bool drag = (cfg->physics & DN_PHYS_DRAG) != 0;
    if (drag) bytes += fplane;
    const size_t rms_floats = e->normalize_obs ? static_cast<size_t>(2 * P.obs_dim + 1) * N : 0;
    bytes += ((rms_floats * sizeof(float) + 255) / 256) * 256;
    cudaError_t ce = cudaMalloc(&e->state_mem, bytes);
    if (ce != cudaSuccess) return cleanup(DN_ENOMEM, std::string("dn_create: cudaMalloc state: ") + cudaGetErrorString(ce));
    char* p = static_cast<char*>(e->state_mem);
    for (int k = 0; k < dn::kPlanes; ++k) { P.s[k] = reinterpret_cast<float4*>(p); p += plane; }
    if (drag) { P.last_rpm_sum = reinterpret_cast<float*>(p); p += fplane; }
    if (rms_floats) P.obs_rms = reinterpret_cast<float*>(p);
    if ((ce = cudaMalloc(&e->d_targets, T * sizeof(float4))) != cudaSuccess ||
        (ce = cudaMalloc(&e->d_segs, 2 * T * sizeof(float4))) != cudaSuccess ||
        (ce = cudaMalloc(&e->d_stats, sizeof(dn::Stats))) != cudaSuccess)
        return cleanup(DN_ENOMEM, std::string("dn_create: cudaMalloc tables: ") + cudaGetErrorString(ce));
    if ((ce = cudaMemcpy(e->d_targets, h_t.data(), T * sizeof(float4), cudaMemcpyHostToDevice)) != cudaSuccess ||
        (ce = cudaMemcpy(e->d_segs, h_s.data(), 2 * T * sizeof(float4), cudaMemcpyHostToDevice)) != cudaSuccess ||
        (ce = cudaMemset(e->d_stats, 0, sizeof(dn::Stats))) != cudaSuccess)
        return cleanup(DN_ECUDA, std::string("dn_create: table upload: ") + cudaGetErrorString(ce));
    P.targets = e->d_targets; P.segs = e->d_segs; P.stats = e->d_stats;

    dn::init_kernel<<<(N + 255) / 256, 256>>>(P, e->d0);
    ce = cudaGetLastError();
    if (ce == cudaSuccess) ce = cudaDeviceSynchronize();
    if (ce != cudaSuccess) return cleanup(DN_ECUDA, std::string("dn_create: init_kernel: ") + cudaGetErrorString(ce));
    e->launches = 1;
    *out = e;
    return DN_OK;
}

int dn_destroy(dn_env* env) {
    if (!env) return DN_OK;
    DeviceGuard guard(env->device);
    cudaFree(env->state_mem); cudaFree(env->d_targets); cudaFree(env->d_segs); cudaFree(env->d_stats);
    delete env;
    return DN_OK;
}

int dn_num_envs(const dn_env* env) { return env ? env->P.n : fail(DN_EINVAL, "null handle"); }
int dn_obs_dim(const dn_env* env) { return env ? env->P.obs_dim : fail(DN_EINVAL, "null handle"); }
int64_t dn_launch_count(const dn_env* env) { return env ? env->launches : 0; }

int dn_reset(dn_env* env, const uint8_t* mask, float* obs_out, void* stream) {
    if (!env) return fail(DN_EINVAL, "dn_reset: null handle");
    DeviceGuard guard(env->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int N = env->P.n;
    if (env->normalize_obs) dn::reset_kernel<true><<<(N + 255) / 256, 256, 0, st>>>(env->P, mask, obs_out);
    else dn::reset_kernel<false><<<(N + 255) / 256, 256, 0, st>>>(env->P, mask, obs_out);
    DN_CUDA(cudaGetLastError());
    env->launches += 1;
    return DN_OK;
}

static int launch_step(dn_env* env, const dn_step_io* io, int num_steps, int per_step, void* stream) {
    if (!env || !io) return fail(DN_EINVAL, "dn_step: null argument");
    if (!io->actions || !io->obs || !io->reward || !io->done) return fail(DN_EINVAL, "dn_step: actions/obs/reward/done are required");
    if (num_steps <= 0) return fail(DN_EINVAL, "dn_step_many: num_steps must be > 0");
    if (reinterpret_cast<uintptr_t>(io->actions) & 15u) return fail(DN_EINVAL, "dn_step: actions must be 16-byte aligned");
    DeviceGuard guard(env->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    dn::StepIO k;
    k.actions = reinterpret_cast<const float4*>(io->actions);
    k.obs = io->obs; k.reward = io->reward; k.done = io->done; k.terminal_obs = io->terminal_obs;
    k.found_targets = io->found_targets; k.episode_return = io->episode_return; k.episode_length = io->episode_length;
    const int N = env->P.n;
    const dim3 grid((N + dn::kBlock - 1) / dn::kBlock), block(dn::kBlock);
    const int phys = env->P.physics & 3;
#define DN_LAUNCH(PH, NO) dn::step_kernel<PH, NO><<<grid, block, 0, st>>>(env->P, k, num_steps, per_step)
    if (env->normalize_obs) {
        switch (phys) { case 0: DN_LAUNCH(0, true); break; case 1: DN_LAUNCH(1, true); break;
                        case 2: DN_LAUNCH(2, true); break; default: DN_LAUNCH(3, true); break; }
    } else {
        switch (phys) { case 0: DN_LAUNCH(0, false); break; case 1: DN_LAUNCH(1, false); break;
                        case 2: DN_LAUNCH(2, false); break; default: DN_LAUNCH(3, false); break; }
    }
#undef DN_LAUNCH
    DN_CUDA(cudaGetLastError());
    env->launches += 1;
    return DN_OK;
}

int dn_step(dn_env* env, const dn_step_io* io, void* stream) { return launch_step(env, io, 1, 1, stream); }

int dn_step_many(dn_env* env, const dn_step_io* io, int num_steps, int per_step_outputs, void* stream) {
    return launch_step(env, io, num_steps, per_step_outputs ? 1 : 0, stream);
}

static int state_xfer(dn_env* env, const dn_state_view* v, bool set, void* stream) {
    if (!env || !v) return fail(DN_EINVAL, "dn_get/set_state: null argument");
    DeviceGuard guard(env->device);
    dn::StateView V;
    V.pos = v->pos; V.quat = v->quat; V.vel = v->vel; V.rpy_rates = v->rpy_rates; V.ang_v = v->ang_v;
    V.prev_vel = v->prev_vel; V.prev_ang_v = v->prev_ang_v; V.dist = v->dist; V.prev_dist = v->prev_dist;
    V.target_idx = v->target_idx; V.steps = v->steps; V.just_found = v->just_found; V.ep_return = v->ep_return;
    V.ep_length = v->ep_length; V.episode_count = v->episode_count; V.last_rpm_sum = v->last_rpm_sum; V.obs_rms = v->obs_rms;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int N = env->P.n;
    if (set) dn::state_xfer_kernel<true><<<(N + 255) / 256, 256, 0, st>>>(env->P, V);
    else dn::state_xfer_kernel<false><<<(N + 255) / 256, 256, 0, st>>>(env->P, V);
    DN_CUDA(cudaGetLastError());
    env->launches += 1;
    return DN_OK;
}

int dn_get_state(dn_env* env, const dn_state_view* view, void* stream) { return state_xfer(env, view, false, stream); }
int dn_set_state(dn_env* env, const dn_state_view* view, void* stream) { return state_xfer(env, view, true, stream); }

int dn_episode_stats(dn_env* env, dn_stats* host_out, int clear, void* stream) {
    if (!env || !host_out) return fail(DN_EINVAL, "dn_episode_stats: null argument");
    DeviceGuard guard(env->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    dn::Stats h;
    DN_CUDA(cudaMemcpyAsync(&h, env->d_stats, sizeof(h), cudaMemcpyDeviceToHost, st));
    if (clear) DN_CUDA(cudaMemsetAsync(env->d_stats, 0, sizeof(dn::Stats), st));
    DN_CUDA(cudaStreamSynchronize(st));
    host_out->return_sum = h.return_sum; host_out->length_sum = h.length_sum; host_out->episodes = h.episodes;
    host_out->successes = h.successes; host_out->found_targets = h.found_targets; host_out->crashes = h.crashes;
    host_out->truncations = h.truncations;
    return DN_OK;
}

}  // extern "C"
