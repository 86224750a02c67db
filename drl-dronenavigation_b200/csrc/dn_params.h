// dn_params.h -- kernel-side parameter block of libdronenav (host + device).
//
// Everything a control step needs that is not per-env state travels in ONE by-value
// kernel argument (`__grid_constant__`, i.e. the constant bank): CF2X constants
// (Sol/resources/safegym/cf2x.urdf:5,11-12,34; Sol/PyBullet/BaseAviary.py:76,163-176),
// the PBDroneEnv constructor arguments (Sol/Model/Environments/PBDroneEnv.py:41-169) and
// pointers to the handle's HBM-resident state.
#pragma once
#include <stdint.h>
#ifndef DN_HOST_EMU
#include <cuda_runtime.h>
#endif

namespace dn {

// One reward family covers PBDroneEnv._computeReward (PBDroneEnv.py:475-571) and its
// variants dummy_env.py:446-550 / ThrustEnv.py:368-513: same state machine, different
// constants.
struct RewardParams {
    float crash;             // returned as is (NOT divided), PBDroneEnv.py:489-490
    float final_bonus;       // all targets reached, :544
    float capture_bonus;     // one target reached, :550
    float capture_orient_w;  // :551
    float exp_w, exp_k;      // exp_w * exp(-exp_k * d), :555
    float progress_w;        // (prev_d - d) * progress_w unless just_found, :556
    float orient_w;          // :557
    float smooth_lin_thr, smooth_ang_thr;  // :599
    float smooth_w;          // 1 = smoothness term on, 0 = off (ThrustEnv)
    float divisor;           // :571
    float inv_divisor;       // 1 / divisor (host, double)
    // ---- other reward families (reward_id >= 3), see dn_device.cuh::reward_alt
    int   mode;              // RW_* below
    float decay_log2;        // HER: log2(discount) / 10, capture bonus * discount^(steps/10) (HerPBDroneEnv.py:371)
    float proj_w;            // DN_REWARD_PROGRESS: weight of the projection progress that replaces (prev_d - d) * progress_w; 0 = off
    // LITERATURE: Rewarder.BootstrappedImiVisionRewardCalculator / ChampRewardCalculator (Rewarder.py:66-150) as one formula:
    // prog (prev_d - d) + perc_poly dc^4 + perc_exp_w exp(perc_exp_k dc^4) + da1 |da| + da2 |da|^2 + w1 |w| + w2 |w|^2
    // + pass [passed] - crash [crashed (or p_z < 0 if lit_pz)]
    float lit_prog, lit_perc_poly, lit_perc_exp_w, lit_perc_exp_k, lit_da1, lit_da2, lit_w1, lit_w2, lit_pass, lit_crash;
    int   lit_pz;
    float pt_x, pt_y_rate, pt_z, pt_w;   // POINT: -pt_w * |(pt_x, pt_y_rate * t_norm, pt_z) - pos|^2 (HoverAviary.py:65-76, FlyThruGateAviary.py:100-112)
};
enum { RW_WAYPOINT = 0, RW_HER = 1, RW_REACHING = 2, RW_POINT = 3, RW_LITERATURE = 4 };

struct Stats {               // device mirror of dn_stats
    double             return_sum;
    unsigned long long length_sum, episodes, successes, found_targets, crashes, truncations;
};
typedef Stats BlockStats;    // one accumulation slot per CTA of the step grid (no atomics)

// Per-env persistent state in HBM: seven float4 planes ("structure of float4 arrays"),
// plane p of env i at s[p][i].  One LDG.128 / STG.128 per plane per env, consecutive
// lanes touch consecutive 16-byte words -> every warp request is 512 contiguous bytes.
//   s0 = pos.xyz        | dist               (BaseAviary.pos ; PBDroneEnv._distance_to_target)
//   s1 = quat.xyzw                           (BaseAviary.quat)
//   s2 = vel.xyz        | prev_dist          (BaseAviary.vel ; _prev_distance_to_target)
//   s3 = rpy_rates.xyz  | ep_return          (BaseAviary.rpy_rates ; Monitor)
//   s4 = ang_v.xyz      | bits{steps:20, just_found:1, target_idx:11}
//   s5 = prev_vel.xyz   | ep_length (int bits)
//   s6 = prev_ang_v.xyz | episode_count (uint bits, Philox counter)
// current_vel / current_ang_v of the reference are vel / ang_v at step entry and are
// not stored.  112 B read + 112 B written per env-step.
constexpr int kPlanes = 7;
constexpr int kConstTargets = 16;

struct Params {
    int   n;
    int   substeps;
    int   act_type;
    int   normalize_actions;
    int   physics;
    int   obs_dim;
    int   cylinder;
    int   circle;
    int   max_steps;
    int   num_targets;
    int   spawn_mode;
    int   reward_id;
    float dt;
    float threshold;
    float x_low, y_low, z_low, x_high, y_high, z_high;
    float max_target_dist;
    float inv_x_high, inv_y_high, inv_z_high, inv_max_target_dist;   // reciprocals, computed in double
    float thr2;              // threshold^2
    float cyl_limit2;        // (threshold + 0.2)^2, segment tube (PBDroneEnv.py:786)
    float init_pos[3];
    float init_quat[4];
    float init_obs[12];      // observation of the spawn pose (entries 0..11)
    float init_seg_base[3];  // INIT_XYZS[0], base of segment 0 (PBDroneEnv.py:746-748)
    // action map (PBDroneEnv.py:113-116,872-895,949-971; env_utils.py:8-59)
    float a_low, a_high, kf, km, pwm_scale, pwm_const, pwm_min, pwm_max, hover_rpm;
    float a_span, inv_a_span, inv_kf, inv_pwm_scale;   // float32 divisors and their RN reciprocals (div_const_rn)
    // rigid body (BaseAviary.py:899-958)
    float gravity, inv_m, torque_arm, ixx, iyy, izz, inv_ixx, inv_iyy, inv_izz;   // torque_arm: L / sqrt(2) (CF2X, RACE), L (CF2P)
    int   frame_plus;        // 1: DroneModel.CF2P torque mix (BaseAviary.py:933-935); km is negated for RACE (:927-928)
    // add-ons (BaseAviary.py:798-865)
    float drag_xy, drag_z, gnd_coeff, prop_radius, gnd_h_clip, collision_half_h;
    float prop_x[4], prop_y[4];
    RewardParams rw;
    unsigned long long seed;
    long long env_id_offset;
    // Tracks of up to kConstTargets targets (circle: 6, reaching: 8) travel in the constant bank with the rest of
    // this block: a dependent global load of targets[idx] in the middle of the epilogue costs an L2 round trip per
    // thread, an indexed constant load does not.  Longer tracks use the global tables below.
    int   const_tables;      // 1: tgt_c / seg_c hold the tables
    float4 tgt_c[16];
    float4 seg_c[32];
    const float4* targets;   // [T] (x,y,z,0)
    const float4* segs;      // [T][2] {ext_p1.xyz, ext_len} {unit.xyz, seg_len}; non-circle cylinder
    float4* s[kPlanes];
    float* last_rpm_sum;     // [N], drag only
    double* obs_rms;         // [(2*obs_dim+1)][N] FP64 mean planes | var planes | count, normalize_obs only
    float4* spawn;           // [N] {spawn point of the current episode, 0}; DN_SPAWN_LINE only (segment 0 of the tube starts there)
    float4* aux;             // [N] {_current_position.xyz (stale across resets), |_current_position - _last_position|} (RW_REACHING);
                             //     PBDroneEnv._last_action (RW_LITERATURE)
    float4* rew_rms;         // [N] {returns, mean, var, count} of normalize.NormalizeReward (normalize.py:100-147); normalize_reward only
    float rew_gamma, rew_eps, rew_clip;   // NormalizeReward gamma / epsilon; TransformReward clip bound (<= 0: off), PBDroneSimulator.py:190-193
    float ep_time_scale;     // S / (PYB_FREQ * EPISODE_LEN_SEC): step_counter / PYB_FREQ / EPISODE_LEN_SEC = ep_len * this
    // DSLPIDControl (Sol/PyBullet/DSLPIDControl.py:20-80) for DN_ACT_PID / VEL / ONE_D_PID; its constants are the CF2X ones
    // whatever the airframe (BaseSingleAgentAviary.py:72-73), gains are compile-time constants in dn_device.cuh
    float ctrl_dt, inv_ctrl_dt;   // CTRL_TIMESTEP = 1 / ctrl_freq and its reciprocal
    float speed_limit;            // 0.03 * MAX_SPEED_KMH * 1000 / 3600 (BaseSingleAgentAviary.py:91)
    float pid_gravity, pid_inv_4kf;   // BaseControl.GRAVITY = g * m, 1 / (4 KF) of the controller's URDF (cf2x)
    float4* pid[3];               // [N] integral_pos_e | integral_rpy_e | last_rpy (w unused); PID action types only
    BlockStats* block_stats; // [ceil(N / CTA)] Monitor statistics slots
    int prefetch_ctas;       // software-prefetch distance in CTAs (= CTAs resident on the whole GPU)
};

struct StepIO {
    const float4* actions;
    float*    obs;
    float*    reward;
    uint8_t*  done;
    float*    terminal_obs;
    int32_t*  found_targets;
    float*    episode_return;
    int32_t*  episode_length;
    int       pdl_prefetch;   // 1: launched as a programmatic dependent grid -> prefetch this thread's lines to L2 while waiting
    // dn_step_host, zero-copy path: the last CTA to finish writes `seq` to a pinned host word, which the host polls instead of the stream
    unsigned int* done_counter;   // device, zero between launches
    unsigned int* host_flag;      // device alias of the pinned word; nullptr: not a host call
    unsigned int  seq;
    // resident step server (dn_host_server; MULTI kernels only): the kernel stays on the GPU between host steps.  The host rings
    // `srv_doorbell` (pinned + mapped) with (sequence << 48) | device pointer of the step's actions (pointer 0: quit); CTA 0 polls
    // it and publishes every command -- or, after `srv_idle_us` without one, "leave" -- through `srv_cmd` (device memory) to the
    // other CTAs, so that all CTAs run exactly the same steps; completion of a step = the sequence number in `host_flag`.
    unsigned long long* srv_doorbell;
    unsigned long long* srv_cmd;
    unsigned long long  srv_word0;    // the word `srv_cmd` holds at launch (the last command of the previous residency)
    unsigned int        srv_idle_us;
};
constexpr unsigned long long kSrvPtrMask = (1ull << 48) - 1ull;

}  // namespace dn
