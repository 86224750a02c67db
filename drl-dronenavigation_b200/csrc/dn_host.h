// dn_host.h -- host-side derivation of the kernel parameter block from a dn_config:
// CF2X constants, derived BaseAviary quantities, the spawn observation and the target /
// segment tables, all computed in double and rounded once to FP32.
// Shared by dronenav.cu (the product) and tests/host_emu (a g++ build of the same step
// logic that lets the CPU test suite exercise the device code path without a GPU).
#pragma once
#include <cmath>
#include <vector>
#include "../../include/dronenav.h"
#include "dn_params.h"

namespace dn { namespace host {

// Airframe constants: Sol/resources/safegym/cf2x.urdf:5,11-12,34,42-78 ; Sol/resources/cf2p.urdf:5,11-12,34,42-78 ;
// Sol/resources/racer.urdf:5,11-12,28,36-72 ; derived quantities BaseAviary.py:76,163-176.  Only cf2x.urdf carries the
// pwm attributes (the THRUST action map needs them).
struct Airframe {
    double M, L, T2W, IXX, IYY, IZZ, KF, KM, MAX_SPEED_KMH;
    double PROP_RADIUS;
    double prop_x[4], prop_y[4];
    bool plus;          // CF2P torque arms (BaseAviary.py:933-935)
    bool km_negated;    // RACE: z_torques = -z_torques (BaseAviary.py:927-928)
    bool has_pwm;       // pwm2rpm_scale / pwm2rpm_const / pwm_min / pwm_max present in the URDF
    static constexpr double COLLISION_H = 0.025;
    static constexpr double GND_EFF_COEFF = 11.36859;
    static constexpr double DRAG_XY = 9.1785e-7, DRAG_Z = 10.311e-7;
    static constexpr double PWM2RPM_SCALE = 0.2685, PWM2RPM_CONST = 4070.3, MIN_PWM = 20000.0, MAX_PWM = 65535.0;
    static constexpr double G = 9.8;
};
inline const Airframe* airframe(int model) {
    static const Airframe cf2x = {0.027, 0.0397, 2.25, 1.4e-5, 1.4e-5, 2.17e-5, 3.16e-10, 7.94e-12, 30.0, 2.31348e-2,
                                  {0.028, -0.028, -0.028, 0.028}, {0.028, 0.028, -0.028, -0.028}, false, false, true};
    static const Airframe cf2p = {0.027, 0.0397, 2.25, 2.3951e-5, 2.3951e-5, 3.2347e-5, 3.16e-10, 7.94e-12, 30.0, 2.31348e-2,
                                  {0.0397, 0.0, -0.0397, 0.0}, {0.0, 0.0397, 0.0, -0.0397}, true, false, false};
    static const Airframe race = {0.830, 0.109, 4.17, 3.113e-3, 3.113e-3, 3.113e-3, 8.47e-9, 2.13e-11, 200.0, 12.7e-2,
                                  {0.085, -0.085, -0.085, 0.085}, {0.0675, 0.0675, -0.0675, -0.0675}, false, true, false};
    switch (model) {
        case DN_MODEL_CF2X: return &cf2x;
        case DN_MODEL_CF2P: return &cf2p;
        case DN_MODEL_RACE: return &race;
        default: return nullptr;
    }
}
inline void quat_from_euler(const double rpy[3], double q[4]) {   // p.getQuaternionFromEuler (BaseAviary.py:567)
    const double r = rpy[0] * 0.5, p = rpy[1] * 0.5, y = rpy[2] * 0.5;
    const double cr = std::cos(r), sr = std::sin(r), cp = std::cos(p), sp = std::sin(p), cy = std::cos(y), sy = std::sin(y);
    q[0] = sr * cp * cy - cr * sp * sy;
    q[1] = cr * sp * cy + sr * cp * sy;
    q[2] = cr * cp * sy - sr * sp * cy;
    q[3] = cr * cp * cy + sr * sp * sy;
    const double n = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    for (int k = 0; k < 4; ++k) q[k] /= n;
}

inline void euler_from_quat(const double q[4], double rpy[3]) {    // p.getEulerFromQuaternion (BaseAviary.py:597)
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

inline bool reward_table(int id, double discount, dn::RewardParams& w) {
    w = dn::RewardParams{};
    w.divisor = 1.f; w.inv_divisor = 1.f; w.mode = dn::RW_WAYPOINT;
    switch (id) {
        case DN_REWARD_DEFAULT:    // PBDroneEnv.py:475-607
            w = {-10.f, 200.f, 75.f, 5.f, 3.f, 2.f, 3000.f, 3.f, 0.7f, 0.3f, 1.f, 25.f, 0.04f}; return true;
        case DN_REWARD_DUMMY:      // dummy_env.py:446-550,587-598 (smoothness thresholds 0.1 / 0.1)
            w = {-10.f, 200.f, 75.f, 5.f, 3.f, 2.f, 3000.f, 3.f, 0.1f, 0.1f, 1.f, 25.f, 0.04f}; return true;
        case DN_REWARD_THRUSTENV:  // ThrustEnv.py:368-463 (-4 crash, +25 / +1000, 20 x progress, no orientation / smoothness)
            w = {-4.f, 1000.f, 25.f, 0.f, 3.f, 2.f, 20.f, 0.f, 0.f, 0.f, 0.f, 25.f, 0.04f}; return true;
        case DN_REWARD_HER:        // HerPBDroneEnv.py:314-398: -3000 crash, 50 exp(-5 d) + 300 (prev_d - d), +5000 discount^(steps/10), +1e6
            w.mode = dn::RW_HER; w.crash = -3000.f; w.final_bonus = 1.0e6f; w.capture_bonus = 5000.f;
            w.exp_w = 50.f; w.exp_k = 5.f; w.progress_w = 300.f;
            w.decay_log2 = (float)(std::log2(discount) / 10.0); return true;
        case DN_REWARD_REACHING:   // dummy_env.py:617-643 / Rewarder.py:8-40: +3 gate, 10 final, -10 collision, travel - distance
            w.mode = dn::RW_REACHING; w.crash = -10.f; w.final_bonus = 10.f; w.capture_bonus = 3.f; return true;
        case DN_REWARD_HOVER:      // HoverAviary.py:65-76: -|(0,0,1) - p|^2
            w.mode = dn::RW_POINT; w.pt_x = 0.f; w.pt_y_rate = 0.f; w.pt_z = 1.f; w.pt_w = 1.f; return true;
        case DN_REWARD_FLYTHRUGATE: // FlyThruGateAviary.py:100-112: -10 |(0, -2 t_norm, 0.75) - p|^2
            w.mode = dn::RW_POINT; w.pt_x = 0.f; w.pt_y_rate = -2.f; w.pt_z = 0.75f; w.pt_w = 10.f; return true;
        case DN_REWARD_PROGRESS:   // PBDroneEnv._computeReward with the progress term of Rewarder.py:43-62, weight 2000 (ThrustEnv.py:416-421)
            w = {-10.f, 200.f, 75.f, 5.f, 3.f, 2.f, 3000.f, 3.f, 0.7f, 0.3f, 1.f, 25.f, 0.04f};
            w.proj_w = 2000.f; return true;
        case DN_REWARD_BOOTSTRAPPED:   // Rewarder.py:66-104: lambda1..4 = 0.5, 0.025, 2e-4, 5e-4; c1 = 10, c2 = 4
            w.mode = dn::RW_LITERATURE; w.lit_prog = 0.5f; w.lit_perc_poly = (float)(0.025 * 2e-4); w.lit_da1 = -2e-4f; w.lit_w1 = -5e-4f;
            w.lit_pass = 10.f; w.lit_crash = 4.f; w.crash = -4.f; return true;
        case DN_REWARD_CHAMP:          // Rewarder.py:107-150: lambda1..5 = 1, 0.02, -10, -2e-4, -1e-4; c1 = 5 (also when p_z < 0)
            w.mode = dn::RW_LITERATURE; w.lit_prog = 1.f; w.lit_perc_exp_w = 0.02f; w.lit_perc_exp_k = -10.f; w.lit_w2 = -2e-4f; w.lit_da2 = -1e-4f;
            w.lit_crash = 5.f; w.lit_pz = 1; w.crash = -5.f; return true;
        default: return false;
    }
}

inline void fill_params(const dn_config& cfg, const RewardParams& rw, Params& P,
                        std::vector<float4>& h_t, std::vector<float4>& h_s, float& d0) {
    const int N = cfg.num_envs, T = cfg.num_targets;
    P.n = N;
    P.substeps = cfg.pyb_freq / cfg.ctrl_freq;
    P.act_type = cfg.act_type;
    P.normalize_actions = cfg.normalize_actions ? 1 : 0;
    P.physics = cfg.physics;
    P.obs_dim = cfg.include_distance ? 13 : 12;
    P.cylinder = cfg.cylinder ? 1 : 0;
    P.circle = cfg.circle ? 1 : 0;
    P.max_steps = cfg.max_steps;
    P.num_targets = T;
    P.spawn_mode = cfg.spawn_mode;
    P.reward_id = cfg.reward_id;
    P.dt = static_cast<float>(1.0 / cfg.pyb_freq);
    P.threshold = static_cast<float>(cfg.threshold);
    const double* ad = cfg.aviary_dim;
    P.x_low = (float)ad[0]; P.y_low = (float)ad[1]; P.z_low = (float)ad[2];
    P.x_high = (float)ad[3]; P.y_high = (float)ad[4]; P.z_high = (float)ad[5];
    const double mtd = std::fmax(std::fmax(std::fabs(ad[0]) + ad[3], std::fabs(ad[1]) + ad[4]), ad[5]);   // PBDroneEnv.py:91
    P.max_target_dist = static_cast<float>(mtd);
    P.inv_x_high = (float)(1.0 / ad[3]); P.inv_y_high = (float)(1.0 / ad[4]); P.inv_z_high = (float)(1.0 / ad[5]);
    P.inv_max_target_dist = (float)(1.0 / mtd);
    P.thr2 = (float)(cfg.threshold * cfg.threshold);
    P.cyl_limit2 = (float)((cfg.threshold + 0.2) * (cfg.threshold + 0.2));
    double q0[4];
    quat_from_euler(cfg.init_rpy, q0);
    for (int k = 0; k < 3; ++k) { P.init_pos[k] = (float)cfg.init_xyz[k]; P.init_seg_base[k] = (float)cfg.init_xyz[k]; }
    for (int k = 0; k < 4; ++k) P.init_quat[k] = (float)q0[k];
    {   // observation of the spawn pose, entries 0..11 (PBDroneEnv.py:338-398), in double
        double rpy[3];
        euler_from_quat(q0, rpy);
        const double pi = 3.14159265358979323846;
        double o[12] = {cfg.init_xyz[0] / ad[3], cfg.init_xyz[1] / ad[4], cfg.init_xyz[2] / ad[5],
                        std::fmin(std::fmax(rpy[0], -pi), pi) / pi, std::fmin(std::fmax(rpy[1], -pi), pi) / pi, rpy[2] / pi,
                        0, 0, 0, 0, 0, 0};
        for (int k = 0; k < 12; ++k) P.init_obs[k] = (float)o[k];
    }
    // action map constants: float32 like the reference (PBDroneEnv.py:113-116)
    const Airframe& A = *airframe(cfg.drone_model);
    const double a_low = A.KF * std::pow(Airframe::PWM2RPM_SCALE * Airframe::MIN_PWM + Airframe::PWM2RPM_CONST, 2);
    const double a_high = A.KF * std::pow(Airframe::PWM2RPM_SCALE * Airframe::MAX_PWM + Airframe::PWM2RPM_CONST, 2);
    P.a_low = (float)a_low; P.a_high = (float)a_high;
    P.kf = (float)A.KF; P.km = (float)(A.km_negated ? -A.KM : A.KM);   // RACE: -z_torques, exact in every rounding (:927-929)
    P.pwm_scale = (float)Airframe::PWM2RPM_SCALE; P.pwm_const = (float)Airframe::PWM2RPM_CONST;
    P.pwm_min = (float)Airframe::MIN_PWM; P.pwm_max = (float)Airframe::MAX_PWM;
    // divisors exactly as numpy forms them in float32, and their correctly rounded float32 reciprocals
    P.a_span = P.a_high - P.a_low;                       // float32 subtraction (PBDroneEnv.py:968)
    P.inv_a_span = (float)(1.0 / (double)P.a_span);
    P.inv_kf = (float)(1.0 / (double)P.kf);
    P.inv_pwm_scale = (float)(1.0 / (double)P.pwm_scale);
    const double gravity = Airframe::G * A.M;
    const double hover_rpm = std::sqrt(gravity / (4 * A.KF));
    const double max_rpm = std::sqrt((A.T2W * gravity) / (4 * A.KF));
    const double max_thrust = 4 * A.KF * max_rpm * max_rpm;
    P.hover_rpm = (float)hover_rpm;
    P.gravity = (float)gravity; P.inv_m = (float)(1.0 / A.M);
    P.frame_plus = A.plus ? 1 : 0;
    P.torque_arm = (float)(A.plus ? A.L : A.L / std::sqrt(2.0));
    P.ixx = (float)A.IXX; P.iyy = (float)A.IYY; P.izz = (float)A.IZZ;
    P.inv_ixx = (float)(1.0 / A.IXX); P.inv_iyy = (float)(1.0 / A.IYY); P.inv_izz = (float)(1.0 / A.IZZ);
    P.drag_xy = (float)Airframe::DRAG_XY; P.drag_z = (float)Airframe::DRAG_Z;
    P.gnd_coeff = (float)Airframe::GND_EFF_COEFF; P.prop_radius = (float)A.PROP_RADIUS;
    P.gnd_h_clip = (float)(0.25 * A.PROP_RADIUS * std::sqrt((15 * max_rpm * max_rpm * A.KF * Airframe::GND_EFF_COEFF) / max_thrust));
    P.collision_half_h = (float)(Airframe::COLLISION_H / 2);
    for (int k = 0; k < 4; ++k) { P.prop_x[k] = (float)A.prop_x[k]; P.prop_y[k] = (float)A.prop_y[k]; }
    // DSLPIDControl always reads the cf2x URDF (BaseSingleAgentAviary.py:72-73; BaseControl.py:33-37)
    {
        const Airframe& Cx = *airframe(DN_MODEL_CF2X);
        P.ctrl_dt = (float)(1.0 / cfg.ctrl_freq); P.inv_ctrl_dt = (float)cfg.ctrl_freq;
        P.speed_limit = (float)(0.03 * A.MAX_SPEED_KMH * (1000.0 / 3600.0));
        P.pid_gravity = (float)(Airframe::G * Cx.M);
        P.pid_inv_4kf = (float)(1.0 / (4.0 * Cx.KF));
    }
    P.rw = rw;
    P.rew_gamma = (float)(cfg.reward_gamma > 0.0 ? cfg.reward_gamma : 0.99);   // gym / normalize.NormalizeReward default
    P.rew_eps = 1e-8f;
    P.rew_clip = (float)(cfg.clip_reward > 0.0 ? cfg.clip_reward : 0.0);
    P.ep_time_scale = (float)((double)(cfg.pyb_freq / cfg.ctrl_freq) / ((double)cfg.pyb_freq * 1.0));   // EPISODE_LEN_SEC = 1: _clipAndNormalizeState overwrites the constructor's 5 (PBDroneEnv.py:68,348)
    P.seed = cfg.seed;
    P.env_id_offset = cfg.env_id_offset;

    // target table + segment table for the non-circle cylinder (PBDroneEnv.py:746-786), in double
    h_t.assign(T, float4{}); h_s.assign(2 * T, float4{});
    for (int k = 0; k < T; ++k) {
        const double* tk = cfg.targets + 3 * k;
        h_t[k] = make_float4((float)tk[0], (float)tk[1], (float)tk[2], 0.f);
        const double* b1 = (k == 0) ? cfg.init_xyz : cfg.targets + 3 * (k - 1);
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
    P.const_tables = (T <= kConstTargets) ? 1 : 0;
    if (P.const_tables) {
        for (int k = 0; k < T; ++k) { P.tgt_c[k] = h_t[k]; P.seg_c[2 * k] = h_s[2 * k]; P.seg_c[2 * k + 1] = h_s[2 * k + 1]; }
    }
    // constructor distance: ||INIT_XYZS[0] - target[0]|| (PBDroneEnv.py:137-138)
    {
        const double* t0 = cfg.targets;
        const double dx = cfg.init_xyz[0] - t0[0], dy = cfg.init_xyz[1] - t0[1], dz = cfg.init_xyz[2] - t0[2];
        d0 = (float)std::sqrt(dx * dx + dy * dy + dz * dz);
    }

}

}}  // namespace dn::host
