// emu.cpp -- TEST TOOL: host build of the device step logic (csrc/dn_device.cuh) behind a
// tiny C interface.  Used by tests/test_host_emulation.py to check the kernel's logic
// against the oracle on machines without a GPU.  It is FP32 like the kernel (FMA
// contraction may differ from nvcc's), it is NOT a CPU fallback of the product and is
// never imported by the package.
#include "cuda_shim.h"
#include "../../drl-dronenavigation_b200/csrc/dn_device.cuh"
#include "../../drl-dronenavigation_b200/csrc/dn_host.h"
#include <vector>

struct Emu {
    dn::Params P;
    std::vector<float4> planes[dn::kPlanes];
    std::vector<float4> targets, segs;
    std::vector<float> last_rpm_sum, obs_rms;
    std::vector<float4> aux, rew_rms, spawn, pid[3];
    int normalize_obs;
    float d0;
};

template <int PHYS>
static void step_all(Emu* e, const float* actions, float* obs, float* reward, uint8_t* done, float* term_obs, int32_t* found) {
    const dn::Params& P = e->P;
    for (int i = 0; i < P.n; ++i) {
        dn::EnvState s;
        dn::load_core(P, i, s);
        float lrs = (PHYS & 1) ? P.last_rpm_sum[i] : 0.f;
        const float4 a = make_float4(actions[4 * i], actions[4 * i + 1], actions[4 * i + 2], actions[4 * i + 3]);
        float* row = obs + (size_t)i * P.obs_dim;
        dn::StepResult r = dn::env_step<PHYS>(P, i, s, a, lrs, row);
        reward[i] = r.reward; done[i] = r.done; found[i] = r.found;
        if (r.finished) for (int k = 0; k < P.obs_dim; ++k) {
            term_obs[(size_t)i * P.obs_dim + k] = row[k];
            row[k] = (k < 3) ? r.spawn_obs[k] : ((k < 12) ? P.init_obs[k] : r.reset_obs_dist);
        }
        dn::store_state(P, i, s);
        if (PHYS & 1) P.last_rpm_sum[i] = lrs;
    }
}

extern "C" {

Emu* emu_create(const dn_config* cfg) {
    dn::RewardParams rw;
    if (!dn::host::reward_table(cfg->reward_id, cfg->discount, rw)) return nullptr;
    if (!dn::host::airframe(cfg->drone_model)) return nullptr;
    Emu* e = new Emu();
    memset(&e->P, 0, sizeof(e->P));
    dn::host::fill_params(*cfg, rw, e->P, e->targets, e->segs, e->d0);
    const int N = cfg->num_envs;
    e->normalize_obs = cfg->normalize_obs;
    for (int k = 0; k < dn::kPlanes; ++k) { e->planes[k].assign(N, float4{0, 0, 0, 0}); e->P.s[k] = e->planes[k].data(); }
    e->P.targets = e->targets.data(); e->P.segs = e->segs.data();
    if (cfg->physics & DN_PHYS_DRAG) { e->last_rpm_sum.assign(N, 0.f); e->P.last_rpm_sum = e->last_rpm_sum.data(); }
    if (rw.mode == dn::RW_REACHING) {
        e->aux.assign(N, make_float4(e->P.init_pos[0], e->P.init_pos[1], e->P.init_pos[2], 0.f)); e->P.aux = e->aux.data();
    }
    if (rw.mode == dn::RW_LITERATURE) { e->aux.assign(N, make_float4(0.f, 0.f, 0.f, 0.f)); e->P.aux = e->aux.data(); }
    if (cfg->spawn_mode != DN_SPAWN_FIXED) {
        e->spawn.assign(N, make_float4(e->P.init_pos[0], e->P.init_pos[1], e->P.init_pos[2], 0.f)); e->P.spawn = e->spawn.data();
    }
    if (cfg->act_type >= DN_ACT_PID) {
        for (int k = 0; k < 3; ++k) { e->pid[k].assign(N, float4{0, 0, 0, 0}); e->P.pid[k] = e->pid[k].data(); }
    }
    if (cfg->normalize_reward) { e->rew_rms.assign(N, make_float4(0.f, 0.f, 1.f, 1e-4f)); e->P.rew_rms = e->rew_rms.data(); }
    for (int i = 0; i < N; ++i) {
        dn::EnvState s;
        memset(&s, 0, sizeof(s));
        s.px = e->P.init_pos[0]; s.py = e->P.init_pos[1]; s.pz = e->P.init_pos[2]; s.dist = s.prev_dist = e->d0;
        s.qx = e->P.init_quat[0]; s.qy = e->P.init_quat[1]; s.qz = e->P.init_quat[2]; s.qw = e->P.init_quat[3];
        dn::store_state(e->P, i, s);
    }
    return e;
}

void emu_destroy(Emu* e) { delete e; }

void emu_step(Emu* e, const float* actions, float* obs, float* reward, uint8_t* done, float* term_obs, int32_t* found) {
    switch (e->P.physics & 3) {
        case 0: step_all<0>(e, actions, obs, reward, done, term_obs, found); break;
        case 1: step_all<1>(e, actions, obs, reward, done, term_obs, found); break;
        case 2: step_all<2>(e, actions, obs, reward, done, term_obs, found); break;
        default: step_all<3>(e, actions, obs, reward, done, term_obs, found); break;
    }
}

void emu_action_to_rpm(Emu* e, const float* a, float* out, long long n) {
    for (long long i = 0; i < n; ++i) out[i] = dn::action_to_rpm(e->P, a[i]);
}

// raw access to the packed planes: [7][N][4] floats
void emu_get_planes(Emu* e, float* out) {
    for (int k = 0; k < dn::kPlanes; ++k) memcpy(out + (size_t)k * e->P.n * 4, e->planes[k].data(), (size_t)e->P.n * 16);
}
void emu_set_planes(Emu* e, const float* in) {
    for (int k = 0; k < dn::kPlanes; ++k) memcpy(e->planes[k].data(), in + (size_t)k * e->P.n * 4, (size_t)e->P.n * 16);
}
void emu_get_pid(Emu* e, float* out) {     // [N,9]
    if (!e->P.pid[0]) return;
    for (int i = 0; i < e->P.n; ++i) for (int k = 0; k < 3; ++k) {
        const float4 v = e->pid[k][i]; out[9 * i + 3 * k] = v.x; out[9 * i + 3 * k + 1] = v.y; out[9 * i + 3 * k + 2] = v.z;
    }
}
void emu_set_pid(Emu* e, const float* in) {
    if (!e->P.pid[0]) return;
    for (int i = 0; i < e->P.n; ++i) for (int k = 0; k < 3; ++k)
        e->pid[k][i] = make_float4(in[9 * i + 3 * k], in[9 * i + 3 * k + 1], in[9 * i + 3 * k + 2], 0.f);
}
void emu_set_last_rpm_sum(Emu* e, const float* in) { if (e->P.last_rpm_sum) memcpy(e->P.last_rpm_sum, in, (size_t)e->P.n * 4); }

}  // extern "C"
