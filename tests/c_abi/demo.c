/* demo.c -- the C ABI of libdronenav.so driven from plain C99 (no CUDA headers, no C++, no Python):
 * what a non-Python host of the reference's env loop would link against.  Used by tests/test_c_abi_from_c.py.
 *
 *   demo <lib> symbols              dlopen + every entry point of include/dronenav.h + argument validation (no GPU needed)
 *   demo <lib> step <envs> <steps> <seed>  circle track (Waypoints.circle(1, 6, 1) minus its first point, PBDroneSimulator.py:111-130),
 *                                   240/30 Hz, hover-band actions from a fixed LCG, pageable host buffers through dn_step_host;
 *                                   prints one line per control step: reward sum, done count, found_targets sum, obs checksum
 */
#include <dlfcn.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "dronenav.h"

#define LOAD(name) do { *(void**)(&p_##name) = dlsym(h, #name); if (!p_##name) { fprintf(stderr, "missing symbol %s\n", #name); return 2; } } while (0)

static int (*p_dn_abi_version)(void);
static const char* (*p_dn_last_error)(void);
static int (*p_dn_create)(const dn_config*, int, dn_env**);
static int (*p_dn_destroy)(dn_env*);
static int (*p_dn_num_envs)(const dn_env*);
static int (*p_dn_obs_dim)(const dn_env*);
static int64_t (*p_dn_launch_count)(const dn_env*);
static int (*p_dn_reset)(dn_env*, const uint8_t*, float*, void*);
static int (*p_dn_step)(dn_env*, const dn_step_io*, void*);
static int (*p_dn_step_many)(dn_env*, const dn_step_io*, int, int, void*);
static int (*p_dn_step_host)(dn_env*, const dn_step_io*);
static int (*p_dn_action_to_rpm)(dn_env*, const float*, float*, int64_t, void*);
static int (*p_dn_get_state)(dn_env*, const dn_state_view*, void*);
static int (*p_dn_set_state)(dn_env*, const dn_state_view*, void*);
static int (*p_dn_episode_stats)(dn_env*, dn_stats*, int, void*);
static int (*p_dn_gae)(const float*, const float*, const uint8_t*, const float*, float, float, float*, float*, int32_t, int32_t, void*);

static void circle_config(dn_config* c, double* targets, int envs) {
    const double pi = 3.14159265358979323846;
    int k;
    memset(c, 0, sizeof(*c));
    for (k = 0; k < 6; ++k) {                       /* points 1..6 of the 7-point circle */
        targets[3 * k] = cos((k + 1) * pi / 3.0);
        targets[3 * k + 1] = sin((k + 1) * pi / 3.0);
        targets[3 * k + 2] = 1.0;
    }
    c->abi_version = DN_ABI_VERSION;
    c->num_envs = envs;
    c->pyb_freq = 240; c->ctrl_freq = 30;
    c->act_type = DN_ACT_THRUST; c->normalize_actions = 1;
    c->physics = DN_PHYS_DYN; c->reward_id = DN_REWARD_DEFAULT;
    c->include_distance = 1; c->cylinder = 1; c->circle = 1;
    c->max_steps = 4096; c->spawn_mode = DN_SPAWN_FIXED;
    c->threshold = 0.3; c->discount = 0.999;
    c->aviary_dim[0] = -2; c->aviary_dim[1] = -2; c->aviary_dim[2] = 0; c->aviary_dim[3] = 2; c->aviary_dim[4] = 2; c->aviary_dim[5] = 2;
    c->init_xyz[0] = 1; c->init_xyz[1] = 0; c->init_xyz[2] = 1;
    c->num_targets = 6; c->targets = targets;
    c->drone_model = DN_MODEL_CF2X;
}

int main(int argc, char** argv) {
    void* h;
    dn_config cfg;
    double targets[18];
    dn_env* env = NULL;
    if (argc < 3) { fprintf(stderr, "usage: demo <lib> symbols | step <envs> <steps> <seed>\n"); return 1; }
    h = dlopen(argv[1], RTLD_NOW);
    if (!h) { fprintf(stderr, "dlopen: %s\n", dlerror()); return 2; }
    LOAD(dn_abi_version); LOAD(dn_last_error); LOAD(dn_create); LOAD(dn_destroy); LOAD(dn_num_envs); LOAD(dn_obs_dim);
    LOAD(dn_launch_count); LOAD(dn_reset); LOAD(dn_step); LOAD(dn_step_many); LOAD(dn_step_host); LOAD(dn_action_to_rpm);
    LOAD(dn_get_state); LOAD(dn_set_state); LOAD(dn_episode_stats); LOAD(dn_gae);
    if (p_dn_abi_version() != DN_ABI_VERSION) { fprintf(stderr, "ABI version %d != header %d\n", p_dn_abi_version(), DN_ABI_VERSION); return 3; }

    if (strcmp(argv[2], "symbols") == 0) {
        circle_config(&cfg, targets, 4);
        cfg.ctrl_freq = 7;                                           /* 240 % 7 != 0 */
        if (p_dn_create(&cfg, 0, &env) != DN_EINVAL || !strstr(p_dn_last_error(), "divisible")) return 4;
        circle_config(&cfg, targets, 4);
        cfg.drone_model = DN_MODEL_RACE; cfg.act_type = DN_ACT_VEL;  /* no controller for the racer */
        if (p_dn_create(&cfg, 0, &env) != DN_EINVAL || !strstr(p_dn_last_error(), "no controller")) return 5;
        if (p_dn_step_host(NULL, NULL) != DN_EINVAL) return 6;
        printf("abi %d sizeof(dn_config) %u sizeof(dn_step_io) %u sizeof(dn_state_view) %u sizeof(dn_stats) %u\n", p_dn_abi_version(),
               (unsigned)sizeof(dn_config), (unsigned)sizeof(dn_step_io), (unsigned)sizeof(dn_state_view), (unsigned)sizeof(dn_stats));
        return 0;
    }

    if (strcmp(argv[2], "step") == 0 && argc >= 6) {
        const int N = atoi(argv[3]), T = atoi(argv[4]);
        int D, t, i, k;
        unsigned int lcg = (unsigned int)atoi(argv[5]);
        float *actions, *obs, *reward;
        uint8_t* done;
        int32_t* found;
        dn_step_io io;
        dn_stats st;
        circle_config(&cfg, targets, N);
        if (p_dn_create(&cfg, 0, &env) != DN_OK) { fprintf(stderr, "dn_create: %s\n", p_dn_last_error()); return 7; }
        D = p_dn_obs_dim(env);
        actions = (float*)malloc(sizeof(float) * 4 * N); obs = (float*)malloc(sizeof(float) * D * N);
        reward = (float*)malloc(sizeof(float) * N); done = (uint8_t*)malloc(N); found = (int32_t*)malloc(sizeof(int32_t) * N);
        memset(&io, 0, sizeof(io));
        io.actions = actions; io.obs = obs; io.reward = reward; io.done = done; io.found_targets = found;
        for (t = 0; t < T; ++t) {
            double rsum = 0.0, osum = 0.0;
            long dsum = 0, fsum = 0;
            for (i = 0; i < 4 * N; ++i) {                           /* hover band: 0.092227 + 0.004 u, u in [-1, 1) from a 31-bit LCG */
                lcg = (1103515245u * lcg + 12345u) & 0x7fffffffu;
                actions[i] = (float)(0.092227 + 0.004 * ((double)lcg / 1073741824.0 - 1.0));
            }
            if (p_dn_step_host(env, &io) != DN_OK) { fprintf(stderr, "dn_step_host: %s\n", p_dn_last_error()); return 8; }
            for (i = 0; i < N; ++i) {
                rsum += reward[i]; dsum += done[i] != 0; fsum += found[i];
                for (k = 0; k < D; ++k) osum += obs[i * D + k];
            }
            printf("%d %.6f %ld %ld %.5f\n", t, rsum, dsum, fsum, osum);
        }
        if (p_dn_episode_stats(env, &st, 0, NULL) != DN_OK) return 9;
        printf("episodes %llu crashes %llu launches %lld\n", (unsigned long long)st.episodes, (unsigned long long)st.crashes,
               (long long)p_dn_launch_count(env));
        p_dn_destroy(env);
        free(actions); free(obs); free(reward); free(done); free(found);
        return 0;
    }
    return 1;
}
