/* host_rollout.c — a plain-C host of the C ABI (include/psk_craft.h): no Python, no torch.
 *
 *   gcc -std=c99 -O2 -I include -I /usr/local/cuda/include examples/host_rollout.c \
 *       -L psketch_b200 -lpsk_b200 -L /usr/local/cuda/lib64 -lcudart -Wl,-rpath,$PWD/psketch_b200 -o /tmp/host_rollout
 *   /tmp/host_rollout [n_envs] [ticks]
 *
 * Builds the stock craft_medium tables by hand (the numbers of resources/craft/recipes.yaml and of
 * the get[wood] hint: go[wood], use), lays out one 8x8 scenario, runs teacher-driven rollouts with
 * psk_craft_rollout and prints the episode statistics.  Every episode of this scenario takes the
 * same number of ticks, so episodes == successes == n * (ticks / episode_length) — the check
 * tests/test_abi.py makes on the program's output (GPU tier).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <cuda_runtime_api.h>

#include "psk_craft.h"

#define CK(x)                                                                 \
    do {                                                                      \
        cudaError_t e_ = (x);                                                 \
        if (e_ != cudaSuccess) {                                              \
            fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_));          \
            return 2;                                                         \
        }                                                                     \
    } while (0)

enum { BOUNDARY = 1, WS0 = 2, WOOD = 9, PLANK = 12 };

int main(int argc, char **argv) {
    const long n = argc > 1 ? atol(argv[1]) : 4096;
    const int ticks = argc > 2 ? atoi(argv[2]) : 24;
    psk_craft_tables t;
    memset(&t, 0, sizeof t);
    t.width = t.height = 8;
    t.n_kinds = 21;
    t.window_w = t.window_h = 3;
    t.water_kind = 5; t.stone_kind = 6; t.bridge_kind = 19; t.axe_kind = 14;
    for (int k = 1; k <= 20; k++)          /* 1 boundary, 2-4 workshops, 5 water, 6 stone, 7.. grabbable */
        t.kind_class[k] = k == 1 ? 1 : k <= 4 ? 2 : k == 5 ? 3 : k == 6 ? 4 : 5;
    t.n_recipes = 1;                        /* plank <- wood at workshop0 (recipes.yaml:17-20) */
    { uint8_t r[8] = {PLANK, WS0, 1, WOOD, 1, 0, 0, 1}; memcpy(t.recipes[0], r, 8); }
    /* task 1 = get[wood]: nodes (sat class, arg, leaf kind, skip_to): get[wood], go[wood], use */
    t.n_tasks = 2;
    t.task_len[1] = 3;
    { uint8_t nd[3][4] = {{1, WOOD, 0, 3}, {2, WOOD, 2, 2}, {0, 0, 1, 3}}; memcpy(t.task_nodes[1], nd, sizeof nd); }
    if (!psk_craft_supported(&t) || psk_craft_n_features(&t) != 404) {
        fprintf(stderr, "tables not supported\n");
        return 2;
    }
    /* one scenario: boundary ring, wood at (5,4); every env starts at (2,4) facing DOWN */
    uint8_t scen[64];
    memset(scen, 0, sizeof scen);
    for (int i = 0; i < 8; i++) scen[i * 8] = scen[i * 8 + 7] = scen[i] = scen[56 + i] = BOUNDARY;
    scen[5 * 8 + 4] = WOOD;
    uint8_t *h_init = (uint8_t *)calloc((size_t)n, PSK_AGENT_BYTES);
    int32_t *h_idx = (int32_t *)calloc((size_t)n, sizeof(int32_t));
    for (long e = 0; e < n; e++) {
        uint8_t *a = h_init + e * PSK_AGENT_BYTES;
        a[PSK_AG_X] = 2; a[PSK_AG_Y] = 4; a[PSK_AG_DIR] = PSK_ACT_DOWN; a[PSK_AG_TASK] = 1; a[PSK_AG_TIMER] = 40;
    }
    uint8_t *d_scen, *d_init, *d_grid, *d_agent, *d_expert;
    int32_t *d_idx, *d_err;
    unsigned long long *d_stats;
    CK(cudaMalloc((void **)&d_scen, 64));
    CK(cudaMalloc((void **)&d_init, (size_t)n * PSK_AGENT_BYTES));
    CK(cudaMalloc((void **)&d_idx, (size_t)n * sizeof(int32_t)));
    CK(cudaMalloc((void **)&d_grid, (size_t)n * 64));
    CK(cudaMalloc((void **)&d_agent, (size_t)n * PSK_AGENT_BYTES));
    CK(cudaMalloc((void **)&d_expert, (size_t)n * ticks));
    CK(cudaMalloc((void **)&d_stats, 4 * sizeof(unsigned long long)));
    CK(cudaMalloc((void **)&d_err, sizeof(int32_t)));
    CK(cudaMemcpy(d_scen, scen, 64, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_init, h_init, (size_t)n * PSK_AGENT_BYTES, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_idx, h_idx, (size_t)n * sizeof(int32_t), cudaMemcpyHostToDevice));
    CK(cudaMemset(d_stats, 0, 4 * sizeof(unsigned long long)));
    CK(cudaMemset(d_err, 0, sizeof(int32_t)));
    psk_craft_state s = {d_grid, d_agent, n, 64, 0};
    psk_craft_episodes ep = {d_scen, d_idx, d_init};
    int rc = psk_craft_reset(s, ep, NULL, NULL);
    if (!rc) rc = psk_craft_rollout(&t, s, ep, ticks, NULL, NULL, 0, d_expert, NULL, NULL, d_stats, d_err, NULL);
    if (rc) {
        fprintf(stderr, "psk call failed: %d\n", rc);
        return 2;
    }
    CK(cudaDeviceSynchronize());
    unsigned long long st[4];
    int32_t err = 0;
    uint8_t first[8];
    CK(cudaMemcpy(st, d_stats, sizeof st, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&err, d_err, sizeof err, cudaMemcpyDeviceToHost));
    for (int k = 0; k < 8 && k < ticks; k++) CK(cudaMemcpy(first + k, d_expert + (size_t)k * n, 1, cudaMemcpyDeviceToHost));
    printf("%s envs=%ld ticks=%d episodes=%llu successes=%llu env_steps=%llu err=%d teacher:", psk_version(), n,
           ticks, st[0], st[1], st[2], (int)err);
    for (int k = 0; k < 8 && k < ticks; k++) printf(" %d", first[k]);
    printf("\n");
    return err ? 1 : 0;
}
