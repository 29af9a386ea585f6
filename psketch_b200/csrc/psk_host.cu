// psk_host.cu — host-buffer entry points of the C ABI (include/psk_craft.h, "host_" section).
//
// A caller that keeps its environments in host memory (as the reference does: every CraftState
// is a numpy object on the CPU) hands whole batches to psk_craft_host_tick.  The batch is cut
// into chunks; chunk i runs on stream i % PSK_HOST_STREAMS as  H2D(state) -> fused tick kernel
// -> D2H(features, teacher actions, flags, new state), so the copy engines and the SMs overlap
// across chunks.  Pass pinned memory (psk_host_alloc, or any cudaHostRegister'ed / torch-pinned
// buffer) or the copies degrade to staged synchronous ones.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

#include <cuda_runtime.h>

#include "../../include/psk_craft.h"

#define PSK_HOST_STREAMS 3

struct psk_craft_host_ctx {
    psk_craft_tables tables;
    int64_t max_envs, chunk;
    int cell_stride, nf;
    int device;
    cudaStream_t streams[PSK_HOST_STREAMS];
    // per-stream chunk buffers
    uint8_t *d_grid[PSK_HOST_STREAMS], *d_agent[PSK_HOST_STREAMS], *d_action[PSK_HOST_STREAMS];
    uint8_t *d_expert[PSK_HOST_STREAMS], *d_done[PSK_HOST_STREAMS], *d_success[PSK_HOST_STREAMS];
    float *d_feat[PSK_HOST_STREAMS];
    // device-resident episode tables (static per context)
    uint8_t *d_scen_grid, *d_init_agent;
    int32_t *d_scen_idx;
    int64_t n_scen, n_eps;
    unsigned long long *d_stats;
    int32_t *d_err;
};

#define CK(x)                                   \
    do {                                        \
        if ((x) != cudaSuccess) return PSK_ERR_CUDA; \
    } while (0)

// A context belongs to the device that was current when it was created; calls made while another
// device is current switch to it for their duration.
struct DeviceScope {
    int prev = -1;
    explicit DeviceScope(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != dev) cudaSetDevice(dev); else prev = -1;
    }
    ~DeviceScope() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

extern "C" {

void *psk_host_alloc(size_t bytes) {
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) return nullptr;
    return p;
}

void psk_host_free(void *p) {
    if (p) cudaFreeHost(p);
}

static int host_ctx_alloc(psk_craft_host_ctx *c) {
    CK(cudaGetDevice(&c->device));
    for (int i = 0; i < PSK_HOST_STREAMS; i++) {
        CK(cudaStreamCreateWithFlags(&c->streams[i], cudaStreamNonBlocking));
        CK(cudaMalloc(&c->d_grid[i], (size_t)c->chunk * c->cell_stride));
        CK(cudaMalloc(&c->d_agent[i], (size_t)c->chunk * PSK_AGENT_BYTES));
        CK(cudaMalloc(&c->d_action[i], (size_t)c->chunk));
        CK(cudaMalloc(&c->d_expert[i], (size_t)c->chunk));
        CK(cudaMalloc(&c->d_done[i], (size_t)c->chunk));
        CK(cudaMalloc(&c->d_success[i], (size_t)c->chunk));
        CK(cudaMalloc(&c->d_feat[i], (size_t)c->chunk * c->nf * sizeof(float)));
    }
    CK(cudaMalloc(&c->d_stats, 4 * sizeof(unsigned long long)));
    CK(cudaMemset(c->d_stats, 0, 4 * sizeof(unsigned long long)));
    CK(cudaMalloc(&c->d_err, sizeof(int32_t)));
    CK(cudaMemset(c->d_err, 0, sizeof(int32_t)));
    return PSK_OK;
}

void psk_craft_host_destroy(psk_craft_host_ctx *c);

int psk_craft_host_create(const psk_craft_tables *t, int64_t max_envs, int64_t chunk_envs,
                          psk_craft_host_ctx **out) {
    if (!t || !out || max_envs <= 0) return PSK_ERR_BADARG;
    if (!psk_craft_supported(t)) return PSK_ERR_UNSUPPORTED;
    psk_craft_host_ctx *c = new (std::nothrow) psk_craft_host_ctx();
    if (!c) return PSK_ERR_BADARG;
    memset(c, 0, sizeof(*c));
    c->tables = *t;
    c->max_envs = max_envs;
    if (chunk_envs <= 0) chunk_envs = 16384;
    if (chunk_envs > max_envs) chunk_envs = max_envs;
    c->chunk = (chunk_envs + 127) / 128 * 128;
    c->cell_stride = ((t->width * t->height + 63) / 64) * 64;
    c->nf = psk_craft_n_features(t);
    const int rc = host_ctx_alloc(c);
    if (rc != PSK_OK) {             // release whatever was allocated before the failure
        psk_craft_host_destroy(c);
        return rc;
    }
    *out = c;
    return PSK_OK;
}

void psk_craft_host_destroy(psk_craft_host_ctx *c) {
    if (!c) return;
    DeviceScope scope(c->device);
    for (int i = 0; i < PSK_HOST_STREAMS; i++) {
        if (c->streams[i]) cudaStreamSynchronize(c->streams[i]);
        cudaFree(c->d_grid[i]); cudaFree(c->d_agent[i]); cudaFree(c->d_action[i]);
        cudaFree(c->d_expert[i]); cudaFree(c->d_done[i]); cudaFree(c->d_success[i]);
        cudaFree(c->d_feat[i]);
        if (c->streams[i]) cudaStreamDestroy(c->streams[i]);
    }
    cudaFree(c->d_scen_grid); cudaFree(c->d_init_agent); cudaFree(c->d_scen_idx);
    cudaFree(c->d_stats); cudaFree(c->d_err);
    delete c;
}

int psk_craft_host_set_episodes(psk_craft_host_ctx *c, const uint8_t *host_scen_grid,
                                int64_t n_scen, const int32_t *host_scen_idx,
                                const uint8_t *host_init_agent, int64_t n) {
    if (!c || !host_scen_grid || !host_scen_idx || !host_init_agent || n <= 0 || n_scen <= 0 ||
        n > c->max_envs)
        return PSK_ERR_BADARG;
    DeviceScope scope(c->device);
    cudaFree(c->d_scen_grid); cudaFree(c->d_init_agent); cudaFree(c->d_scen_idx);
    c->d_scen_grid = nullptr; c->d_init_agent = nullptr; c->d_scen_idx = nullptr;
    CK(cudaMalloc(&c->d_scen_grid, (size_t)n_scen * c->cell_stride));
    CK(cudaMalloc(&c->d_scen_idx, (size_t)n * sizeof(int32_t)));
    CK(cudaMalloc(&c->d_init_agent, (size_t)n * PSK_AGENT_BYTES));
    CK(cudaMemcpy(c->d_scen_grid, host_scen_grid, (size_t)n_scen * c->cell_stride, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(c->d_scen_idx, host_scen_idx, (size_t)n * sizeof(int32_t), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(c->d_init_agent, host_init_agent, (size_t)n * PSK_AGENT_BYTES, cudaMemcpyHostToDevice));
    c->n_scen = n_scen;
    c->n_eps = n;
    return PSK_OK;
}

int psk_craft_host_tick(psk_craft_host_ctx *c, uint8_t *host_grid, uint8_t *host_agent,
                        const uint8_t *host_action_in, float *host_features,
                        uint8_t *host_expert, uint8_t *host_done, uint8_t *host_success,
                        int64_t n, unsigned long long *host_stats, int32_t *host_err_flags) {
    if (!c || !host_grid || !host_agent || !host_expert || n < 0 || n > c->n_eps)
        return PSK_ERR_BADARG;
    DeviceScope scope(c->device);
    const int cs = c->cell_stride;
    int k = 0;
    for (int64_t off = 0; off < n; off += c->chunk, k++) {
        const int s = k % PSK_HOST_STREAMS;
        const int64_t m = (n - off) < c->chunk ? (n - off) : c->chunk;
        cudaStream_t st = c->streams[s];
        CK(cudaMemcpyAsync(c->d_grid[s], host_grid + off * cs, (size_t)m * cs, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(c->d_agent[s], host_agent + off * PSK_AGENT_BYTES, (size_t)m * PSK_AGENT_BYTES,
                           cudaMemcpyHostToDevice, st));
        if (host_action_in)
            CK(cudaMemcpyAsync(c->d_action[s], host_action_in + off, (size_t)m, cudaMemcpyHostToDevice, st));
        psk_craft_state state = {c->d_grid[s], c->d_agent[s], m, cs, 0};
        psk_craft_episodes ep = {c->d_scen_grid, c->d_scen_idx + off, c->d_init_agent + off * PSK_AGENT_BYTES};
        int rc = psk_craft_tick(&c->tables, state, ep, host_action_in ? c->d_action[s] : nullptr,
                                host_features ? c->d_feat[s] : nullptr, c->d_expert[s],
                                c->d_done[s], c->d_success[s], c->d_stats, c->d_err, 1, st);
        if (rc) return rc;
        if (host_features)
            CK(cudaMemcpyAsync(host_features + off * c->nf, c->d_feat[s], (size_t)m * c->nf * sizeof(float),
                               cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(host_expert + off, c->d_expert[s], (size_t)m, cudaMemcpyDeviceToHost, st));
        if (host_done) CK(cudaMemcpyAsync(host_done + off, c->d_done[s], (size_t)m, cudaMemcpyDeviceToHost, st));
        if (host_success)
            CK(cudaMemcpyAsync(host_success + off, c->d_success[s], (size_t)m, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(host_grid + off * cs, c->d_grid[s], (size_t)m * cs, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(host_agent + off * PSK_AGENT_BYTES, c->d_agent[s], (size_t)m * PSK_AGENT_BYTES,
                           cudaMemcpyDeviceToHost, st));
    }
    for (int i = 0; i < PSK_HOST_STREAMS; i++) CK(cudaStreamSynchronize(c->streams[i]));
    if (host_stats) CK(cudaMemcpy(host_stats, c->d_stats, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    if (host_err_flags) {
        CK(cudaMemcpy(host_err_flags, c->d_err, sizeof(int32_t), cudaMemcpyDeviceToHost));
        if (*host_err_flags) CK(cudaMemset(c->d_err, 0, sizeof(int32_t)));
    }
    return PSK_OK;
}

}  // extern "C"
