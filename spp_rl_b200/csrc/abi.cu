// C ABI of spp_rl_b200 (see include/spp_rl_b200.h): handle, device memory, parameter I/O, replay ring
// host state machine, and the launchers of the fused update kernels.
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <string>
#include <vector>

#include "../../include/spp_rl_b200.h"
#include "layout.h"
#include "ring_kernels.h"
#include "update_kernel.cuh"

namespace spp {
cudaError_t launch_update_burst(const UpdateArgs& a, int grid, cudaStream_t stream);
cudaError_t launch_update_burst_twoslot(const UpdateArgs& a, int grid, cudaStream_t stream);      // update_kernel_twoslot.cu
cudaError_t launch_acm_train(const UpdateArgs& a, int grid, cudaStream_t stream);
cudaError_t launch_rollout(const RolloutArgs& r, int grid, cudaStream_t stream);
}

using namespace spp;

static thread_local std::string g_err;
static std::atomic<int64_t> g_launches{0};

static int fail(int code, const std::string& msg) { g_err = msg; return code; }
int spp_set_error_(int code, const std::string& msg) { return fail(code, msg); }     // shared with ppo_abi.cu
void spp_count_launch_() { g_launches++; }
#define CK(expr)                                                                                      \
    do {                                                                                              \
        cudaError_t e_ = (expr);                                                                      \
        if (e_ != cudaSuccess)                                                                        \
            return fail(SPP_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_));           \
    } while (0)

// Host -> device writes of ring rows, statistics and parameters go through the population's own (non-blocking) stream, so they
// are ordered against in-flight work of the asynchronous entry points; pageable sources are staged before the call returns.
#define H2D(dst, src, bytes) CK(cudaMemcpyAsync((dst), (src), (bytes), cudaMemcpyHostToDevice, p->stream))
#define H2D_DONE() CK(cudaStreamSynchronize(p->stream))

struct TensorMap {
    std::string name;
    int layer;      // index into NetDesc::L
    int is_bias;    // 0: rows of W, 1: bias vector, 2: gain row of the BasicAcM pseudo layer
    int row0;       // first row inside the layer (heads: fc_prob rows [0,ob), fc_scale rows [ob,2ob))
    int rows, cols; // reference shape
    int col0;       // gains: first column inside the pseudo layer
};

struct DevBuf {
    void* p = nullptr; size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e == cudaSuccess) cap = bytes;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

struct spp_population {
    spp_config cfg;
    int P = 0, device = 0, sm_count = 0;
    Layout L;
    Layout L_acm;               // same nets, scratch sized for acm_batch_size (ACM regression bursts)
    Hyper h;
    float* scratch_acm = nullptr;
    float *params = nullptr, *mom_m = nullptr, *mom_v = nullptr, *scratch = nullptr, *norm = nullptr, *acm_lim = nullptr;
    int* steps = nullptr;
    double* alpha_state = nullptr;
    // ring
    int64_t S = 0;
    float *r_obs = nullptr, *r_act = nullptr, *r_rew = nullptr, *r_aacm = nullptr;
    int32_t *r_oidx = nullptr, *r_nidx = nullptr;
    uint8_t *r_done = nullptr, *r_end = nullptr;
    int64_t* r_len = nullptr;    // device copy of current_len per agent
    DevBuf d_stat_partial, d_stat_moments, d_stat_state, d_stat_hist;      // spp_ring_obs_stats scratch
    std::vector<int64_t> obs_cur, ts_cur, cur_len;
    bool len_dirty = true;
    // staging
    DevBuf d_obs, d_nobs, d_act, d_rew, d_done, d_aacm, d_eps, d_idx, d_losses, d_tmp;
    DevBuf d_roll_scratch, d_env, d_cur;     // rollout: scratch for E rows per agent, synthetic env state, cursors
    DevBuf d_progress;                       // update bursts of populations larger than the grid: per-agent step counters (update_burst_kernel)
    DevBuf d_add_flags, d_add_src[4];        // spp_ring_add_rollout_store staging (kept: cudaMalloc / cudaFree cost 100+ ms next to a busy torch allocator)
    int roll_E = 0;
    Layout L_roll;
    cudaStream_t stream = nullptr;
    uint64_t seq = 0;
    std::vector<TensorMap> tensors[NET_COUNT];
};

static const NetDesc& net_desc(const spp_population* p, int net) {
    switch (net) {
        case NET_ACTOR: case NET_ACTOR_TARG: return p->L.actor;
        case NET_ACM: return p->L.acm;
        default: return p->L.critic;
    }
}

static void build_tensor_maps(spp_population* p) {
    const Layout& L = p->L;
    auto lin = [](std::vector<TensorMap>& v, const char* nm, int layer, int row0, int rows, int cols) {
        v.push_back({std::string(nm) + ".weight", layer, 0, row0, rows, cols, 0});
        v.push_back({std::string(nm) + ".bias", layer, 1, row0, rows, 1, 0});
    };
    for (int net : {NET_ACTOR, NET_ACTOR_TARG}) {
        auto& v = p->tensors[net];
        lin(v, "fc1", 0, 0, kHidden, L.ob);
        lin(v, "fc2", 1, 0, kHidden, kHidden);
        if (L.algo == ALGO_SAC) {
            lin(v, "fc_prob", 2, 0, L.ob, kHidden);
            lin(v, "fc_scale", 2, L.ob, L.ob, kHidden);
        } else {
            lin(v, "fc3", 2, 0, L.ob, kHidden);
        }
    }
    for (int net : {NET_CRITIC_1, NET_CRITIC_2, NET_CRITIC_1_TARG, NET_CRITIC_2_TARG}) {
        auto& v = p->tensors[net];
        lin(v, "fc1", 0, 0, kHidden, L.ob + L.act_dim);
        lin(v, "fc2", 1, 0, kHidden, kHidden);
        lin(v, "fc3", 2, 0, 1, kHidden);
    }
    {
        auto& v = p->tensors[NET_ACM];
        if (L.acm_kind == ACM_BASIC) {   // state_dict order of BasicAcM: t, t1, fc1, fc2, fc21, fc3
            v.push_back({"t", 4, 2, 0, 1, 1, 0});
            v.push_back({"t1", 4, 2, 0, L.ac, 1, 4});
            lin(v, "fc1", 0, 0, L.hm1, 2 * L.ob);
            lin(v, "fc2", 1, 0, L.hm2, L.hm1);
            lin(v, "fc21", 3, 0, L.hm2, 2 * L.ob);
            lin(v, "fc3", 2, 0, L.ac, L.hm2);
        } else {
            lin(v, "fc1", 0, 0, L.hm1, 2 * L.ob);
            lin(v, "fc2", 1, 0, L.hm2, L.hm1);
            lin(v, "fc3", 2, 0, L.ac, L.hm2);
        }
    }
}

static inline int map_col(const LayerDesc& l, int c) { return (l.split > 0 && c >= l.split) ? pad4(l.split) + (c - l.split) : c; }

// ------------------------------------------------------------------------------------------------
#include "ppo_rollout.h"
int spp_population_acm_view_(spp_population* p, int agent, spp::PopulationAcmView* out) {
    if (!p || !out || agent < 0 || agent >= p->P) return fail(SPP_ERR_ARG, "bad population / agent");
    const Layout& L = p->L;
    out->device = p->device; out->ob = L.ob; out->ac = L.ac; out->lda = L.lda; out->ldo = L.ldo; out->acm_kind = L.acm_kind;
    out->hm1 = L.hm1; out->hm2 = L.hm2; out->ldm1 = L.ldm1; out->ldm2 = L.ldm2; out->acm_desc = L.acm;
    out->acm = p->params + (size_t)agent * L.params_size + L.net_off[NET_ACM];
    out->acm_lim = p->acm_lim;
    return SPP_OK;
}

extern "C" {

int spp_abi_version(void) { return SPP_ABI_VERSION; }
const char* spp_last_error(void) { return g_err.c_str(); }
int64_t spp_kernel_launches(void) { return g_launches.load(); }

int spp_device_info(int device, int* sm_count, int* cc_major, int* cc_minor, char* name, int name_cap) {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (cc_major) *cc_major = prop.major;
    if (cc_minor) *cc_minor = prop.minor;
    if (name && name_cap > 0) { strncpy(name, prop.name, name_cap - 1); name[name_cap - 1] = 0; }
    return SPP_OK;
}

int spp_population_destroy(spp_population* p) {
    if (!p) return SPP_OK;
    cudaSetDevice(p->device);
    if (p->stream) cudaStreamSynchronize(p->stream);
    for (void* q : {(void*)p->params, (void*)p->mom_m, (void*)p->mom_v, (void*)p->scratch, (void*)p->norm, (void*)p->acm_lim,
                    (void*)p->steps, (void*)p->alpha_state, (void*)p->r_obs, (void*)p->r_act, (void*)p->r_rew, (void*)p->r_aacm,
                    (void*)p->r_oidx, (void*)p->r_nidx, (void*)p->r_done, (void*)p->r_end, (void*)p->r_len, (void*)p->scratch_acm})
        if (q) cudaFree(q);
    for (DevBuf* b : {&p->d_obs, &p->d_nobs, &p->d_act, &p->d_rew, &p->d_done, &p->d_aacm, &p->d_eps, &p->d_idx, &p->d_losses, &p->d_tmp, &p->d_roll_scratch, &p->d_env, &p->d_cur,
                      &p->d_stat_partial, &p->d_stat_moments, &p->d_stat_state, &p->d_stat_hist, &p->d_progress, &p->d_add_flags, &p->d_add_src[0], &p->d_add_src[1],
                      &p->d_add_src[2], &p->d_add_src[3]})
        b->release();
    if (p->stream) cudaStreamDestroy(p->stream);
    delete p;
    return SPP_OK;
}

int spp_population_create(const spp_config* cfg, int population, int device, spp_population** out) {
    if (!cfg || !out || population < 1) return fail(SPP_ERR_ARG, "spp_population_create: null argument or population < 1");
    if (cfg->ob_dim < 1 || cfg->ob_dim > 128 || cfg->ac_dim < 1 || cfg->ac_dim > 32)
        return fail(SPP_ERR_ARG, "spp_population_create: ob_dim must be in [1,128], ac_dim in [1,32]");
    if (cfg->update_batch_size < 1 || cfg->update_batch_size > 4096) return fail(SPP_ERR_ARG, "update_batch_size must be in [1,4096]");
    if (cfg->acm_batch_size < 0 || cfg->acm_batch_size > 4096) return fail(SPP_ERR_ARG, "acm_batch_size must be in [0,4096]");
    if (cfg->algo != SPP_ALGO_SAC && cfg->algo != SPP_ALGO_DDPG) return fail(SPP_ERR_ARG, "unknown algo");
    if (cfg->acm_kind != SPP_ACM_MLP && cfg->acm_kind != SPP_ACM_BASIC) return fail(SPP_ERR_ARG, "unknown acm_kind");
    if (!cfg->acm_critic && !cfg->store_actions && cfg->buffer_size > 0)
        return fail(SPP_ERR_ARG, "acm_critic=0 needs store_actions=1 (the critic reads the state-target action)");
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return fail(SPP_ERR_ARG, "no such CUDA device");
    CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) return fail(SPP_ERR_UNSUPPORTED, "spp_rl_b200 is built for sm_100a (B200) only");

    spp_population* p = new spp_population();
    p->cfg = *cfg; p->P = population; p->device = device; p->sm_count = prop.multiProcessorCount;
    p->L = make_layout(cfg->algo, cfg->ob_dim, cfg->ac_dim, cfg->acm_kind, cfg->acm_critic ? 1 : 0, cfg->update_batch_size);
    Hyper& h = p->h;
    h.gamma = (float)cfg->gamma; h.tau = (float)cfg->tau; h.one_minus_tau = (float)(1.0 - cfg->tau);
    h.custom_loss = (float)cfg->custom_loss; h.target_entropy = (float)cfg->target_entropy;
    h.actor_lr = cfg->actor_lr; h.critic_lr = cfg->critic_lr; h.alpha_lr = cfg->alpha_lr; h.acm_lr = cfg->acm_lr;
    p->L_acm = make_layout(cfg->algo, cfg->ob_dim, cfg->ac_dim, cfg->acm_kind, cfg->acm_critic ? 1 : 0,
                           cfg->acm_batch_size > 0 ? cfg->acm_batch_size : 128);
    h.norm_closs = cfg->norm_closs ? 1 : 0; h.norm_clamp = cfg->min_max_denormalize ? 0 : 1;
    build_tensor_maps(p);
    const Layout& L = p->L;
    const size_t P = (size_t)population;
    auto alloc = [&](void** q, size_t bytes) -> cudaError_t {
        cudaError_t e = cudaMalloc(q, bytes ? bytes : 16);
        if (e == cudaSuccess) e = cudaMemset(*q, 0, bytes ? bytes : 16);
        return e;
    };
#define ALLOC(ptr, bytes)                                                                          \
    do {                                                                                           \
        cudaError_t e_ = alloc((void**)&(ptr), (bytes));                                           \
        if (e_ != cudaSuccess) {                                                                   \
            std::string m = std::string("cudaMalloc(" #ptr "): ") + cudaGetErrorString(e_);        \
            spp_population_destroy(p);                                                             \
            return fail(SPP_ERR_CUDA, m);                                                          \
        }                                                                                          \
    } while (0)
    ALLOC(p->params, P * L.params_size * sizeof(float));
    ALLOC(p->mom_m, P * L.train_size * sizeof(float));
    ALLOC(p->mom_v, P * L.train_size * sizeof(float));
    ALLOC(p->scratch, P * L.s.size * sizeof(float));
    ALLOC(p->scratch_acm, P * p->L_acm.s.size * sizeof(float));
    ALLOC(p->norm, P * NORM_COUNT * L.ldo * sizeof(float));
    ALLOC(p->acm_lim, 32 * sizeof(float));
    ALLOC(p->steps, P * 4 * sizeof(int));
    ALLOC(p->alpha_state, P * 4 * sizeof(double));
    p->S = cfg->buffer_size;
    if (p->S > 0) {
        if (p->S > 2000000000LL) { spp_population_destroy(p); return fail(SPP_ERR_ARG, "buffer_size too large"); }
        ALLOC(p->r_obs, P * p->S * L.ldo * sizeof(float));
        if (cfg->store_actions) ALLOC(p->r_act, P * p->S * L.ldo * sizeof(float));
        ALLOC(p->r_rew, P * p->S * sizeof(float));
        ALLOC(p->r_aacm, P * p->S * L.lda * sizeof(float));
        ALLOC(p->r_oidx, P * p->S * sizeof(int32_t));
        ALLOC(p->r_nidx, P * p->S * sizeof(int32_t));
        ALLOC(p->r_done, P * p->S);
        ALLOC(p->r_end, P * p->S);
    }
    ALLOC(p->r_len, P * sizeof(int64_t));
#undef ALLOC
    p->obs_cur.assign(P, 0); p->ts_cur.assign(P, 0); p->cur_len.assign(P, 0);
    if (cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking) != cudaSuccess) {
        spp_population_destroy(p);
        return fail(SPP_ERR_CUDA, "cudaStreamCreate failed");
    }
    // defaults: limits 1, identity (de)normalisation, log_alpha = log(alpha)
    std::vector<float> ones(128, 1.f);
    *out = p;
    int rc = spp_set_limits(p, ones.data(), ones.data());
    if (rc == SPP_OK) rc = spp_set_norm_stats(p, -1, nullptr, nullptr, nullptr, nullptr);
    if (rc == SPP_OK && cfg->algo == SPP_ALGO_SAC) rc = spp_alpha_set(p, -1, std::log(cfg->alpha > 0 ? cfg->alpha : 0.2));
    if (rc != SPP_OK) { std::string m = g_err; spp_population_destroy(p); *out = nullptr; return fail(rc, m); }
    return SPP_OK;
}

int spp_sync(spp_population* p) {
    if (!p) return fail(SPP_ERR_ARG, "null population");
    CK(cudaSetDevice(p->device));
    CK(cudaStreamSynchronize(p->stream));
    CK(cudaDeviceSynchronize());
    return SPP_OK;
}

int spp_set_limits(spp_population* p, const float* actor_lim, const float* acm_lim) {
    if (!p || !actor_lim || !acm_lim) return fail(SPP_ERR_ARG, "spp_set_limits: null argument");
    CK(cudaSetDevice(p->device));
    const Layout& L = p->L;
    std::vector<float> v(L.ldo, 0.f);
    for (int j = 0; j < L.ob; ++j) v[j] = actor_lim[j];
    for (int a = 0; a < p->P; ++a)
        H2D(p->norm + ((size_t)a * NORM_COUNT + NORM_LIM) * L.ldo, v.data(), L.ldo * sizeof(float));
    std::vector<float> m(32, 0.f);
    for (int j = 0; j < L.ac; ++j) m[j] = acm_lim[j];
    H2D(p->acm_lim, m.data(), 32 * sizeof(float));
    H2D_DONE();
    return SPP_OK;
}

int spp_set_norm_stats(spp_population* p, int a, const float* min_obs, const float* max_obs, const float* obs_mean,
                       const float* obs_std) {
    if (!p) return fail(SPP_ERR_ARG, "null population");
    if (a < -1 || a >= p->P) return fail(SPP_ERR_ARG, "agent index out of range");
    CK(cudaSetDevice(p->device));
    const Layout& L = p->L;
    std::vector<float> v(4 * (size_t)L.ldo, 0.f);
    float* doff = v.data(); float* dsc = doff + L.ldo; float* nsub = dsc + L.ldo; float* ndiv = nsub + L.ldo;
    for (int j = 0; j < L.ldo; ++j) { doff[j] = 0.f; dsc[j] = (j < L.ob) ? 1.f : 0.f; nsub[j] = 0.f; ndiv[j] = 1.f; }
    if (p->cfg.min_max_denormalize) {
        if (min_obs && max_obs)
            for (int j = 0; j < L.ob; ++j) {   // memory.py:80-82 and :119-121, evaluated in float32 like torch
                const float mean = (max_obs[j] + min_obs[j]) / 2.f;
                doff[j] = mean;
                dsc[j] = (max_obs[j] - min_obs[j]) / 2.f;
                nsub[j] = mean;
                ndiv[j] = (max_obs[j] - mean) + 1e-8f;
            }
    } else if (obs_mean && obs_std) {
        for (int j = 0; j < L.ob; ++j) {       // memory.py:123 and utils.py:70
            doff[j] = obs_mean[j];
            dsc[j] = obs_std[j] + 1e-8f;
            nsub[j] = obs_mean[j];
            ndiv[j] = obs_std[j] + 1e-8f;
        }
    }
    for (int i = (a < 0 ? 0 : a); i < (a < 0 ? p->P : a + 1); ++i)
        H2D(p->norm + (size_t)i * NORM_COUNT * L.ldo, v.data(), 4 * (size_t)L.ldo * sizeof(float));
    H2D_DONE();
    return SPP_OK;
}

// ---- parameters ------------------------------------------------------------------------------------
int spp_net_tensor_count(spp_population* p, int net) {
    if (!p || net < 0 || net >= NET_COUNT) return fail(SPP_ERR_ARG, "bad net id");
    return (int)p->tensors[net].size();
}

int spp_net_tensor_info(spp_population* p, int net, int t, char* name, int name_cap, int* rows, int* cols) {
    if (!p || net < 0 || net >= NET_COUNT || t < 0 || t >= (int)p->tensors[net].size()) return fail(SPP_ERR_ARG, "bad net/tensor id");
    const TensorMap& m = p->tensors[net][t];
    if (name && name_cap > 0) { strncpy(name, m.name.c_str(), name_cap - 1); name[name_cap - 1] = 0; }
    if (rows) *rows = m.rows;
    if (cols) *cols = m.cols;
    return SPP_OK;
}

// direction: 0 upload params, 1 download params, 2 download m, 3 download v
// Weights exist twice on the device (natural W [rows x ld] and W^T [ld x ld_t], see layout.h).  Uploads write both;
// downloads read W^T where it exists (target nets are only maintained in that copy).  Adam moments follow the
// natural layout except for the actor heads, whose dW tile is computed transposed (update_kernel.cu).
static int tensor_io(spp_population* p, int a, int net, int t, float* host, int dir) {
    if (!p || !host) return fail(SPP_ERR_ARG, "null argument");
    if (net < 0 || net >= NET_COUNT || t < 0 || t >= (int)p->tensors[net].size()) return fail(SPP_ERR_ARG, "bad net/tensor id");
    if (a < (dir == 0 ? -1 : 0) || a >= p->P) return fail(SPP_ERR_ARG, "agent index out of range");
    if (dir >= 2 && net > NET_ACM) return fail(SPP_ERR_ARG, "target nets have no optimiser state");
    CK(cudaSetDevice(p->device));
    const Layout& L = p->L;
    const TensorMap& m = p->tensors[net][t];
    const LayerDesc& l = net_desc(p, net).L[m.layer];
    const size_t stride = (dir >= 2) ? L.train_size : L.params_size;
    float* arena = dir == 2 ? p->mom_m : dir == 3 ? p->mom_v : p->params;
    const bool is_w = (m.is_bias == 0);
    const bool has_t = is_w && l.off_wt >= 0;
    const bool heads = (net == NET_ACTOR || net == NET_ACTOR_TARG) && m.layer == 2;
    // which copy a download reads: parameters from W^T when present; moments from the layout the optimiser uses
    const bool read_t = has_t && (dir == 1 || heads);
    std::vector<float> nat, tr;
    if (is_w) {
        nat.assign((size_t)m.rows * l.ld, 0.f);
        if (has_t) tr.assign((size_t)l.ld * l.ld_t, 0.f);   // whole W^T block of the layer (heads: both halves)
    } else {
        nat.assign(m.rows, 0.f);
    }
    for (int i = (a < 0 ? 0 : a); i < (a < 0 ? p->P : a + 1); ++i) {
        float* base = arena + (size_t)i * stride + L.net_off[net];
        float* dev_nat = is_w ? base + l.off_w + (size_t)m.row0 * l.ld
                              : base + (m.is_bias == 1 ? l.off_b + m.row0 : l.off_w + m.col0);
        float* dev_t = has_t ? base + l.off_wt : nullptr;
        if (dir == 0) {
            if (is_w) {
                for (int r = 0; r < m.rows; ++r)
                    for (int c = 0; c < m.cols; ++c) nat[(size_t)r * l.ld + map_col(l, c)] = host[(size_t)r * m.cols + c];
                H2D(dev_nat, nat.data(), nat.size() * sizeof(float));
                H2D_DONE();
                if (has_t) {   // read-modify-write of the W^T block: only columns [row0, row0 + rows) belong to this tensor
                    CK(cudaMemcpy(tr.data(), dev_t, tr.size() * sizeof(float), cudaMemcpyDeviceToHost));
                    for (int r = 0; r < m.rows; ++r)
                        for (int c = 0; c < m.cols; ++c)
                            tr[(size_t)map_col(l, c) * l.ld_t + m.row0 + r] = host[(size_t)r * m.cols + c];
                    H2D(dev_t, tr.data(), tr.size() * sizeof(float));
                    H2D_DONE();
                }
            } else {
                for (int r = 0; r < m.rows; ++r) nat[r] = host[r];
                H2D(dev_nat, nat.data(), nat.size() * sizeof(float));
                H2D_DONE();
            }
        } else {
            CK(cudaStreamSynchronize(p->stream));
            if (is_w && read_t) {
                CK(cudaMemcpy(tr.data(), dev_t, tr.size() * sizeof(float), cudaMemcpyDeviceToHost));
                for (int r = 0; r < m.rows; ++r)
                    for (int c = 0; c < m.cols; ++c)
                        host[(size_t)r * m.cols + c] = tr[(size_t)map_col(l, c) * l.ld_t + m.row0 + r];
            } else {
                CK(cudaMemcpy(nat.data(), dev_nat, nat.size() * sizeof(float), cudaMemcpyDeviceToHost));
                if (is_w)
                    for (int r = 0; r < m.rows; ++r)
                        for (int c = 0; c < m.cols; ++c) host[(size_t)r * m.cols + c] = nat[(size_t)r * l.ld + map_col(l, c)];
                else
                    for (int r = 0; r < m.rows; ++r) host[r] = nat[r];
            }
        }
    }
    return SPP_OK;
}

int spp_params_upload(spp_population* p, int a, int net, int t, const float* host) { return tensor_io(p, a, net, t, const_cast<float*>(host), 0); }
int spp_params_download(spp_population* p, int a, int net, int t, float* host) { return tensor_io(p, a, net, t, host, 1); }

int spp_adam_download(spp_population* p, int a, int net, int t, float* exp_avg, float* exp_avg_sq, int* step) {
    int rc = SPP_OK;
    if (exp_avg) rc = tensor_io(p, a, net, t, exp_avg, 2);
    if (rc == SPP_OK && exp_avg_sq) rc = tensor_io(p, a, net, t, exp_avg_sq, 3);
    if (rc == SPP_OK && step) {
        if (!p || net < 0 || net > NET_ACM || a < 0 || a >= p->P) return fail(SPP_ERR_ARG, "bad net/agent");
        CK(cudaStreamSynchronize(p->stream));
        CK(cudaMemcpy(step, p->steps + (size_t)a * 4 + net, sizeof(int), cudaMemcpyDeviceToHost));
    }
    return rc;
}

int spp_adam_reset(spp_population* p, int a, int net) {
    if (!p || net < 0 || net > NET_ACM || a < -1 || a >= p->P) return fail(SPP_ERR_ARG, "bad net/agent");
    CK(cudaSetDevice(p->device));
    CK(cudaStreamSynchronize(p->stream));
    const Layout& L = p->L;
    const size_t n = net_desc(p, net).size;
    for (int i = (a < 0 ? 0 : a); i < (a < 0 ? p->P : a + 1); ++i) {
        CK(cudaMemset(p->mom_m + (size_t)i * L.train_size + L.net_off[net], 0, n * sizeof(float)));
        CK(cudaMemset(p->mom_v + (size_t)i * L.train_size + L.net_off[net], 0, n * sizeof(float)));
        CK(cudaMemset(p->steps + (size_t)i * 4 + net, 0, sizeof(int)));
    }
    return SPP_OK;
}

int spp_sync_targets(spp_population* p, int a) {
    if (!p || a < -1 || a >= p->P) return fail(SPP_ERR_ARG, "bad agent");
    CK(cudaSetDevice(p->device));
    CK(cudaStreamSynchronize(p->stream));
    const Layout& L = p->L;
    for (int i = (a < 0 ? 0 : a); i < (a < 0 ? p->P : a + 1); ++i) {
        float* base = p->params + (size_t)i * L.params_size;
        CK(cudaMemcpy(base + L.net_off[NET_CRITIC_1_TARG], base + L.net_off[NET_CRITIC_1], L.critic.size * sizeof(float), cudaMemcpyDeviceToDevice));
        CK(cudaMemcpy(base + L.net_off[NET_CRITIC_2_TARG], base + L.net_off[NET_CRITIC_2], L.critic.size * sizeof(float), cudaMemcpyDeviceToDevice));
        CK(cudaMemcpy(base + L.net_off[NET_ACTOR_TARG], base + L.net_off[NET_ACTOR], L.actor.size * sizeof(float), cudaMemcpyDeviceToDevice));
    }
    return SPP_OK;
}

int spp_alpha_get(spp_population* p, int a, double* log_alpha, double* alpha) {
    if (!p || a < 0 || a >= p->P) return fail(SPP_ERR_ARG, "bad agent");
    CK(cudaSetDevice(p->device));
    CK(cudaStreamSynchronize(p->stream));
    double la = 0;
    CK(cudaMemcpy(&la, p->alpha_state + (size_t)a * 4, sizeof(double), cudaMemcpyDeviceToHost));
    if (log_alpha) *log_alpha = la;
    if (alpha) *alpha = std::exp(la);
    return SPP_OK;
}

int spp_alpha_set(spp_population* p, int a, double log_alpha) {
    if (!p || a < -1 || a >= p->P) return fail(SPP_ERR_ARG, "bad agent");
    CK(cudaSetDevice(p->device));
    const double st[4] = {log_alpha, 0.0, 0.0, 0.0};
    for (int i = (a < 0 ? 0 : a); i < (a < 0 ? p->P : a + 1); ++i)
        H2D(p->alpha_state + (size_t)i * 4, st, sizeof(st));
    H2D_DONE();
    return SPP_OK;
}

// ReplayBuffer._sample_batch normalises obs / next_obs when the buffer was built with obs_norm=True (rltoolkit/buffer/replay_buffer.py:
// 246-248, gate :76-80): the fused ring updates then gather NORMALISED rows (statistics of spp_set_norm_stats).  Explicit minibatches
// (spp_update_host) are taken as given -- the reference's update() receives what sample_batch returned.
int spp_set_obs_norm(spp_population* p, int on) {
    if (!p) return fail(SPP_ERR_ARG, "null population");
    p->h.obs_norm = on ? 1 : 0;
    return SPP_OK;
}

int spp_set_learning_rates(spp_population* p, double actor_lr, double critic_lr, double alpha_lr, double acm_lr) {
    if (!p) return fail(SPP_ERR_ARG, "null population");
    if (actor_lr >= 0) p->h.actor_lr = actor_lr;
    if (critic_lr >= 0) p->h.critic_lr = critic_lr;
    if (alpha_lr >= 0) p->h.alpha_lr = alpha_lr;
    if (acm_lr >= 0) p->h.acm_lr = acm_lr;
    return SPP_OK;
}

// ---- replay ring -----------------------------------------------------------------------------------
static int ring_check(spp_population* p, int a) {
    if (!p) return fail(SPP_ERR_ARG, "null population");
    if (a < 0 || a >= p->P) return fail(SPP_ERR_ARG, "agent index out of range");
    if (p->S <= 0) return fail(SPP_ERR_STATE, "population was created without a replay ring (buffer_size = 0)");
    return SPP_OK;
}

int spp_ring_add_obs(spp_population* p, int a, const float* obs, int64_t* out_idx) {
    int rc = ring_check(p, a); if (rc) return rc;
    if (!obs) return fail(SPP_ERR_ARG, "null obs");
    CK(cudaSetDevice(p->device));
    const Layout& L = p->L;
    const int64_t i = p->obs_cur[a];
    H2D(p->r_obs + ((size_t)a * p->S + i) * L.ldo, obs, L.ob * sizeof(float));
    H2D_DONE();
    p->obs_cur[a] = (i + 1) % p->S;                        // replay_buffer.py:56-60
    if (out_idx) *out_idx = i;
    return SPP_OK;
}

int spp_ring_add_acm_action(spp_population* p, int a, const float* acm_action) {
    int rc = ring_check(p, a); if (rc) return rc;
    if (!acm_action) return fail(SPP_ERR_ARG, "null acm_action");
    CK(cudaSetDevice(p->device));
    H2D(p->r_aacm + ((size_t)a * p->S + p->ts_cur[a]) * p->L.lda, acm_action, p->L.ac * sizeof(float));
    H2D_DONE();
    return SPP_OK;                                         // replay_buffer.py:332-333
}

int spp_ring_add_timestep(spp_population* p, int a, int64_t obs_idx, int64_t next_obs_idx, const float* action,
                          float reward, int done, int end) {
    int rc = ring_check(p, a); if (rc) return rc;
    if (obs_idx < 0 || obs_idx >= p->S || next_obs_idx < 0 || next_obs_idx >= p->S) return fail(SPP_ERR_ARG, "obs index out of range");
    CK(cudaSetDevice(p->device));
    const Layout& L = p->L;
    const int64_t ts = p->ts_cur[a];
    const size_t row = (size_t)a * p->S + ts;
    const int32_t oi = (int32_t)obs_idx, ni = (int32_t)next_obs_idx;
    const uint8_t d = done ? 1 : 0, e = end ? 1 : 0;
    H2D(p->r_oidx + row, &oi, 4);
    H2D(p->r_nidx + row, &ni, 4);
    if (action && p->r_act) H2D(p->r_act + row * L.ldo, action, L.ob * sizeof(float));
    H2D(p->r_rew + row, &reward, 4);
    H2D(p->r_done + row, &d, 1);
    H2D(p->r_end + row, &e, 1);
    H2D_DONE();      // one synchronisation for the six row writes
    // replay_buffer.py:70-75
    if (next_obs_idx < ts) { p->cur_len[a] = ts + 1; p->ts_cur[a] = 0; }
    else p->ts_cur[a] = ts + 1;
    p->cur_len[a] = p->ts_cur[a] > p->cur_len[a] ? p->ts_cur[a] : p->cur_len[a];
    p->len_dirty = true;
    return SPP_OK;
}

// ReplayBufferAcM.add_buffer (rltoolkit/buffer/replay_buffer.py:284-297) for the [T][E] store a device rollout left in `store`
// (spp_ppo_rollout_synthetic): the store's trajectories are taken environment by environment, each cut into rollouts at its `end`
// flags, as the chain of observations (terminal observations included) + one ACM action per transition that MemoryAcM would hold.
// The loop below is the reference's, joint behaviour included (the first transition after a joint is skipped and the chain index
// runs one ahead from there on) -- on integers only: it decides which store row ends up in which ring
// slot under the cursor state machine of add_obs / add_timestep (:56-75), and one kernel then moves the rows on the device.
int spp_ring_add_rollout_store(spp_population* p, int a, spp_ppo* store) {
    int rc = ring_check(p, a); if (rc) return rc;
    spp::PpoStoreView st;
    rc = spp_ppo_store_view_(store, &st); if (rc) return rc;
    const Layout& L = p->L;
    if (st.device != p->device || st.ob != L.ob || st.lda != L.lda || st.ldo != L.ldo) return fail(SPP_ERR_ARG, "population and rollout store disagree (device / shapes)");
    CK(cudaSetDevice(p->device));
    const int E = st.E, T = st.T;
    const int64_t N = (int64_t)E * T, S = p->S;
    // `end` flags, environment-major (the order the chain is walked in), one byte each: transposed on the device, 1 byte per row D2H
    const bool dbg_t = getenv("SPP_DEBUG_TIMING") != nullptr;
    auto now = [] { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; };
    const double t_a = now();
    std::vector<uint8_t> endT((size_t)N);
    {
        CK(p->d_add_flags.ensure((size_t)N));
        CK(cudaStreamWaitEvent(p->stream, st.ready, 0));      // the rollout that filled the store, not whatever the policy's stream runs now
        CK(spp::launch_end_flags_env_major(st.end, E, T, (uint8_t*)p->d_add_flags.p, p->stream));
        CK(cudaMemcpyAsync(endT.data(), p->d_add_flags.p, (size_t)N, cudaMemcpyDeviceToHost, p->stream));
        CK(cudaStreamSynchronize(p->stream));
    }
    const double t_b = now();
    constexpr int64_t kNext = 1ll << 62;
    // The chain is never materialised: a cursor walks it.  Chain entries, in order: for every transition q (environment-major) its
    // observation, followed by its NEXT observation when the transition ends a rollout; the entry after a terminal one is a joint.
    struct ChainCursor {      // (e, t) kept incrementally: no 64-bit division in the 8 M-step walk
        const uint8_t* endT; int64_t q; int e, t; bool terminal; bool joint; int E, T; int64_t N;
        int64_t src() const { const int64_t row = (int64_t)t * E + e; return terminal ? (row | kNext) : row; }
        void advance() {
            if (!terminal && q < N && endT[q]) { terminal = true; joint = false; }
            else { joint = terminal; terminal = false; ++q; if (q < N && ++t == T) { t = 0; ++e; } }
        }
    } c{endT.data(), 0, 0, 0, false, false, E, T, N};
    std::vector<int64_t> obs_src((size_t)S, -1), ts_src((size_t)S, -1);
    std::vector<int32_t> ts_oidx((size_t)S, 0), ts_nidx((size_t)S, 0);
    int64_t obs_cur = p->obs_cur[a], ts_cur = p->ts_cur[a], cur_len = p->cur_len[a];
    auto add_obs = [&](int64_t src) { const int64_t i = obs_cur; obs_src[i] = src; obs_cur = i + 1 == S ? 0 : i + 1; return i; };
    // Slots written more than ~3 ring laps before the end are overwritten later (the observation cursor advances once per step, the
    // timestep cursor at most once and never runs more than a lap behind): only the cursors are walked there, nothing is recorded.
    const int64_t k_rec = N - 3 * S - 64;
    bool rec = k_rec <= 0;
    auto add_obs_fast = [&]() { const int64_t i = obs_cur; obs_cur = i + 1 == S ? 0 : i + 1; return i; };
    int64_t obs_idx = rec ? add_obs(c.src()) : add_obs_fast();
    int ke = 0, kt = 0;                                // the k-th ACM action in environment-major order is store row kt * E + ke
    for (int64_t k = 0; k < N; ++k) {
        const int64_t krow = (int64_t)kt * E + ke;
        if (++kt == T) { kt = 0; ++ke; }
        if (k == k_rec) rec = true;
        c.advance();                                   // i += 1
        const int64_t next_idx = rec ? add_obs(c.src()) : add_obs_fast();
        if (c.joint) { c.advance(); continue; }        // a new rollout starts here: i += 1, no timestep (quirk 19)
        const int64_t ts = ts_cur;
        if (rec) {
            ts_src[ts] = krow;
            ts_oidx[ts] = (int32_t)obs_idx; ts_nidx[ts] = (int32_t)next_idx;
        }
        if (next_idx < ts) { cur_len = ts + 1; ts_cur = 0; } else ts_cur = ts + 1;      // replay_buffer.py:70-75
        cur_len = ts_cur > cur_len ? ts_cur : cur_len;
        obs_idx = next_idx;
    }
    const double t_c = now();
    DevBuf& d1 = p->d_add_src[0]; DevBuf& d2 = p->d_add_src[1]; DevBuf& d3 = p->d_add_src[2]; DevBuf& d4 = p->d_add_src[3];
    CK(d1.ensure((size_t)S * 8)); CK(d2.ensure((size_t)S * 8)); CK(d3.ensure((size_t)S * 4)); CK(d4.ensure((size_t)S * 4));
    CK(cudaMemcpyAsync(d1.p, obs_src.data(), (size_t)S * 8, cudaMemcpyHostToDevice, p->stream));
    CK(cudaMemcpyAsync(d2.p, ts_src.data(), (size_t)S * 8, cudaMemcpyHostToDevice, p->stream));
    CK(cudaMemcpyAsync(d3.p, ts_oidx.data(), (size_t)S * 4, cudaMemcpyHostToDevice, p->stream));
    CK(cudaMemcpyAsync(d4.p, ts_nidx.data(), (size_t)S * 4, cudaMemcpyHostToDevice, p->stream));
    const size_t base = (size_t)a * S;
    CK(spp::launch_ring_add_store(p->r_obs + base * L.ldo, p->r_oidx + base, p->r_nidx + base, p->r_aacm + base * L.lda, p->r_rew + base,
                                  p->r_done + base, p->r_end + base, S, L.ob, L.ac, L.ldo, L.lda, (const int64_t*)d1.p, (const int64_t*)d2.p,
                                  (const int32_t*)d3.p, (const int32_t*)d4.p, st, p->stream));
    g_launches++;
    CK(cudaStreamSynchronize(p->stream));
    p->obs_cur[a] = obs_cur; p->ts_cur[a] = ts_cur; p->cur_len[a] = cur_len;
    p->len_dirty = true;
    if (dbg_t) fprintf(stderr, "spp_ring_add_rollout_store: flags %.1f ms, walk %.1f ms, rows %.1f ms\n", t_b - t_a, t_c - t_b, now() - t_c);
    return SPP_OK;
}

int spp_ring_reset(spp_population* p, int a) {
    if (!p || a < -1 || a >= p->P) return fail(SPP_ERR_ARG, "bad agent");
    for (int i = (a < 0 ? 0 : a); i < (a < 0 ? p->P : a + 1); ++i) { p->obs_cur[i] = 0; p->ts_cur[i] = 0; p->cur_len[i] = 0; }
    p->len_dirty = true;
    return SPP_OK;
}

int spp_ring_state(spp_population* p, int a, int64_t out[3]) {
    if (!p || a < 0 || a >= p->P || !out) return fail(SPP_ERR_ARG, "bad argument");
    out[0] = p->obs_cur[a]; out[1] = p->ts_cur[a]; out[2] = p->cur_len[a];
    return SPP_OK;
}

static RingView ring_view(const spp_population* p) {
    RingView R{p->r_obs, p->r_oidx, p->r_nidx, p->r_act, p->r_rew, p->r_done, p->r_aacm, p->S, p->L.ob, p->L.ac, p->L.ldo, p->L.lda};
    return R;
}

static int push_ring_len(spp_population* p) {
    if (!p->len_dirty) return SPP_OK;
    CK(cudaMemcpyAsync(p->r_len, p->cur_len.data(), p->P * sizeof(int64_t), cudaMemcpyHostToDevice, p->stream));
    CK(cudaStreamSynchronize(p->stream));
    p->len_dirty = false;
    return SPP_OK;
}

int spp_ring_sample_batch(spp_population* p, int a, const int64_t* idx, int n, float* obs, float* next_obs, float* action,
                          float* reward, int8_t* done, float* acm_action) {
    int rc = ring_check(p, a); if (rc) return rc;
    if (!idx || n < 0 || !obs || !next_obs || !reward || !done || !acm_action) return fail(SPP_ERR_ARG, "null argument");
    if (n == 0) return SPP_OK;
    for (int i = 0; i < n; ++i)
        if (idx[i] < 0 || idx[i] >= p->cur_len[a]) return fail(SPP_ERR_ARG, "sample index outside [0, len(buffer))");
    CK(cudaSetDevice(p->device));
    const Layout& L = p->L;
    const size_t fo = (size_t)n * L.ob, fa = (size_t)n * L.ac;
    const size_t bytes = (3 * fo + fa + n) * sizeof(float) + n + 64;
    CK(p->d_tmp.ensure(bytes));
    CK(p->d_idx.ensure((size_t)n * sizeof(int64_t)));
    float* d = (float*)p->d_tmp.p;
    GatherOut o{d, d + fo, (action && p->r_act) ? d + 2 * fo : nullptr, d + 3 * fo + fa, (int8_t*)(d + 3 * fo + fa + n), d + 3 * fo};
    CK(cudaMemcpyAsync(p->d_idx.p, idx, (size_t)n * sizeof(int64_t), cudaMemcpyHostToDevice, p->stream));
    CK(launch_ring_gather(ring_view(p), a, (const int64_t*)p->d_idx.p, n, o, p->stream));
    g_launches++;
    CK(cudaMemcpyAsync(obs, o.obs, fo * sizeof(float), cudaMemcpyDeviceToHost, p->stream));
    CK(cudaMemcpyAsync(next_obs, o.nobs, fo * sizeof(float), cudaMemcpyDeviceToHost, p->stream));
    if (action) {
        if (o.act) CK(cudaMemcpyAsync(action, o.act, fo * sizeof(float), cudaMemcpyDeviceToHost, p->stream));
        else memset(action, 0, fo * sizeof(float));
    }
    CK(cudaMemcpyAsync(acm_action, o.aacm, fa * sizeof(float), cudaMemcpyDeviceToHost, p->stream));
    CK(cudaMemcpyAsync(reward, o.rew, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, p->stream));
    CK(cudaMemcpyAsync(done, o.done, (size_t)n, cudaMemcpyDeviceToHost, p->stream));
    CK(cudaStreamSynchronize(p->stream));
    return SPP_OK;
}

int spp_ring_fill_synthetic(spp_population* p, uint64_t seed, int64_t n, int episode_len) {
    if (!p || p->S <= 0) return fail(SPP_ERR_STATE, "no replay ring");
    if (n < 1 || episode_len < 1) return fail(SPP_ERR_ARG, "n and episode_len must be positive");
    const int64_t n_eps = (n + episode_len - 1) / episode_len;
    if (n + n_eps > p->S) return fail(SPP_ERR_ARG, "n transitions plus one terminal observation per episode must fit buffer_size");
    CK(cudaSetDevice(p->device));
    const Layout& L = p->L;
    CK(launch_ring_fill(ring_view(p), p->r_obs, p->r_oidx, p->r_nidx, p->r_act, p->r_rew, p->r_done, p->r_end, p->r_aacm, p->P, n,
                        episode_len, seed, p->norm, NORM_COUNT * L.ldo, p->sm_count * 8, p->stream));
    g_launches++;
    CK(cudaStreamSynchronize(p->stream));
    for (int a = 0; a < p->P; ++a) { p->obs_cur[a] = (n + n_eps) % p->S; p->ts_cur[a] = n % p->S; p->cur_len[a] = n; }
    p->len_dirty = true;
    return SPP_OK;
}

// MetaReplayBuffer.update_obs_mean_std (rltoolkit/buffer/replay_buffer.py:83-96) for every agent: out [P][6][ob] =
// mean, population std, and the order statistics x[floor(v1)], x[ceil(v1)], x[floor(v99)], x[ceil(v99)] with v_q = q (n - 1)
int spp_ring_obs_stats(spp_population* p, double* out, double* bytes_out) {
    if (!p || !out) return fail(SPP_ERR_ARG, "spp_ring_obs_stats: null argument");
    if (p->S <= 0) return fail(SPP_ERR_STATE, "no replay ring");
    for (int a = 0; a < p->P; ++a) if (p->cur_len[a] < 1) return fail(SPP_ERR_STATE, "replay ring is empty");
    CK(cudaSetDevice(p->device));
    int rc = push_ring_len(p); if (rc) return rc;
    const int ob = p->L.ob, P = p->P;
    int nb = (8 * p->sm_count) / P; if (nb < 8) nb = 8; if (nb > 1024) nb = 1024;
    const size_t n_state = (size_t)P * ob * 4 * 2;
    CK(p->d_stat_partial.ensure((size_t)P * nb * ob * sizeof(double)));
    CK(p->d_stat_moments.ensure((size_t)2 * P * ob * sizeof(double)));
    CK(p->d_stat_state.ensure(n_state * sizeof(unsigned long long)));
    CK(p->d_stat_hist.ensure((size_t)P * ob * 4 * 256 * sizeof(unsigned int)));
    std::vector<unsigned long long> h_state(n_state);
    CK(launch_ring_obs_stats(ring_view(p), P, p->r_len, p->cur_len.data(), nb, (double*)p->d_stat_partial.p, (double*)p->d_stat_moments.p,
                             (unsigned long long*)p->d_stat_state.p, (unsigned int*)p->d_stat_hist.p, h_state.data(), p->stream));
    g_launches += 12;
    std::vector<double> mom((size_t)2 * P * ob);
    CK(cudaMemcpyAsync(mom.data(), p->d_stat_moments.p, mom.size() * sizeof(double), cudaMemcpyDeviceToHost, p->stream));
    CK(cudaMemcpyAsync(h_state.data(), p->d_stat_state.p, n_state * sizeof(unsigned long long), cudaMemcpyDeviceToHost, p->stream));
    CK(cudaStreamSynchronize(p->stream));
    double bytes = 0;
    for (int a = 0; a < P; ++a) {
        bytes += 6.0 * (double)p->cur_len[a] * (ob * 4.0 + 4.0);      // 6 passes over len rows: ob floats + the obs index
        for (int j = 0; j < ob; ++j) {
            double* o = out + (size_t)a * 6 * ob;
            o[0 * ob + j] = mom[(size_t)a * ob + j];
            o[1 * ob + j] = mom[(size_t)(P + a) * ob + j];
            for (int T = 0; T < 4; ++T)
                o[(2 + T) * ob + j] = (double)ring_stats_key_to_float((uint32_t)h_state[(((size_t)a * ob + j) * 4 + T) * 2]);
        }
    }
    if (bytes_out) *bytes_out = bytes;
    return SPP_OK;
}

int spp_ring_gather_bench_device(spp_population* p, int n_batches, uint64_t seed, double* bytes_out, void* stream) {
    if (!p || p->S <= 0) return fail(SPP_ERR_STATE, "no replay ring");
    if (n_batches < 1) return fail(SPP_ERR_ARG, "n_batches must be positive");
    CK(cudaSetDevice(p->device));
    int rc = push_ring_len(p); if (rc) return rc;
    const Layout& L = p->L;
    const size_t rows = (size_t)p->P * n_batches * L.B;
    const size_t bytes = rows * ((2 * (size_t)L.ldo + L.lda + 1) * sizeof(float) + 1) + 256;
    CK(p->d_tmp.ensure(bytes));
    float* o_obs = (float*)p->d_tmp.p; float* o_nobs = o_obs + rows * L.ldo; float* o_aacm = o_nobs + rows * L.ldo;
    float* o_rew = o_aacm + rows * L.lda; uint8_t* o_done = (uint8_t*)(o_rew + rows);
    cudaStream_t s = stream ? (cudaStream_t)stream : p->stream;
    CK(launch_ring_gather_bench(ring_view(p), p->P, n_batches, L.B, p->r_len, seed, o_obs, o_nobs, o_aacm, o_rew, o_done,
                                p->sm_count * 8, s));
    g_launches++;
    // algorithmic bytes: B * ((2 ob + ac + 1) * 4 + 1) read and the same written, plus 2 index words read
    if (bytes_out) *bytes_out = (double)rows * (2.0 * ((2.0 * L.ob + L.ac + 1) * 4 + 1) + 8.0);
    return SPP_OK;
}

// rows [first, first + n) of the dense minibatches the last spp_ring_gather_bench_device call left on the device (test hook: the
// bench kernel draws its indices on the device -- Philox4x32-10(seed, agent, row in the agent's block) -> mulhi(x.x:x.y, len) --
// so a test can recompute them and compare with spp_ring_sample_batch)
int spp_ring_gather_bench_rows(spp_population* p, int n_batches, int64_t first, int n, float* obs, float* nobs, float* aacm, float* rew,
                               int8_t* done) {
    if (!p || !obs || !nobs || !aacm || !rew || !done || n < 1 || first < 0) return fail(SPP_ERR_ARG, "bad argument");
    const Layout& L = p->L;
    const size_t rows = (size_t)p->P * n_batches * L.B;
    if ((size_t)first + n > rows || !p->d_tmp.p) return fail(SPP_ERR_ARG, "rows outside the last gather");
    CK(cudaSetDevice(p->device));
    CK(cudaDeviceSynchronize());
    const float* o_obs = (const float*)p->d_tmp.p; const float* o_nobs = o_obs + rows * L.ldo; const float* o_aacm = o_nobs + rows * L.ldo;
    const float* o_rew = o_aacm + rows * L.lda; const uint8_t* o_done = (const uint8_t*)(o_rew + rows);
    std::vector<float> tmp((size_t)n * (L.ldo > L.lda ? L.ldo : L.lda));
    auto rows_out = [&](const float* src, int ld, int w, float* dst) -> cudaError_t {
        cudaError_t e = cudaMemcpy(tmp.data(), src + (size_t)first * ld, (size_t)n * ld * 4, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) return e;
        for (int r = 0; r < n; ++r) memcpy(dst + (size_t)r * w, tmp.data() + (size_t)r * ld, (size_t)w * 4);
        return cudaSuccess;
    };
    CK(rows_out(o_obs, L.ldo, L.ob, obs)); CK(rows_out(o_nobs, L.ldo, L.ob, nobs)); CK(rows_out(o_aacm, L.lda, L.ac, aacm));
    CK(cudaMemcpy(rew, o_rew + first, (size_t)n * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(done, o_done + first, (size_t)n, cudaMemcpyDeviceToHost));
    return SPP_OK;
}

// ---- update ----------------------------------------------------------------------------------------
// 1: the 128 x 128 GEMMs of the update burst run on tcgen05 (3-pass tf32 split); 0: FFMA tiles.  Both are device paths;
// the switch exists for A/B measurement and kernel bisection (environment SPP_UMMA=0 or spp_set_gemm_path).
static std::atomic<int> g_gemm_path{[] { const char* e = getenv("SPP_UMMA"); return (e && e[0] == '0') ? 0 : ((e && e[0] == '2') ? 2 : 1); }()};

static void fill_args(spp_population* p, UpdateArgs& a, int G) {
    memset(&a, 0, sizeof(a));
    a.L = p->L; a.h = p->h;
    a.params = p->params; a.mom_m = p->mom_m; a.mom_v = p->mom_v; a.scratch = p->scratch;
    a.steps = p->steps; a.alpha_state = p->alpha_state; a.norm = p->norm; a.acm_lim = p->acm_lim;
    a.ring = RingPtrs{p->r_obs, p->r_oidx, p->r_nidx, p->r_act, p->r_rew, p->r_done, p->r_aacm, p->S};
    a.ring_len = p->r_len;
    a.G = G; a.population = p->P;
    a.seq = p->seq++;
    a.use_umma = g_gemm_path.load();
}

static int grid_for(const spp_population* p) { return p->P < p->sm_count ? p->P : p->sm_count; }

// Update bursts of a population larger than the grid run as interleaved (step, agent) work items (update_burst_kernel): the
// per-agent progress words are zeroed on the launch stream right before the kernel.  SPP_BALANCE=0 keeps whole agents per CTA (A/B).
static cudaError_t launch_update_balanced(spp_population* p, UpdateArgs& a, cudaStream_t s) {
    const char* env = getenv("SPP_BALANCE");      // read per launch: the parity test flips it inside one process
    const bool balance = !(env && env[0] == '0');
    const int grid = grid_for(p);
    a.progress = nullptr;
    if (balance && p->P > grid && a.G > 1 && (long long)p->P * a.G < (1ll << 31)) {
        cudaError_t e = p->d_progress.ensure((size_t)p->P * sizeof(unsigned int));
        if (e != cudaSuccess) return e;
        e = cudaMemsetAsync(p->d_progress.p, 0, (size_t)p->P * sizeof(unsigned int), s);
        if (e != cudaSuccess) return e;
        a.progress = (unsigned int*)p->d_progress.p;
    }
    // the wide products' main loop: landing-zone loop (default path), two-slot loop for the single-pass variant and on request
    const char* loop = getenv("SPP_UMMA_LOOP");
    const bool twoslot = a.use_umma == 2 || (loop && loop[0] == 't');
    return twoslot ? launch_update_burst_twoslot(a, grid, s) : launch_update_burst(a, grid, s);
}

int spp_update_host(spp_population* p, int G, const float* obs, const float* next_obs, const float* action,
                    const float* reward, const int8_t* done, const float* acm_action, const float* eps, uint64_t seed,
                    float* losses) {
    if (!p || !obs || !next_obs || !reward || !done || !acm_action) return fail(SPP_ERR_ARG, "spp_update_host: null batch tensor");
    if (G < 1) return fail(SPP_ERR_ARG, "grad_steps must be positive");
    if (!p->cfg.acm_critic && !action) return fail(SPP_ERR_ARG, "acm_critic=0 needs the action tensor");
    CK(cudaSetDevice(p->device));
    const Layout& L = p->L;
    const size_t rows = (size_t)p->P * G * L.B;
    cudaStream_t s = p->stream;
    CK(p->d_obs.ensure(rows * L.ob * 4)); CK(p->d_nobs.ensure(rows * L.ob * 4));
    CK(p->d_rew.ensure(rows * 4)); CK(p->d_done.ensure(rows)); CK(p->d_aacm.ensure(rows * L.ac * 4));
    CK(cudaMemcpyAsync(p->d_obs.p, obs, rows * L.ob * 4, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(p->d_nobs.p, next_obs, rows * L.ob * 4, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(p->d_rew.p, reward, rows * 4, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(p->d_done.p, done, rows, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(p->d_aacm.p, acm_action, rows * L.ac * 4, cudaMemcpyHostToDevice, s));
    const bool need_act = !p->cfg.acm_critic;
    if (need_act) {
        CK(p->d_act.ensure(rows * L.ob * 4));
        CK(cudaMemcpyAsync(p->d_act.p, action, rows * L.ob * 4, cudaMemcpyHostToDevice, s));
    }
    const bool sac = p->cfg.algo == SPP_ALGO_SAC;
    if (eps && sac) {
        CK(p->d_eps.ensure(rows * 2 * L.ob * 4));
        CK(cudaMemcpyAsync(p->d_eps.p, eps, rows * 2 * L.ob * 4, cudaMemcpyHostToDevice, s));
    }
    CK(p->d_losses.ensure((size_t)p->P * G * LOSS_COUNT * 4));
    UpdateArgs a;
    fill_args(p, a, G);
    a.batch = BatchPtrs{(const float*)p->d_obs.p, (const float*)p->d_nobs.p, need_act ? (const float*)p->d_act.p : nullptr,
                        (const float*)p->d_rew.p, (const int8_t*)p->d_done.p, (const float*)p->d_aacm.p};
    a.eps = (eps && sac) ? (const float*)p->d_eps.p : nullptr;
    a.seed = seed;
    a.losses = (float*)p->d_losses.p;
    CK(launch_update_balanced(p, a, s));
    g_launches++;
    if (losses) CK(cudaMemcpyAsync(losses, p->d_losses.p, (size_t)p->P * G * LOSS_COUNT * 4, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return SPP_OK;
}

int spp_update_ring(spp_population* p, int G, const int64_t* idx, const float* eps, uint64_t seed, float* losses) {
    if (!p) return fail(SPP_ERR_ARG, "null population");
    if (p->S <= 0) return fail(SPP_ERR_STATE, "no replay ring");
    if (G < 1) return fail(SPP_ERR_ARG, "grad_steps must be positive");
    CK(cudaSetDevice(p->device));
    const Layout& L = p->L;
    const size_t rows = (size_t)p->P * G * L.B;
    for (int a = 0; a < p->P; ++a)
        if (p->cur_len[a] < 1) return fail(SPP_ERR_STATE, "replay ring is empty");
    cudaStream_t s = p->stream;
    if (idx) {
        for (int a = 0; a < p->P; ++a)
            for (size_t k = 0; k < (size_t)G * L.B; ++k) {
                const int64_t v = idx[(size_t)a * G * L.B + k];
                if (v < 0 || v >= p->cur_len[a]) return fail(SPP_ERR_ARG, "sample index outside [0, len(buffer))");
            }
        CK(p->d_idx.ensure(rows * sizeof(int64_t)));
        CK(cudaMemcpyAsync(p->d_idx.p, idx, rows * sizeof(int64_t), cudaMemcpyHostToDevice, s));
    }
    int rc = push_ring_len(p); if (rc) return rc;
    const bool sac = p->cfg.algo == SPP_ALGO_SAC;
    if (eps && sac) {
        CK(p->d_eps.ensure(rows * 2 * L.ob * 4));
        CK(cudaMemcpyAsync(p->d_eps.p, eps, rows * 2 * L.ob * 4, cudaMemcpyHostToDevice, s));
    }
    CK(p->d_losses.ensure((size_t)p->P * G * LOSS_COUNT * 4));
    UpdateArgs a;
    fill_args(p, a, G);
    a.idx = idx ? (const int64_t*)p->d_idx.p : nullptr;
    a.eps = (eps && sac) ? (const float*)p->d_eps.p : nullptr;
    a.seed = seed;
    a.losses = (float*)p->d_losses.p;
    CK(launch_update_balanced(p, a, s));
    g_launches++;
    if (losses) CK(cudaMemcpyAsync(losses, p->d_losses.p, (size_t)p->P * G * LOSS_COUNT * 4, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return SPP_OK;
}

int spp_update_ring_device(spp_population* p, int G, uint64_t seed, float* losses_dev, void* stream) {
    if (!p) return fail(SPP_ERR_ARG, "null population");
    if (p->S <= 0) return fail(SPP_ERR_STATE, "no replay ring");
    if (G < 1) return fail(SPP_ERR_ARG, "grad_steps must be positive");
    for (int a = 0; a < p->P; ++a)
        if (p->cur_len[a] < 1) return fail(SPP_ERR_STATE, "replay ring is empty");
    CK(cudaSetDevice(p->device));
    int rc = push_ring_len(p); if (rc) return rc;
    UpdateArgs a;
    fill_args(p, a, G);
    a.seed = seed;
    a.losses = losses_dev;
    CK(launch_update_balanced(p, a, stream ? (cudaStream_t)stream : p->stream));
    g_launches++;
    return SPP_OK;
}

// Stage timeline of one update burst (agent 0): the kernel stamps %globaltimer at its kStageMarks stage boundaries; out_us
// receives the mean duration of each of the kStageMarks - 1 intervals over the G steps, in microseconds.
int spp_update_stage_profile(spp_population* p, int G, uint64_t seed, double* out_us, int cap, int* n_out) {
    if (!p || !out_us) return fail(SPP_ERR_ARG, "null argument");
    if (p->S <= 0) return fail(SPP_ERR_STATE, "no replay ring");
    if (G < 1 || cap < kStageMarks - 1) return fail(SPP_ERR_ARG, "G must be positive and cap >= 23");
    for (int a = 0; a < p->P; ++a)
        if (p->cur_len[a] < 1) return fail(SPP_ERR_STATE, "replay ring is empty");
    CK(cudaSetDevice(p->device));
    int rc = push_ring_len(p); if (rc) return rc;
    const size_t n = (size_t)G * kStageMarks;
    CK(p->d_tmp.ensure(n * sizeof(unsigned long long)));
    UpdateArgs a;
    fill_args(p, a, G);
    a.seed = seed;
    a.timing = (unsigned long long*)p->d_tmp.p;
    CK(launch_update_balanced(p, a, p->stream));
    g_launches++;
    std::vector<unsigned long long> t(n);
    CK(cudaMemcpyAsync(t.data(), p->d_tmp.p, n * sizeof(unsigned long long), cudaMemcpyDeviceToHost, p->stream));
    CK(cudaStreamSynchronize(p->stream));
    for (int i = 0; i + 1 < kStageMarks; ++i) {
        double s = 0;
        for (int g = 0; g < G; ++g) s += (double)(t[(size_t)g * kStageMarks + i + 1] - t[(size_t)g * kStageMarks + i]);
        out_us[i] = s / G * 1e-3;
    }
    if (n_out) *n_out = kStageMarks - 1;
    return SPP_OK;
}

// ---- ACM regression bursts --------------------------------------------------------------------------
static void fill_acm_args(spp_population* p, UpdateArgs& a, int n) {
    fill_args(p, a, n);
    a.L = p->L_acm;
    a.scratch = p->scratch_acm;
}

static int acm_host_pass(spp_population* p, int n_batches, const float* x, const float* y, int last_rows, float* losses, int eval);

int spp_acm_update_host(spp_population* p, int n_batches, const float* x, const float* y, int last_rows, float* losses) {
    return acm_host_pass(p, n_batches, x, y, last_rows, losses, 0);
}

int spp_acm_eval_host(spp_population* p, int n_batches, const float* x, const float* y, int last_rows, float* losses) {
    return acm_host_pass(p, n_batches, x, y, last_rows, losses, 1);
}

static int acm_host_pass(spp_population* p, int n_batches, const float* x, const float* y, int last_rows, float* losses, int eval) {
    if (!p || !x || !y) return fail(SPP_ERR_ARG, "spp_acm_update_host / spp_acm_eval_host: null argument");
    if (n_batches < 1) return fail(SPP_ERR_ARG, "n_batches must be positive");
    CK(cudaSetDevice(p->device));
    const Layout& L = p->L_acm;
    const size_t rows = (size_t)p->P * n_batches * L.B;
    cudaStream_t s = p->stream;
    CK(p->d_obs.ensure(rows * 2 * L.ob * 4)); CK(p->d_aacm.ensure(rows * L.ac * 4));
    CK(cudaMemcpyAsync(p->d_obs.p, x, rows * 2 * L.ob * 4, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(p->d_aacm.p, y, rows * L.ac * 4, cudaMemcpyHostToDevice, s));
    CK(p->d_losses.ensure((size_t)p->P * n_batches * 4));
    UpdateArgs a;
    fill_acm_args(p, a, n_batches);
    if (last_rows < 0 || last_rows > L.B) return fail(SPP_ERR_ARG, "last_rows outside [0, acm_batch_size]");
    a.acm_last_rows = (last_rows == L.B) ? 0 : last_rows;
    a.acm_eval = eval;
    a.batch.acm_x = (const float*)p->d_obs.p;
    a.batch.acm_y = (const float*)p->d_aacm.p;
    a.losses = (float*)p->d_losses.p;
    CK(launch_acm_train(a, grid_for(p), s));
    g_launches++;
    if (losses) CK(cudaMemcpyAsync(losses, p->d_losses.p, (size_t)p->P * n_batches * 4, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return SPP_OK;
}

int spp_acm_update_ring(spp_population* p, int n_batches, const int64_t* idx, int last_rows, uint64_t seed, float* losses) {
    if (!p) return fail(SPP_ERR_ARG, "null population");
    if (p->S <= 0) return fail(SPP_ERR_STATE, "no replay ring");
    if (n_batches < 1) return fail(SPP_ERR_ARG, "n_batches must be positive");
    CK(cudaSetDevice(p->device));
    const Layout& L = p->L_acm;
    const size_t per = (size_t)n_batches * L.B, rows = (size_t)p->P * per;
    for (int a = 0; a < p->P; ++a)
        if (p->cur_len[a] < 1) return fail(SPP_ERR_STATE, "replay ring is empty");
    cudaStream_t s = p->stream;
    if (idx) {
        for (int a = 0; a < p->P; ++a)
            for (size_t k = 0; k < per; ++k) {
                if (last_rows > 0 && k >= per - L.B + (size_t)last_rows) break;      // unused tail of a partial last batch
                if (idx[(size_t)a * per + k] < 0 || idx[(size_t)a * per + k] >= p->cur_len[a])
                    return fail(SPP_ERR_ARG, "sample index outside [0, len(buffer))");
            }
        CK(p->d_idx.ensure(rows * sizeof(int64_t)));
        CK(cudaMemcpyAsync(p->d_idx.p, idx, rows * sizeof(int64_t), cudaMemcpyHostToDevice, s));
    }
    int rc = push_ring_len(p); if (rc) return rc;
    CK(p->d_losses.ensure((size_t)p->P * n_batches * 4));
    UpdateArgs a;
    fill_acm_args(p, a, n_batches);
    if (last_rows < 0 || last_rows > L.B) return fail(SPP_ERR_ARG, "last_rows outside [0, acm_batch_size]");
    a.acm_last_rows = (last_rows == L.B) ? 0 : last_rows;
    a.idx = idx ? (const int64_t*)p->d_idx.p : nullptr;
    a.seed = seed;
    a.losses = (float*)p->d_losses.p;
    CK(launch_acm_train(a, grid_for(p), s));
    g_launches++;
    if (losses) CK(cudaMemcpyAsync(losses, p->d_losses.p, (size_t)p->P * n_batches * 4, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return SPP_OK;
}

// ---- rollout -----------------------------------------------------------------------------------------
static int prepare_rollout(spp_population* p, int E, RolloutArgs& r) {
    if (E < 1 || E > 4096) return fail(SPP_ERR_ARG, "environments per agent must be in [1,4096]");
    if (p->roll_E != E) {
        p->L_roll = make_layout(p->cfg.algo, p->cfg.ob_dim, p->cfg.ac_dim, p->cfg.acm_kind, p->cfg.acm_critic ? 1 : 0, E);
        const size_t bytes = (size_t)p->P * p->L_roll.s.size * sizeof(float);
        CK(p->d_roll_scratch.ensure(bytes));
        CK(cudaMemsetAsync(p->d_roll_scratch.p, 0, bytes, p->stream));
        CK(p->d_env.ensure((size_t)p->P * E * p->L_roll.ldo * sizeof(float)));
        CK(cudaMemsetAsync(p->d_env.p, 0, (size_t)p->P * E * p->L_roll.ldo * sizeof(float), p->stream));
        p->roll_E = E;
    }
    memset(&r, 0, sizeof(r));
    fill_args(p, r.u, 1);
    r.u.L = p->L_roll;
    r.u.scratch = (float*)p->d_roll_scratch.p;
    r.steps = 1;
    return SPP_OK;
}

int spp_rollout_step_host(spp_population* p, int E, const float* obs, const float* noise, const float* eps, int random_phase,
                          double act_noise, int obs_norm, int denormalize_actor_out, float* out_target, float* out_action) {
    if (!p || !obs || !noise || !out_target || !out_action) return fail(SPP_ERR_ARG, "spp_rollout_step_host: null argument");
    if (random_phase < 0 || random_phase > 2) return fail(SPP_ERR_ARG, "random_phase must be 0 (actor), 1 (lim * noise) or 2 (noise IS the state target)");
    CK(cudaSetDevice(p->device));
    RolloutArgs r;
    int rc = prepare_rollout(p, E, r); if (rc) return rc;
    const Layout& L = p->L_roll;
    const size_t n_ob = (size_t)p->P * E * L.ob, n_ac = (size_t)p->P * E * L.ac;
    cudaStream_t s = p->stream;
    CK(p->d_obs.ensure(n_ob * 4)); CK(p->d_nobs.ensure(n_ob * 4)); CK(p->d_act.ensure(n_ob * 4)); CK(p->d_aacm.ensure(n_ac * 4));
    CK(cudaMemcpyAsync(p->d_obs.p, obs, n_ob * 4, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(p->d_nobs.p, noise, n_ob * 4, cudaMemcpyHostToDevice, s));
    if (eps) { CK(p->d_eps.ensure(n_ob * 4)); CK(cudaMemcpyAsync(p->d_eps.p, eps, n_ob * 4, cudaMemcpyHostToDevice, s)); }
    r.in_obs = (const float*)p->d_obs.p; r.in_noise = (const float*)p->d_nobs.p; r.in_eps = eps ? (const float*)p->d_eps.p : nullptr;
    r.out_target = (float*)p->d_act.p; r.out_action = (float*)p->d_aacm.p;
    r.random_phase = random_phase; r.obs_norm = obs_norm ? 1 : 0; r.denormalize_out = denormalize_actor_out ? 1 : 0;
    r.act_noise = (float)act_noise;
    CK(launch_rollout(r, grid_for(p), s));
    g_launches++;
    CK(cudaMemcpyAsync(out_target, r.out_target, n_ob * 4, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(out_action, r.out_action, n_ac * 4, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return SPP_OK;
}

int spp_rollout_synthetic_device(spp_population* p, int E, int steps, uint64_t seed, double act_noise, int random_phase, void* stream) {
    if (!p) return fail(SPP_ERR_ARG, "null population");
    if (random_phase < 0 || random_phase > 1) return fail(SPP_ERR_ARG, "random_phase must be 0 (actor + noise) or 1 (lim * N(0,1), frames < random_frames)");
    if (p->S <= 0) return fail(SPP_ERR_STATE, "no replay ring");
    if (steps < 1) return fail(SPP_ERR_ARG, "steps must be positive");
    if ((int64_t)(steps + 1) * E > p->S) return fail(SPP_ERR_ARG, "steps * E must fit the ring");
    CK(cudaSetDevice(p->device));
    RolloutArgs r;
    int rc = prepare_rollout(p, E, r); if (rc) return rc;
    cudaStream_t s = stream ? (cudaStream_t)stream : p->stream;
    CK(p->d_cur.ensure((size_t)p->P * 2 * sizeof(int64_t)));
    std::vector<int64_t> cur(2 * (size_t)p->P);
    for (int a = 0; a < p->P; ++a) { cur[a] = p->obs_cur[a]; cur[p->P + a] = p->ts_cur[a]; }
    CK(cudaMemcpyAsync(p->d_cur.p, cur.data(), cur.size() * sizeof(int64_t), cudaMemcpyHostToDevice, p->stream));
    CK(cudaStreamSynchronize(p->stream));
    r.env_state = (float*)p->d_env.p;
    r.w_obs = p->r_obs; r.w_oidx = p->r_oidx; r.w_nidx = p->r_nidx; r.w_act = p->r_act; r.w_rew = p->r_rew;
    r.w_done = p->r_done; r.w_end = p->r_end; r.w_aacm = p->r_aacm;
    r.obs_cur = (const int64_t*)p->d_cur.p; r.ts_cur = r.obs_cur + p->P;
    r.steps = steps; r.denormalize_out = 1; r.act_noise = (float)act_noise; r.u.seed = seed; r.random_phase = random_phase;
    CK(launch_rollout(r, grid_for(p), s));
    g_launches++;
    for (int a = 0; a < p->P; ++a) {   // host mirror of the cursors (no episode resets in the synthetic environment)
        const int64_t n = (int64_t)steps * E;
        p->obs_cur[a] = (p->obs_cur[a] + n) % p->S;
        const int64_t ts = p->ts_cur[a] + n;
        p->cur_len[a] = ts >= p->S ? p->S : (ts > p->cur_len[a] ? ts : p->cur_len[a]);
        p->ts_cur[a] = ts % p->S;
    }
    p->len_dirty = true;
    return SPP_OK;
}

// ---- introspection ---------------------------------------------------------------------------------
int spp_set_gemm_path(int tensor_cores) {
    if (tensor_cores < 0 || tensor_cores > 2) return fail(SPP_ERR_ARG, "gemm path must be 0 (FFMA), 1 (tcgen05, 3-pass tf32) or 2 (tcgen05, single tf32 pass)");
    g_gemm_path.store(tensor_cores);
    return SPP_OK;
}

int spp_debug_scratch(spp_population* p, int a, const char* name, float* host, int cap, int* rows, int* ld) {
    if (!p || !name || !host || a < 0 || a >= p->P) return fail(SPP_ERR_ARG, "bad argument");
    const Layout& L = p->L; const ScratchDesc& s = L.s;
    struct E { const char* n; int off, rows, ld; };
    const int B = L.B;
    const E table[] = {
        {"xo", s.xo, B, L.ldo}, {"xn", s.xn, B, L.ldo}, {"xc", s.xc, B, L.ldc}, {"xcp", s.xcp, B, L.ldc}, {"xm", s.xm, B, L.ldm},
        {"ha1", s.ha1, B, kHidden}, {"ha2", s.ha2, B, kHidden}, {"ml", s.ml, B, L.ldh}, {"zt", s.zt, B, L.ldo}, {"epsb", s.epsb, B, L.ldo},
        {"hc1_0", s.hc1[0], B, kHidden}, {"hc1_1", s.hc1[1], B, kHidden}, {"hc2_0", s.hc2[0], B, kHidden}, {"hc2_1", s.hc2[1], B, kHidden},
        {"dz2_0", s.dz2[0], B, kHidden}, {"dz2_1", s.dz2[1], B, kHidden}, {"dz1_0", s.dz1[0], B, kHidden}, {"dz1_1", s.dz1[1], B, kHidden},
        {"hm1", s.hm1, B, L.ldm1}, {"hm2", s.hm2, B, L.ldm2}, {"tm3", s.tm3, B, L.lda}, {"dm3", s.dm3, B, L.lda}, {"dm2", s.dm2, B, L.ldm2},
        {"dm1", s.dm1, B, L.ldm1}, {"dxc", s.dxc, B, L.ldc}, {"dxm", s.dxm, B, L.ldm}, {"dml", s.dml, B, L.ldh},
        {"dza2", s.dza2, B, kHidden}, {"dza1", s.dza1, B, kHidden}, {"vec", s.vec, 9, L.Bp}, {"gvec", s.gvec, 8, 512},
    };
    for (const E& e : table)
        if (strcmp(e.n, name) == 0) {
            if (cap < e.rows * e.ld) return fail(SPP_ERR_ARG, "host buffer too small");
            CK(cudaSetDevice(p->device));
            CK(cudaStreamSynchronize(p->stream));
            CK(cudaMemcpy(host, p->scratch + (size_t)a * s.size + e.off, (size_t)e.rows * e.ld * sizeof(float), cudaMemcpyDeviceToHost));
            if (rows) *rows = e.rows;
            if (ld) *ld = e.ld;
            return SPP_OK;
        }
    return fail(SPP_ERR_ARG, std::string("unknown scratch buffer ") + name);
}

}  // extern "C"
