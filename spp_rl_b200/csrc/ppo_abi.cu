// C ABI of the SPP-PPO path (include/spp_rl_b200.h, "spp_ppo_*"): one policy, data-parallel over rows.
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/spp_rl_b200.h"
#include "ppo_kernels.cuh"
#include "update_kernel.cuh"

using namespace spp;

extern "C" const char* spp_last_error(void);
int spp_set_error_(int code, const std::string& msg);     // defined in abi.cu
void spp_count_launch_();

#define PCK(expr)                                                                                   \
    do {                                                                                            \
        cudaError_t e_ = (expr);                                                                    \
        if (e_ != cudaSuccess)                                                                      \
            return spp_set_error_(SPP_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_)); \
    } while (0)

// ---- NCCL, bound at run time (dlopen): the library has no link-time dependency on it, so it loads on boxes without NCCL and
//      re-uses the copy torch already mapped when it is called from a torch.distributed process.  Only the five entry points the
//      data-parallel SPP-PPO path needs; enum values as in nccl.h (ncclFloat32 = 7, ncclFloat64 = 8, ncclSum = 0).
#include <dlfcn.h>
namespace {
struct NcclUniqueId { char internal[128]; };
typedef struct ncclComm* ncclComm_t;
struct NcclApi {
    int (*GetUniqueId)(NcclUniqueId*) = nullptr;
    int (*CommInitRank)(ncclComm_t*, int, NcclUniqueId, int) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    int (*GetVersion)(int*) = nullptr;
    bool ok = false;
    std::string err;
};
NcclApi& nccl() {
    static NcclApi api;
    static bool tried = false;
    if (tried) return api;
    tried = true;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);      // the copy already in the process (torch's), if any
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) { api.err = std::string("dlopen(libnccl.so.2): ") + dlerror(); return api; }
    api.GetUniqueId = (int (*)(NcclUniqueId*))dlsym(h, "ncclGetUniqueId");
    api.CommInitRank = (int (*)(ncclComm_t*, int, NcclUniqueId, int))dlsym(h, "ncclCommInitRank");
    api.AllReduce = (int (*)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t))dlsym(h, "ncclAllReduce");
    api.CommDestroy = (int (*)(ncclComm_t))dlsym(h, "ncclCommDestroy");
    api.GetErrorString = (const char* (*)(int))dlsym(h, "ncclGetErrorString");
    api.GetVersion = (int (*)(int*))dlsym(h, "ncclGetVersion");
    api.ok = api.GetUniqueId && api.CommInitRank && api.AllReduce && api.CommDestroy && api.GetErrorString;
    if (!api.ok) api.err = "libnccl.so.2 lacks an expected symbol";
    return api;
}
}  // namespace
#define NCK(expr)                                                                                       \
    do {                                                                                                \
        int r_ = (expr);                                                                                \
        if (r_ != 0) return spp_set_error_(SPP_ERR_CUDA, std::string(#expr) + ": " + nccl().GetErrorString(r_)); \
    } while (0)

#include "ppo_rollout.h"
#include "ppo_p2p.h"

struct PTensor { std::string name; int layer; int is_bias; int rows, cols; };

struct spp_ppo {
    spp_ppo_config cfg;
    int device = 0, sm_count = 0, grid = 0;
    int sm_use = 0;             // SMs the row-sharded kernels spread over: sm_count minus what spp_ppo_set_reserved_sms keeps free
    PpoLayout L;
    PpoHyper h;
    int64_t cap_rows = 0, cap_batch = 0;
    float *actor = nullptr, *actor_m = nullptr, *actor_v = nullptr, *critic = nullptr, *critic_m = nullptr, *critic_v = nullptr;
    float* norm = nullptr;
    PpoData d{};
    PpoBatch b{};
    PpoScratch s{};
    float *part = nullptr, *gbuf = nullptr, *scal = nullptr, *gscal = nullptr, *raw = nullptr;
    double* dstats = nullptr;
    int64_t* dperm = nullptr;
    int64_t* dperm_epoch = nullptr;      // one epoch's permutation (spp_ppo_update_actor), allocated on first use and kept
    int part_stride = 0;
    int step_actor = 0, step_critic = 0;
    ncclComm_t comm = nullptr;  // data-parallel runs (spp_ppo_comm_init): gradients + scalars are all-reduced inside the entry points
    int rank = 0, world = 1;
    int64_t n_allreduce = 0;    // collectives issued so far (reported by the benches)
    // NVLink peer-memory all-reduce (ppo_p2p.cu): exchange buffer [2][total] floats + flag words, mapped into every peer by CUDA IPC
    float* p2p_buf = nullptr; uint32_t* p2p_flags = nullptr; unsigned* p2p_ticket = nullptr; int* p2p_err = nullptr;
    float* p2p_peer[kP2pMaxRanks] = {}; uint32_t* p2p_peer_flags[kP2pMaxRanks] = {}; void* p2p_mapped[kP2pMaxRanks] = {};
    int p2p_on = 0; uint32_t p2p_epoch = 0; unsigned p2p_launches = 0; int64_t n_p2p = 0;
    int defer_reduce = 0, pend_parts = 0, pend_elems = 0;      // inside the library's own loops the local reduce is fused into the all-reduce kernel
    float* slog = nullptr;      // [kLogSlots][PS_COUNT + ldo]: per-step scalars (+ log_scale) recorded on the device, read once per loop
    // device rollout (spp_ppo_rollout_synthetic): persistent environment state and the extra store columns the ACM ring needs
    float* env_state = nullptr; int* env_len = nullptr; int env_E = 0;
    float* st_aacm = nullptr; float* st_raw_next = nullptr; int st_lda = 0;
    cudaEvent_t store_ev = nullptr;    // recorded behind the rollout kernel: consumers on other streams wait for it, not for the whole stream
    int store_E = 0, store_T = 0;      // shape of the [T][E] store the last device rollout left (0: rows came from spp_ppo_load_rollout)
    int plain_ppo = 0;          // 1: PPO.update_actor (custom_loss == 0): no distance term, log-prob of the stored actions as they are
    float* h_log = nullptr;     // pinned staging for the per-step scalar logs [kLogSlots][8 + ldo]
    double* h_stats = nullptr;  // pinned staging for the advantage statistics (a pageable copy serialises with the other streams' copies)
    int critic_tc = 1;          // critic fit on tcgen05 (ppo_critic_tc.cu) when the shapes allow; 0: the FFMA tile kernel (spp_ppo_set_critic_path)
    int a2c = 0;                // 1: A2C policy gradient -mean(logp * adv) (a2c.py:267-285, on_policy.py:100-124): no ratio, no entropy term
    float* gacc = nullptr;      // A2C_AcM.update_actor_acm never zeroes the actor's gradients: they accumulate here (on_policy.py:117-123)
    int64_t scratch_rows = 0;
    cudaStream_t stream = nullptr;
    std::vector<PTensor> tensors[2];
    std::vector<void*> allocs;
};

constexpr int kLogSlots = 4096;
// device log -> host through PINNED staging: a pageable cudaMemcpyAsync serialises with the other streams' work (measured: 13 ms per
// call beside the ACM burst of the population's stream, against 0.04 ms pinned)
static cudaError_t read_log(spp_ppo* p, float* dst, size_t floats);
static cudaError_t read_log(spp_ppo* p, float* dst, size_t floats) {
    cudaError_t e;
    if (!p->h_log && (e = cudaHostAlloc((void**)&p->h_log, (size_t)kLogSlots * (PS_COUNT + p->L.ldo) * 4, cudaHostAllocDefault)) != cudaSuccess) return e;
    if ((e = cudaMemcpyAsync(p->h_log, p->slog, floats * 4, cudaMemcpyDeviceToHost, p->stream)) != cudaSuccess) return e;
    if ((e = cudaStreamSynchronize(p->stream)) != cudaSuccess) return e;
    memcpy(dst, p->h_log, floats * 4);
    return cudaSuccess;
}
static const NetDesc& pnet(const spp_ppo* p, int net) { return net == 0 ? p->L.actor : p->L.critic; }

static int rows_per_cta(int64_t n, int grid) {
    int64_t r = (n + grid - 1) / grid;
    r = ((r + 127) / 128) * 128;
    return (int)(r < 128 ? 128 : r);
}

static void fill(const spp_ppo* p, PpoArgs& a, int64_t rows_for_chunks) {
    memset(&a, 0, sizeof(a));
    a.L = p->L; a.h = p->h; a.d = p->d; a.b = p->b; a.s = p->s;
    a.actor = p->actor; a.actor_m = p->actor_m; a.actor_v = p->actor_v;
    a.critic = p->critic; a.critic_m = p->critic_m; a.critic_v = p->critic_v;
    a.norm = p->norm; a.part = p->part; a.part_stride = p->part_stride; a.gbuf = p->gbuf; a.scal = p->scal; a.gscal = p->gscal;
    a.rows_per_cta = rows_per_cta(rows_for_chunks, p->grid);
    if (p->plain_ppo) a.h.custom_loss = 0.f;
    if (p->a2c) a.h.entropy_coef = 0.f;      // A2C has no entropy bonus: the Adam kernel adds none to d log_scale
    a.a2c = p->a2c;
}

int spp_ppo_store_view_(spp_ppo* p, PpoStoreView* out) {
    if (!p || !out) return spp_set_error_(SPP_ERR_ARG, "null argument");
    if (p->store_E < 1 || !p->st_aacm) return spp_set_error_(SPP_ERR_STATE, "the policy holds no device rollout (spp_ppo_rollout_synthetic)");
    out->device = p->device; out->E = p->store_E; out->T = p->store_T; out->ob = p->L.ob; out->ldo = p->L.ldo; out->lda = p->st_lda;
    out->raw_obs = p->raw; out->raw_next = p->st_raw_next; out->aacm = p->st_aacm; out->end = p->d.end; out->stream = p->stream; out->ready = p->store_ev;
    return SPP_OK;
}

extern "C" {

int spp_ppo_destroy(spp_ppo* p) {
    if (!p) return SPP_OK;
    cudaSetDevice(p->device);
    if (p->stream) cudaStreamSynchronize(p->stream);
    if (p->comm) { nccl().CommDestroy(p->comm); p->comm = nullptr; }
    for (int r = 0; r < kP2pMaxRanks; ++r) if (p->p2p_mapped[r]) cudaIpcCloseMemHandle(p->p2p_mapped[r]);
    if (p->h_stats) cudaFreeHost(p->h_stats);
    if (p->h_log) cudaFreeHost(p->h_log);
    if (p->p2p_buf) cudaFree(p->p2p_buf);
    if (p->p2p_ticket) cudaFree(p->p2p_ticket);
    for (void* q : {(void*)p->env_state, (void*)p->env_len, (void*)p->st_aacm, (void*)p->st_raw_next, (void*)p->dperm_epoch, (void*)p->gacc}) if (q) cudaFree(q);
    if (p->store_ev) cudaEventDestroy(p->store_ev);
    for (void* q : p->allocs) cudaFree(q);
    if (p->stream) cudaStreamDestroy(p->stream);
    delete p;
    return SPP_OK;
}

int spp_ppo_create(const spp_ppo_config* cfg, int device, spp_ppo** out) {
    if (!cfg || !out) return spp_set_error_(SPP_ERR_ARG, "spp_ppo_create: null argument");
    if (cfg->ob_dim < 1 || cfg->ob_dim > 128 || cfg->ac_dim < 1 || cfg->ac_dim > 32) return spp_set_error_(SPP_ERR_ARG, "bad dims");
    if (cfg->max_rows < 1 || cfg->max_batch_rows < 1) return spp_set_error_(SPP_ERR_ARG, "max_rows and max_batch_rows must be positive");
    int ndev = 0;
    PCK(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return spp_set_error_(SPP_ERR_ARG, "no such CUDA device");
    PCK(cudaSetDevice(device));
    cudaDeviceProp prop;
    PCK(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) return spp_set_error_(SPP_ERR_UNSUPPORTED, "spp_rl_b200 is built for sm_100a (B200) only");
    spp_ppo* p = new spp_ppo();
    p->cfg = *cfg; p->device = device; p->sm_count = prop.multiProcessorCount; p->sm_use = p->sm_count; p->grid = 2 * prop.multiProcessorCount;      // two resident CTAs per SM (kPpoCtasPerSm)
    p->L = make_ppo_layout(cfg->ob_dim, cfg->ac_dim);
    p->h.gamma = (float)cfg->gamma; p->h.discount = (float)(cfg->gae_lambda * cfg->gamma); p->h.discount_d = cfg->gae_lambda * cfg->gamma;
    p->h.epsilon = (float)cfg->ppo_epsilon; p->h.entropy_coef = (float)cfg->entropy_coef; p->h.custom_loss = (float)cfg->custom_loss;
    p->cap_rows = cfg->max_rows; p->cap_batch = cfg->max_batch_rows;
    const int ldo = p->L.ldo;
    auto alloc = [&](void** q, size_t bytes) -> cudaError_t {
        cudaError_t e = cudaMalloc(q, bytes ? bytes : 16);
        if (e == cudaSuccess) { e = cudaMemset(*q, 0, bytes ? bytes : 16); p->allocs.push_back(*q); }
        return e;
    };
#define PALLOC(ptr, bytes)                                                                           \
    do {                                                                                             \
        cudaError_t e_ = alloc((void**)&(ptr), (bytes));                                             \
        if (e_ != cudaSuccess) {                                                                     \
            std::string m = std::string("cudaMalloc(" #ptr "): ") + cudaGetErrorString(e_);          \
            spp_ppo_destroy(p);                                                                      \
            return spp_set_error_(SPP_ERR_CUDA, m);                                                  \
        }                                                                                            \
    } while (0)
    const size_t N = (size_t)p->cap_rows, NB = (size_t)p->cap_batch;
    PALLOC(p->actor, p->L.actor.size * 4); PALLOC(p->actor_m, p->L.actor.size * 4); PALLOC(p->actor_v, p->L.actor.size * 4);
    PALLOC(p->critic, p->L.critic.size * 4); PALLOC(p->critic_m, p->L.critic.size * 4); PALLOC(p->critic_v, p->L.critic.size * 4);
    PALLOC(p->norm, NORM_COUNT * ldo * 4);
    PALLOC(p->d.x, N * ldo * 4); PALLOC(p->d.xn, N * ldo * 4); PALLOC(p->d.act, N * ldo * 4); PALLOC(p->raw, N * ldo * 4);
    PALLOC(p->d.logp, N * 4); PALLOC(p->d.rew, N * 4); PALLOC(p->d.done, N * 4); PALLOC(p->d.end, N * 4);
    PALLOC(p->d.v, N * 4); PALLOC(p->d.nv, N * 4); PALLOC(p->d.q, N * 4); PALLOC(p->d.adv, N * 4);
    PALLOC(p->d.traj_start, N * 8); PALLOC(p->d.traj_len, N * 8);
    PALLOC(p->b.x, NB * ldo * 4); PALLOC(p->b.act, NB * ldo * 4); PALLOC(p->b.xn, NB * ldo * 4); PALLOC(p->b.logp, NB * 4); PALLOC(p->b.adv, NB * 4);
    p->scratch_rows = (int64_t)p->grid * rows_per_cta((int64_t)(N > NB ? N : NB), p->grid);
    const size_t SR = (size_t)p->scratch_rows;
    PALLOC(p->s.h1, SR * kPpoHidden * 4); PALLOC(p->s.h2, NB * kPpoHidden * 4 + 4096); PALLOC(p->s.dz2, SR * kPpoHidden * 4); PALLOC(p->s.dz1, SR * kPpoHidden * 4);
    PALLOC(p->s.mean, (NB + 128) * ldo * 4); PALLOC(p->s.t3, (NB + 128) * ldo * 4); PALLOC(p->s.d3, (NB + 128) * ldo * 4); PALLOC(p->s.newlogp, (NB + 128) * 4);
    p->part_stride = p->L.actor.size > p->L.critic.size ? p->L.actor.size : p->L.critic.size;
    PALLOC(p->part, (size_t)p->grid * p->part_stride * 4);
    // reduced gradient vector and the reduced scalars in ONE buffer: a data-parallel step all-reduces them with one collective
    PALLOC(p->gbuf, (size_t)(p->part_stride + PS_COUNT) * 4); p->gscal = p->gbuf + p->part_stride;
    PALLOC(p->scal, (size_t)p->grid * PS_COUNT * 4);
    PALLOC(p->dstats, (size_t)p->grid * 2 * 8);
    PALLOC(p->dperm, NB * 8);
    PALLOC(p->slog, (size_t)kLogSlots * (PS_COUNT + ldo) * 4);
#undef PALLOC
    if (cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking) != cudaSuccess) { spp_ppo_destroy(p); return spp_set_error_(SPP_ERR_CUDA, "stream"); }
    auto lin = [](std::vector<PTensor>& v, const char* nm, int layer, int rows, int cols) {
        v.push_back({std::string(nm) + ".weight", layer, 0, rows, cols});
        v.push_back({std::string(nm) + ".bias", layer, 1, rows, 1});
    };
    p->tensors[0].push_back({"log_scale", 3, 2, cfg->ob_dim, 1});       // state_dict order of basic_model.Actor
    lin(p->tensors[0], "fc1", 0, kPpoHidden, cfg->ob_dim); lin(p->tensors[0], "fc2", 1, kPpoHidden, kPpoHidden); lin(p->tensors[0], "fc3", 2, cfg->ob_dim, kPpoHidden);
    lin(p->tensors[1], "fc1", 0, kPpoHidden, cfg->ob_dim); lin(p->tensors[1], "fc2", 1, kPpoHidden, kPpoHidden); lin(p->tensors[1], "fc3", 2, 1, kPpoHidden);
    *out = p;
    std::vector<float> ones(128, 1.f);
    int rc = spp_ppo_set_limits(p, ones.data());
    if (rc == SPP_OK) rc = spp_ppo_set_norm_stats(p, nullptr, nullptr, nullptr, nullptr);
    if (rc != SPP_OK) { spp_ppo_destroy(p); *out = nullptr; }
    return rc;
}

int spp_ppo_set_limits(spp_ppo* p, const float* actor_lim) {
    if (!p || !actor_lim) return spp_set_error_(SPP_ERR_ARG, "null argument");
    PCK(cudaSetDevice(p->device));
    std::vector<float> v(p->L.ldo, 0.f);
    for (int j = 0; j < p->L.ob; ++j) v[j] = actor_lim[j];
    PCK(cudaMemcpy(p->norm + NORM_LIM * p->L.ldo, v.data(), p->L.ldo * 4, cudaMemcpyHostToDevice));
    return SPP_OK;
}

int spp_ppo_set_norm_stats(spp_ppo* p, const float* min_obs, const float* max_obs, const float* obs_mean, const float* obs_std) {
    if (!p) return spp_set_error_(SPP_ERR_ARG, "null argument");
    PCK(cudaSetDevice(p->device));
    const int ldo = p->L.ldo, ob = p->L.ob;
    std::vector<float> v(4 * (size_t)ldo, 0.f);
    float* doff = v.data(); float* dsc = doff + ldo; float* nsub = dsc + ldo; float* ndiv = nsub + ldo;
    for (int j = 0; j < ldo; ++j) { dsc[j] = (j < ob) ? 1.f : 0.f; ndiv[j] = 1.f; }
    if (p->cfg.min_max_denormalize) {
        if (min_obs && max_obs)
            for (int j = 0; j < ob; ++j) {
                const float mean = (max_obs[j] + min_obs[j]) / 2.f;
                doff[j] = mean; dsc[j] = (max_obs[j] - min_obs[j]) / 2.f; nsub[j] = mean; ndiv[j] = (max_obs[j] - mean) + 1e-8f;
            }
    } else if (obs_mean && obs_std) {
        for (int j = 0; j < ob; ++j) { doff[j] = obs_mean[j]; dsc[j] = obs_std[j] + 1e-8f; nsub[j] = obs_mean[j]; ndiv[j] = obs_std[j] + 1e-8f; }
    }
    PCK(cudaMemcpy(p->norm, v.data(), 4 * (size_t)ldo * 4, cudaMemcpyHostToDevice));
    return SPP_OK;
}

int spp_ppo_tensor_count(spp_ppo* p, int net) { return (p && (net == 0 || net == 1)) ? (int)p->tensors[net].size() : -1; }

int spp_ppo_tensor_info(spp_ppo* p, int net, int t, char* name, int name_cap, int* rows, int* cols) {
    if (!p || net < 0 || net > 1 || t < 0 || t >= (int)p->tensors[net].size()) return spp_set_error_(SPP_ERR_ARG, "bad net/tensor id");
    const PTensor& m = p->tensors[net][t];
    if (name && name_cap > 0) { strncpy(name, m.name.c_str(), name_cap - 1); name[name_cap - 1] = 0; }
    if (rows) *rows = m.rows;
    if (cols) *cols = m.cols;
    return SPP_OK;
}

// dir 0 upload, 1 download params, 2 download exp_avg, 3 download exp_avg_sq
static int ppo_tensor_io(spp_ppo* p, int net, int t, float* host, int dir) {
    if (!p || !host || net < 0 || net > 1 || t < 0 || t >= (int)p->tensors[net].size()) return spp_set_error_(SPP_ERR_ARG, "bad argument");
    PCK(cudaSetDevice(p->device));
    PCK(cudaStreamSynchronize(p->stream));
    const PTensor& m = p->tensors[net][t];
    const LayerDesc& l = pnet(p, net).L[m.layer];
    float* arena = net == 0 ? (dir == 2 ? p->actor_m : dir == 3 ? p->actor_v : p->actor) : (dir == 2 ? p->critic_m : dir == 3 ? p->critic_v : p->critic);
    if (m.is_bias == 0) {
        std::vector<float> nat((size_t)m.rows * l.ld, 0.f), tr;
        if (dir == 0) {
            for (int r = 0; r < m.rows; ++r) for (int c = 0; c < m.cols; ++c) nat[(size_t)r * l.ld + c] = host[(size_t)r * m.cols + c];
            PCK(cudaMemcpy(arena + l.off_w, nat.data(), nat.size() * 4, cudaMemcpyHostToDevice));
            if (l.off_wt >= 0) {
                tr.assign((size_t)l.ld * l.ld_t, 0.f);
                for (int r = 0; r < m.rows; ++r) for (int c = 0; c < m.cols; ++c) tr[(size_t)c * l.ld_t + r] = host[(size_t)r * m.cols + c];
                PCK(cudaMemcpy(arena + l.off_wt, tr.data(), tr.size() * 4, cudaMemcpyHostToDevice));
            }
        } else {
            PCK(cudaMemcpy(nat.data(), arena + l.off_w, nat.size() * 4, cudaMemcpyDeviceToHost));
            for (int r = 0; r < m.rows; ++r) for (int c = 0; c < m.cols; ++c) host[(size_t)r * m.cols + c] = nat[(size_t)r * l.ld + c];
        }
    } else {
        float* dev = arena + (m.is_bias == 1 ? l.off_b : l.off_w);
        if (dir == 0) PCK(cudaMemcpy(dev, host, m.rows * 4, cudaMemcpyHostToDevice));
        else PCK(cudaMemcpy(host, dev, m.rows * 4, cudaMemcpyDeviceToHost));
    }
    return SPP_OK;
}
int spp_ppo_params_upload(spp_ppo* p, int net, int t, const float* host) { return ppo_tensor_io(p, net, t, const_cast<float*>(host), 0); }
int spp_ppo_params_download(spp_ppo* p, int net, int t, float* host) { return ppo_tensor_io(p, net, t, host, 1); }
int spp_ppo_adam_download(spp_ppo* p, int net, int t, float* exp_avg, float* exp_avg_sq, int* step) {
    int rc = SPP_OK;
    if (exp_avg) rc = ppo_tensor_io(p, net, t, exp_avg, 2);
    if (rc == SPP_OK && exp_avg_sq) rc = ppo_tensor_io(p, net, t, exp_avg_sq, 3);
    if (rc == SPP_OK && step) *step = net == 0 ? p->step_actor : p->step_critic;
    return rc;
}

int spp_ppo_adam_reset(spp_ppo* p, int net) {      // fresh optimiser state (a new module assigned to model.actor / model.critic)
    if (!p || net < 0 || net > 1) return spp_set_error_(SPP_ERR_ARG, "bad argument");
    PCK(cudaSetDevice(p->device));
    PCK(cudaStreamSynchronize(p->stream));
    const size_t n = (size_t)pnet(p, net).size * 4;
    PCK(cudaMemset(net == 0 ? p->actor_m : p->critic_m, 0, n));
    PCK(cudaMemset(net == 0 ? p->actor_v : p->critic_v, 0, n));
    if (net == 0) p->step_actor = 0; else p->step_critic = 0;
    if (net == 0 && p->gacc) PCK(cudaMemset(p->gacc, 0, (size_t)p->part_stride * 4));      // a new module has no .grad yet
    return SPP_OK;
}

int spp_ppo_load_rollout(spp_ppo* p, int64_t N, const float* obs, const float* next_obs, const float* actions, const float* logp,
                         const float* rew, const float* done, const float* end, const int64_t* traj_start, const int64_t* traj_len,
                         int n_traj, int64_t traj_stride, int64_t global_rows) {
    if (!p || !obs || !next_obs || !actions || !logp || !rew || !done || !end || !traj_start || !traj_len)
        return spp_set_error_(SPP_ERR_ARG, "spp_ppo_load_rollout: null argument");
    if (N < 1 || N > p->cap_rows) return spp_set_error_(SPP_ERR_ARG, "N outside [1, max_rows]");
    if (n_traj < 1 || n_traj > N) return spp_set_error_(SPP_ERR_ARG, "bad trajectory count");
    PCK(cudaSetDevice(p->device));
    cudaStream_t s = p->stream;
    const int ob = p->L.ob, ldo = p->L.ldo;
    const int clamp = p->cfg.min_max_denormalize ? 0 : 1;
    p->d.N = N; p->d.Ntot = global_rows > 0 ? global_rows : N; p->d.n_traj = n_traj; p->d.traj_stride = traj_stride;
    p->store_E = p->store_T = 0;
    // Memory.norm_obs / norm_next_obs (rltoolkit/buffer/memory.py:170-176): normalise once on the device
    PCK(cudaMemcpyAsync(p->raw, obs, (size_t)N * ob * 4, cudaMemcpyHostToDevice, s));
    PCK(launch_ppo_normalize_rows(p->raw, p->d.x, N, ob, ldo, p->norm, clamp, p->grid * 4, s));
    PCK(cudaStreamSynchronize(s));
    PCK(cudaMemcpyAsync(p->raw, next_obs, (size_t)N * ob * 4, cudaMemcpyHostToDevice, s));
    PCK(launch_ppo_normalize_rows(p->raw, p->d.xn, N, ob, ldo, p->norm, clamp, p->grid * 4, s));
    PCK(cudaMemcpy2DAsync(p->d.act, ldo * 4, actions, ob * 4, ob * 4, N, cudaMemcpyHostToDevice, s));
    PCK(cudaMemcpyAsync(p->d.logp, logp, N * 4, cudaMemcpyHostToDevice, s));
    PCK(cudaMemcpyAsync(p->d.rew, rew, N * 4, cudaMemcpyHostToDevice, s));
    PCK(cudaMemcpyAsync(p->d.done, done, N * 4, cudaMemcpyHostToDevice, s));
    PCK(cudaMemcpyAsync(p->d.end, end, N * 4, cudaMemcpyHostToDevice, s));
    PCK(cudaMemcpyAsync(p->d.traj_start, traj_start, (size_t)n_traj * 8, cudaMemcpyHostToDevice, s));
    PCK(cudaMemcpyAsync(p->d.traj_len, traj_len, (size_t)n_traj * 8, cudaMemcpyHostToDevice, s));
    PCK(cudaStreamSynchronize(s));
    spp_count_launch_(); spp_count_launch_();
    return SPP_OK;
}

// ---- data parallelism: one NCCL communicator per policy, collectives on the policy's own stream ----------
int spp_comm_unique_id(char out[128]) {
    if (!out) return spp_set_error_(SPP_ERR_ARG, "null");
    if (!nccl().ok) return spp_set_error_(SPP_ERR_UNSUPPORTED, nccl().err);
    NcclUniqueId id;
    NCK(nccl().GetUniqueId(&id));
    memcpy(out, id.internal, 128);
    return SPP_OK;
}

int spp_ppo_comm_init(spp_ppo* p, const char id_bytes[128], int rank, int world) {
    if (!p || !id_bytes || world < 1 || rank < 0 || rank >= world) return spp_set_error_(SPP_ERR_ARG, "spp_ppo_comm_init: bad argument");
    if (!nccl().ok) return spp_set_error_(SPP_ERR_UNSUPPORTED, nccl().err);
    PCK(cudaSetDevice(p->device));
    if (p->comm) { nccl().CommDestroy(p->comm); p->comm = nullptr; }
    p->rank = rank; p->world = world;
    if (world == 1) return SPP_OK;
    NcclUniqueId id;
    memcpy(id.internal, id_bytes, 128);
    NCK(nccl().CommInitRank(&p->comm, world, id, rank));
    // the first collective of a communicator sets up its channels (seconds): pay for it here, not inside the first optimiser step
    PCK(cudaMemsetAsync(p->dstats, 0, 3 * sizeof(double), p->stream));
    NCK(nccl().AllReduce(p->dstats, p->dstats, 3, 8 /* ncclFloat64 */, 0, p->comm, p->stream));
    PCK(cudaStreamSynchronize(p->stream));
    return SPP_OK;
}

int spp_ppo_comm_info(spp_ppo* p, int* world, int64_t* allreduces, int* nccl_version) {
    if (!p) return spp_set_error_(SPP_ERR_ARG, "null");
    if (world) *world = p->comm ? p->world : 1;
    if (allreduces) *allreduces = p->n_allreduce;
    if (nccl_version) { *nccl_version = 0; if (nccl().ok && nccl().GetVersion) nccl().GetVersion(nccl_version); }
    return SPP_OK;
}

// local reduction of the gradient kernel's per-CTA slots -- or, inside the library's own data-parallel loops with the peer-memory path
// on, only a note of what to reduce: the all-reduce kernel does it on the way
static int reduce_local(spp_ppo* p, const PpoArgs& a, int n_part, int n_elems) {
    if (p->defer_reduce && (p->p2p_on || !p->comm)) { p->pend_parts = n_part; p->pend_elems = n_elems; return SPP_OK; }
    PCK(launch_ppo_reduce(a, n_part, n_elems, p->stream)); spp_count_launch_();
    return SPP_OK;
}

// gradient vector and the 8 reduced scalars sit in one buffer: ONE collective per optimiser step covers both
static int allreduce_grads(spp_ppo* p, const StepArgs* step = nullptr) {
    if (!p->comm) {
        if (p->pend_parts) {      // single GPU inside the library's loops: reduce (+ record + Adam when `step` is given) in one launch
            PpoArgs a; fill(p, a, 1);
            StepArgs t; memset(&t, 0, sizeof(t));
            if (step) t = *step;
            PCK(launch_ppo_reduce_step(a, p->pend_parts, p->pend_elems, t, p->stream)); spp_count_launch_();
            p->pend_parts = 0; p->pend_elems = 0;
        }
        return SPP_OK;
    }
    if (p->p2p_on) {      // fused reduce + all-reduce over NVLink peer memory (ppo_p2p.cu)
        P2pArgs x;
        memset(&x, 0, sizeof(x));
        if (step) x.step = *step;
        x.part = p->part; x.part_stride = p->part_stride; x.n_part = p->pend_parts; x.n_elems = p->pend_elems; x.scal = p->scal;
        x.gbuf = p->gbuf; x.total = p->part_stride + PS_COUNT;
        for (int r = 0; r < p->world; ++r) { x.peer[r] = p->p2p_peer[r]; x.peer_flags[r] = p->p2p_peer_flags[r]; }
        x.epoch = ++p->p2p_epoch; x.rank = p->rank; x.world = p->world;
        x.ticket = p->p2p_ticket; x.ticket_target = ++p->p2p_launches * (unsigned)((x.total + 255) / 256); x.err = p->p2p_err;
        PCK(launch_ppo_reduce_p2p(x, p->stream)); spp_count_launch_();
        p->pend_parts = 0; p->pend_elems = 0;
        p->n_allreduce++; p->n_p2p++;
        return SPP_OK;
    }
    NCK(nccl().AllReduce(p->gbuf, p->gbuf, (size_t)(p->part_stride + PS_COUNT), 7 /* ncclFloat32 */, 0 /* ncclSum */, p->comm, p->stream));
    p->n_allreduce++;
    return SPP_OK;
}

// ---- NVLink peer-memory exchange: every rank exports its buffer (CUDA IPC), the host side (any transport) carries the 64-byte handles ----
int spp_ppo_p2p_handle(spp_ppo* p, char out[64]) {
    if (!p || !out) return spp_set_error_(SPP_ERR_ARG, "null");
    PCK(cudaSetDevice(p->device));
    const size_t total = (size_t)p->part_stride + PS_COUNT;
    const size_t bytes = ((2 * total * 4 + 255) / 256) * 256 + 2 * kP2pMaxRanks * 4 + 256;
    if (!p->p2p_buf) {
        PCK(cudaMalloc(&p->p2p_buf, bytes));      // its own allocation: IPC exports whole allocations
        PCK(cudaMemset(p->p2p_buf, 0, bytes));
        p->p2p_flags = (uint32_t*)((char*)p->p2p_buf + ((2 * total * 4 + 255) / 256) * 256);
        PCK(cudaMalloc(&p->p2p_ticket, 256)); PCK(cudaMemset(p->p2p_ticket, 0, 256));
        p->p2p_err = (int*)(p->p2p_ticket + 16);
    }
    cudaIpcMemHandle_t h;
    PCK(cudaIpcGetMemHandle(&h, p->p2p_buf));
    static_assert(sizeof(h) == 64, "CUDA IPC handle size");
    memcpy(out, &h, 64);
    return SPP_OK;
}

int spp_ppo_p2p_init(spp_ppo* p, const char* handles, int rank, int world) {
    if (!p || !handles || world < 2 || world > kP2pMaxRanks || rank < 0 || rank >= world) return spp_set_error_(SPP_ERR_ARG, "spp_ppo_p2p_init: bad argument");
    if (!p->comm || p->rank != rank || p->world != world) return spp_set_error_(SPP_ERR_STATE, "spp_ppo_comm_init first (same rank / world)");
    if (!p->p2p_buf) return spp_set_error_(SPP_ERR_STATE, "spp_ppo_p2p_handle first");
    PCK(cudaSetDevice(p->device));
    const size_t total = (size_t)p->part_stride + PS_COUNT;
    const size_t flag_off = ((2 * total * 4 + 255) / 256) * 256;
    for (int r = 0; r < world; ++r) {
        void* base = p->p2p_buf;
        if (r != rank) {
            cudaIpcMemHandle_t h;
            memcpy(&h, handles + (size_t)r * 64, 64);
            cudaError_t e = cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess) {
                (void)cudaGetLastError();
                return spp_set_error_(SPP_ERR_UNSUPPORTED, std::string("cudaIpcOpenMemHandle: ") + cudaGetErrorString(e));
            }
            p->p2p_mapped[r] = base;
        }
        p->p2p_peer[r] = (float*)base;
        p->p2p_peer_flags[r] = (uint32_t*)((char*)base + flag_off);
    }
    p->p2p_on = 1;
    return SPP_OK;
}

int spp_ppo_p2p_info(spp_ppo* p, int* on, int64_t* steps, int* err) {
    if (!p) return spp_set_error_(SPP_ERR_ARG, "null");
    if (on) *on = p->p2p_on;
    if (steps) *steps = p->n_p2p;
    if (err) {
        *err = 0;
        if (p->p2p_err) { PCK(cudaSetDevice(p->device)); PCK(cudaStreamSynchronize(p->stream)); PCK(cudaMemcpy(err, p->p2p_err, 4, cudaMemcpyDeviceToHost)); }
    }
    return SPP_OK;
}

// (all-reduce +) record + Adam of one optimiser step: ONE launch when the fused paths apply (peer-memory all-reduce, or a single GPU),
// else the NCCL all-reduce followed by the record and Adam kernels
static int record_step(spp_ppo* p, int slot, bool with_log_scale);
static int finish_step(spp_ppo* p, int net, int slot, bool with_ls) {
    if (p->defer_reduce && (p->p2p_on || !p->comm) && (p->pend_parts > 0 || p->p2p_on)) {
        StepArgs t; memset(&t, 0, sizeof(t));
        const int step = net == 0 ? ++p->step_actor : ++p->step_critic;
        const double lr = net == 0 ? p->cfg.actor_lr : p->cfg.critic_lr;
        const double bc1 = 1.0 - pow(0.9, (double)step), bc2 = 1.0 - pow(0.999, (double)step);
        t.d = net == 0 ? p->L.actor : p->L.critic;
        t.W = net == 0 ? p->actor : p->critic; t.Mo = net == 0 ? p->actor_m : p->critic_m; t.Vo = net == 0 ? p->actor_v : p->critic_v;
        t.s = AdamScalars{(float)(lr / bc1), (float)sqrt(bc2)};
        t.extra_ls_grad = net == 0 ? -(p->a2c ? 0.f : p->h.entropy_coef) : 0.f; t.ls_layer = net == 0 ? 3 : -1;
        t.slog = p->slog + (size_t)slot * (PS_COUNT + p->L.ldo); t.ldo = p->L.ldo; t.with_ls = with_ls ? 1 : 0; t.enabled = 1;
        return allreduce_grads(p, &t);
    }
    int rc = allreduce_grads(p); if (rc) return rc;
    rc = record_step(p, slot, with_ls); if (rc) return rc;
    return net == 0 ? spp_ppo_actor_apply(p) : spp_ppo_critic_apply(p);
}

__global__ void ppo_record_kernel(const float* __restrict__ gscal, const float* __restrict__ log_scale, int ldo, float* __restrict__ slot) {
    const int i = threadIdx.x;
    if (i < PS_COUNT) slot[i] = gscal[i];
    else if (i < PS_COUNT + ldo) slot[i] = log_scale ? log_scale[i - PS_COUNT] : 0.f;
}
static int record_step(spp_ppo* p, int slot, bool with_log_scale) {
    ppo_record_kernel<<<1, 256, 0, p->stream>>>(p->gscal, with_log_scale ? p->actor + p->L.actor.L[3].off_w : nullptr, p->L.ldo,
                                                p->slog + (size_t)slot * (PS_COUNT + p->L.ldo));
    PCK(cudaGetLastError());
    return SPP_OK;
}

// ---- device-resident rollout (P1): E vectorised synthetic environments x T steps into the [T][E] store --------------------------
int spp_ppo_rollout_synthetic(spp_ppo* p, spp_population* pop, int agent, int E, int T, int max_ep_len, double done_prob, uint64_t seed,
                              int denormalize_actor_out, int reset_envs, const float* noise_act, const float* noise_env,
                              const float* u_done, const float* noise_reset) {
    if (!p || !pop) return spp_set_error_(SPP_ERR_ARG, "spp_ppo_rollout_synthetic: null argument");
    if (E < 1 || T < 1 || (int64_t)E * T > p->cap_rows) return spp_set_error_(SPP_ERR_ARG, "E * T outside [1, max_rows]");
    if (max_ep_len < 1 || done_prob < 0.0 || done_prob > 1.0) return spp_set_error_(SPP_ERR_ARG, "bad max_ep_len / done_prob");
    PopulationAcmView av;
    int rc = spp_population_acm_view_(pop, agent, &av); if (rc) return rc;
    if (av.device != p->device || av.ob != p->L.ob || av.ac != p->L.ac) return spp_set_error_(SPP_ERR_ARG, "population and policy disagree (device / shapes)");
    PCK(cudaSetDevice(p->device));
    cudaStream_t s = p->stream;
    const int ldo = p->L.ldo, ob = p->L.ob;
    const int64_t N = (int64_t)E * T;
    if (p->env_E != E) {      // (re)create the environments
        if (p->env_state) { cudaFree(p->env_state); cudaFree(p->env_len); p->env_state = nullptr; p->env_len = nullptr; }
        PCK(cudaMalloc(&p->env_state, (size_t)E * ldo * 4)); PCK(cudaMalloc(&p->env_len, (size_t)E * 4));
        p->env_E = E; reset_envs = 1;
    }
    if (reset_envs) { PCK(cudaMemsetAsync(p->env_state, 0, (size_t)E * ldo * 4, s)); PCK(cudaMemsetAsync(p->env_len, 0, (size_t)E * 4, s)); }
    if (!p->st_aacm || p->st_lda != av.lda) {
        if (p->st_aacm) { cudaFree(p->st_aacm); p->st_aacm = nullptr; }
        if (!p->st_raw_next) PCK(cudaMalloc(&p->st_raw_next, (size_t)p->cap_rows * ldo * 4));
        PCK(cudaMalloc(&p->st_aacm, (size_t)p->cap_rows * av.lda * 4));
        p->st_lda = av.lda;
    }
    PpoRolloutArgs a;
    memset(&a, 0, sizeof(a));
    a.L = p->L; a.actor = p->actor; a.norm = p->norm;
    a.acm = av.acm; a.acm_desc = av.acm_desc; a.acm_kind = av.acm_kind; a.ac = av.ac; a.lda = av.lda; a.hm1 = av.hm1; a.hm2 = av.hm2;
    a.ldm1 = av.ldm1; a.ldm2 = av.ldm2; a.acm_lim = av.acm_lim;
    a.E = E; a.T = T; a.max_ep_len = max_ep_len; a.done_prob = (float)done_prob;
    a.state = p->env_state; a.ep_len = p->env_len;
    a.x = p->d.x; a.xn = p->d.xn; a.act = p->d.act; a.logp = p->d.logp; a.rew = p->d.rew; a.done = p->d.done; a.end = p->d.end;
    a.aacm = p->st_aacm; a.raw_obs = p->raw; a.raw_next = p->st_raw_next;
    a.seed = seed; a.denorm_out = denormalize_actor_out ? 1 : 0; a.clamp = p->cfg.min_max_denormalize ? 0 : 1;
    {   // environments per CTA: as few as fill two CTAs per SM (the k-chains of a step are split over the spare threads)
        const int per = (E + 2 * p->sm_count - 1) / (2 * p->sm_count);
        a.rows_per_cta = per <= 2 ? 2 : per <= 4 ? 4 : per <= 8 ? 8 : per <= 16 ? 16 : per <= 24 ? 24 : 32;
    }
    float* tmp[4] = {nullptr, nullptr, nullptr, nullptr};
    const float* src[4] = {noise_act, noise_env, u_done, noise_reset};
    const size_t bytes[4] = {(size_t)N * ob * 4, (size_t)N * ob * 4, (size_t)N * 4, (size_t)N * ob * 4};
    for (int i = 0; i < 4; ++i)
        if (src[i]) { PCK(cudaMalloc(&tmp[i], bytes[i])); PCK(cudaMemcpyAsync(tmp[i], src[i], bytes[i], cudaMemcpyHostToDevice, s)); }
    a.noise_act = tmp[0]; a.noise_env = tmp[1]; a.u_done = tmp[2]; a.noise_reset = tmp[3];
    // trajectories of the store: environment e owns rows e, e + E, ... (T of them)
    std::vector<int64_t> ts(E), tl(E, T);
    for (int e = 0; e < E; ++e) ts[e] = e;
    PCK(cudaMemcpyAsync(p->d.traj_start, ts.data(), (size_t)E * 8, cudaMemcpyHostToDevice, s));
    PCK(cudaMemcpyAsync(p->d.traj_len, tl.data(), (size_t)E * 8, cudaMemcpyHostToDevice, s));
    PCK(cudaStreamSynchronize(s));      // ts / tl are host temporaries
    PCK(launch_ppo_rollout(a, s)); spp_count_launch_();
    if (!p->store_ev) PCK(cudaEventCreateWithFlags(&p->store_ev, cudaEventDisableTiming));
    PCK(cudaEventRecord(p->store_ev, s));
    p->d.N = N; p->d.Ntot = N; p->d.n_traj = E; p->d.traj_stride = E;
    p->store_E = E; p->store_T = T;
    bool any = false;
    for (int i = 0; i < 4; ++i) any = any || tmp[i];
    if (any) { PCK(cudaStreamSynchronize(s)); for (int i = 0; i < 4; ++i) if (tmp[i]) cudaFree(tmp[i]); }
    return SPP_OK;
}

// one column of the store as the reference's Memory would hold it: [N][ob] ("x", "xn", "act", "raw_obs", "raw_next"), [N][ac] ("aacm"),
// [N] ("logp", "rew", "done", "end", "adv", "v")
int spp_ppo_store_download(spp_ppo* p, const char* name, float* host) {
    if (!p || !name || !host) return spp_set_error_(SPP_ERR_ARG, "null argument");
    if (p->d.N < 1) return spp_set_error_(SPP_ERR_STATE, "no rollout loaded");
    PCK(cudaSetDevice(p->device));
    PCK(cudaStreamSynchronize(p->stream));
    const std::string n(name);
    const int64_t N = p->d.N;
    const int ob = p->L.ob, ldo = p->L.ldo;
    const float* m = nullptr; int w = 0, ld = 0;
    if (n == "x") { m = p->d.x; w = ob; ld = ldo; } else if (n == "xn") { m = p->d.xn; w = ob; ld = ldo; }
    else if (n == "act") { m = p->d.act; w = ob; ld = ldo; } else if (n == "raw_obs") { m = p->raw; w = ob; ld = ldo; }
    else if (n == "raw_next") { m = p->st_raw_next; w = ob; ld = ldo; } else if (n == "aacm") { m = p->st_aacm; w = p->L.ac; ld = p->st_lda; }
    else if (n == "logp") { m = p->d.logp; w = 1; ld = 1; } else if (n == "rew") { m = p->d.rew; w = 1; ld = 1; }
    else if (n == "done") { m = p->d.done; w = 1; ld = 1; } else if (n == "end") { m = p->d.end; w = 1; ld = 1; }
    else if (n == "adv") { m = p->d.adv; w = 1; ld = 1; } else if (n == "v") { m = p->d.v; w = 1; ld = 1; }
    if (!m) return spp_set_error_(SPP_ERR_ARG, "unknown (or not yet produced) store column: " + n);
    PCK(cudaMemcpy2D(host, (size_t)w * 4, m, (size_t)ld * 4, (size_t)w * 4, N, cudaMemcpyDeviceToHost));
    return SPP_OK;
}

// ---- critic ------------------------------------------------------------------------------------------
int spp_ppo_critic_targets(spp_ppo* p) {      // q = r + gamma (1 - done) V(next_obs), detached (a2c.py:203-206)
    if (!p || p->d.N < 1) return spp_set_error_(SPP_ERR_STATE, "no rollout loaded");
    PCK(cudaSetDevice(p->device));
    PpoArgs a; fill(p, a, p->d.N); a.mode = 1;
    PCK((p->critic_tc && ppo_critic_tc_supported(p->L.ob, p->L.ldo)) ? launch_ppo_critic_values_tc(a, p->grid, p->stream)
                                                                     : launch_ppo_critic_values(a, p->grid, p->stream)); spp_count_launch_();
    return SPP_OK;
}

int spp_ppo_critic_grad(spp_ppo* p) {         // local gradient of 0.5 mean((q - V)^2) -> reduced buffer; loss sum in gscal
    if (!p || p->d.N < 1) return spp_set_error_(SPP_ERR_STATE, "no rollout loaded");
    PCK(cudaSetDevice(p->device));
    PpoArgs a; fill(p, a, p->d.N);
    if (p->critic_tc && ppo_critic_tc_supported(p->L.ob, p->L.ldo)) {      // one persistent CTA per SM, rows in tiles of 128
        a.rows_per_cta = rows_per_cta(p->d.N, p->sm_use);
        PCK(launch_ppo_critic_grad_tc(a, p->sm_use, p->stream)); spp_count_launch_();
        return reduce_local(p, a, p->sm_use, p->L.critic.size);
    }
    PCK(launch_ppo_critic_grad(a, p->grid, p->stream)); spp_count_launch_();
    return reduce_local(p, a, p->grid, p->L.critic.size);
}

// The row-sharded kernels size their grids for whole SMs (one or two resident CTAs each).  A kernel of ANOTHER stream that holds an SM
// for long (the ACM regression burst of the population beside the policy: one 213 KB CTA) would push one CTA of every launch into a
// second wave -- twice the kernel time.  n > 0 leaves n SMs out of every grid (rows are re-split over the rest).
int spp_ppo_set_reserved_sms(spp_ppo* p, int n) {
    if (!p || n < 0 || n >= p->sm_count) return spp_set_error_(SPP_ERR_ARG, "bad argument");
    p->sm_use = p->sm_count - n;
    p->grid = 2 * p->sm_use;
    return SPP_OK;
}

int spp_ppo_set_critic_path(spp_ppo* p, int tensor_cores) {
    if (!p || tensor_cores < 0 || tensor_cores > 1) return spp_set_error_(SPP_ERR_ARG, "bad argument");
    p->critic_tc = tensor_cores;
    return SPP_OK;
}

int spp_ppo_critic_apply(spp_ppo* p) {
    if (!p) return spp_set_error_(SPP_ERR_ARG, "null");
    PCK(cudaSetDevice(p->device));
    PpoArgs a; fill(p, a, p->d.N);
    PCK(launch_ppo_adam(a, 1, ++p->step_critic, p->cfg.critic_lr, p->stream)); spp_count_launch_();
    return SPP_OK;
}

int spp_ppo_scalars(spp_ppo* p, float out[8]) {
    if (!p || !out) return spp_set_error_(SPP_ERR_ARG, "null");
    PCK(cudaSetDevice(p->device));
    PCK(cudaMemcpyAsync(out, p->gscal, PS_COUNT * 4, cudaMemcpyDeviceToHost, p->stream));
    PCK(cudaStreamSynchronize(p->stream));
    return SPP_OK;
}

int spp_ppo_update_critic(spp_ppo* p, int n_target_updates, int n_updates_per_target, float* mean_loss) {
    if (!p) return spp_set_error_(SPP_ERR_ARG, "null");
    const int steps = n_target_updates * n_updates_per_target;
    if (steps < 1 || steps > kLogSlots) return spp_set_error_(SPP_ERR_ARG, "critic steps outside [1, 4096]");
    // grad kernel -> reduce -> [all-reduce] -> record the loss sum -> Adam, all enqueued on the policy's stream; ONE host read at the end
    int k = 0;
    struct Defer { spp_ppo* p; Defer(spp_ppo* q) : p(q) { p->defer_reduce = 1; } ~Defer() { p->defer_reduce = 0; p->pend_parts = 0; } } defer(p);
    for (int t = 0; t < n_target_updates; ++t) {
        int rc = spp_ppo_critic_targets(p); if (rc) return rc;
        for (int u = 0; u < n_updates_per_target; ++u, ++k) {
            rc = spp_ppo_critic_grad(p); if (rc) return rc;
            rc = finish_step(p, 1, k, false); if (rc) return rc;
        }
    }
    const int w = PS_COUNT + p->L.ldo;
    std::vector<float> log((size_t)steps * w);
    PCK(read_log(p, log.data(), log.size()));
    double tot = 0.0;
    for (int i = 0; i < steps; ++i) tot += 0.5 * (double)log[(size_t)i * w + PS_LOSS] / (double)p->d.Ntot;      // a2c.py:208-216
    if (mean_loss) *mean_loss = (float)(tot / (double)steps);
    return SPP_OK;
}

// ---- advantages ------------------------------------------------------------------------------------------
int spp_ppo_advantages(spp_ppo* p, float* adv_host) {
    if (!p || p->d.N < 1) return spp_set_error_(SPP_ERR_STATE, "no rollout loaded");
    PCK(cudaSetDevice(p->device));
    PpoArgs a; fill(p, a, p->d.N); a.mode = 0;
    if (p->a2c) { a.h.discount = 0.f; a.h.discount_d = 0.0; }      // A2C.calculate_advantage (a2c.py:227-245): q - V(s), no GAE carry
    PCK((p->critic_tc && ppo_critic_tc_supported(p->L.ob, p->L.ldo)) ? launch_ppo_critic_values_tc(a, p->grid, p->stream)
                                                                     : launch_ppo_critic_values(a, p->grid, p->stream)); spp_count_launch_();
    PCK(launch_ppo_gae(a, p->stream)); spp_count_launch_();
    if (adv_host) PCK(cudaMemcpyAsync(adv_host, p->d.adv, p->d.N * 4, cudaMemcpyDeviceToHost, p->stream));
    PCK(cudaStreamSynchronize(p->stream));
    return SPP_OK;
}

int spp_ppo_load_advantages(spp_ppo* p, const float* adv_host) {      // advantages computed elsewhere (update_actor(advantages, buffer))
    if (!p || !adv_host || p->d.N < 1) return spp_set_error_(SPP_ERR_ARG, "null argument or no rollout loaded");
    PCK(cudaSetDevice(p->device));
    PCK(cudaMemcpyAsync(p->d.adv, adv_host, (size_t)p->d.N * 4, cudaMemcpyHostToDevice, p->stream));
    PCK(cudaStreamSynchronize(p->stream));
    return SPP_OK;
}

int spp_ppo_adv_stats(spp_ppo* p, double out[3]) {       // local (n, sum, sum of squares) in fp64
    if (!p || !out) return spp_set_error_(SPP_ERR_ARG, "null");
    PCK(cudaSetDevice(p->device));
    PpoArgs a; fill(p, a, p->d.N);
    PCK(launch_ppo_adv_stats(a, p->dstats, p->grid, p->stream)); spp_count_launch_();
    if (!p->h_stats) PCK(cudaHostAlloc((void**)&p->h_stats, (2 * (size_t)p->sm_count * 2 + 8) * 8, cudaHostAllocDefault));
    double* h = p->h_stats;
    PCK(cudaMemcpyAsync(h, p->dstats, 2 * (size_t)p->grid * 8, cudaMemcpyDeviceToHost, p->stream));
    PCK(cudaStreamSynchronize(p->stream));
    double s = 0, s2 = 0;
    for (int i = 0; i < p->grid; ++i) { s += h[2 * i]; s2 += h[2 * i + 1]; }
    out[0] = (double)p->d.N; out[1] = s; out[2] = s2;
    return SPP_OK;
}

// A2C_AcM.update_actor_acm (on_policy.py:117-123) steps the optimiser on gradients it never zeroes: accumulate != 0 adds the freshly
// reduced gradient (spp_ppo_actor_minibatch_grad) to the running sum and hands the sum to spp_ppo_actor_apply; 0 restarts the sum
// (A2C.update_actor, which does call zero_grad, a2c.py:282).
__global__ void ppo_grad_accumulate_kernel(float* __restrict__ gacc, float* __restrict__ gbuf, int n, int accumulate) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float g = accumulate ? __fadd_rn(gacc[i], gbuf[i]) : gbuf[i];
        gacc[i] = g; gbuf[i] = g;
    }
}
int spp_ppo_grad_accumulate(spp_ppo* p, int accumulate) {
    if (!p) return spp_set_error_(SPP_ERR_ARG, "null");
    PCK(cudaSetDevice(p->device));
    if (!p->gacc) {
        PCK(cudaMalloc(&p->gacc, (size_t)p->part_stride * 4));
        PCK(cudaMemsetAsync(p->gacc, 0, (size_t)p->part_stride * 4, p->stream));
    }
    ppo_grad_accumulate_kernel<<<(p->L.actor.size + 255) / 256, 256, 0, p->stream>>>(p->gacc, p->gbuf, p->L.actor.size, accumulate);
    PCK(cudaGetLastError()); spp_count_launch_();
    return SPP_OK;
}

// eps < 0: the PPO datasets' 1.2e-7 (advantage_dataset.py:10-12); A2C passes its own 1e-8 (a2c.py:275-277)
int spp_ppo_normalize_adv_eps(spp_ppo* p, const double* global_stats, double eps);
int spp_ppo_normalize_adv(spp_ppo* p, const double* global_stats) { return spp_ppo_normalize_adv_eps(p, global_stats, -1.0); }

int spp_ppo_normalize_adv_eps(spp_ppo* p, const double* global_stats, double eps) {   // (A - mean) / (std_unbiased + eps)
    if (!p) return spp_set_error_(SPP_ERR_ARG, "null");
    double st[3];
    if (global_stats) { st[0] = global_stats[0]; st[1] = global_stats[1]; st[2] = global_stats[2]; }
    else {
        int rc = spp_ppo_adv_stats(p, st); if (rc) return rc;
        if (p->comm) {      // (n, sum, sum of squares) over all ranks, in fp64 (torch.std parity needs it)
            double* hs = p->h_stats + 2 * (size_t)p->sm_count * 2;      // pinned (allocated by spp_ppo_adv_stats above)
            hs[0] = st[0]; hs[1] = st[1]; hs[2] = st[2];
            PCK(cudaMemcpyAsync(p->dstats, hs, 3 * sizeof(double), cudaMemcpyHostToDevice, p->stream));
            NCK(nccl().AllReduce(p->dstats, p->dstats, 3, 8 /* ncclFloat64 */, 0, p->comm, p->stream));
            p->n_allreduce++;
            PCK(cudaMemcpyAsync(hs, p->dstats, 3 * sizeof(double), cudaMemcpyDeviceToHost, p->stream));
            PCK(cudaStreamSynchronize(p->stream));
            st[0] = hs[0]; st[1] = hs[1]; st[2] = hs[2];
        }
    }
    const double n = st[0], mean = st[1] / n;
    const double var = n > 1 ? (st[2] - n * mean * mean) / (n - 1) : 0.0;
    const float denom = (float)std::sqrt(var > 0 ? var : 0.0) + (eps < 0 ? 1.2e-7f : (float)eps);
    PpoArgs a; fill(p, a, p->d.N);
    PCK(launch_ppo_adv_apply(a, (float)mean, denom, p->grid * 4, p->stream)); spp_count_launch_();
    return SPP_OK;
}

// ---- actor -----------------------------------------------------------------------------------------------
// A rank of a data-parallel run may own NO row of a global minibatch (a short final minibatch, a small ppo_batch_size): its
// contribution is a zero gradient vector and zero scalars -- it must still join the all-reduce and the Adam step.
static int empty_minibatch(spp_ppo* p, int64_t n_global) {
    PCK(cudaMemsetAsync(p->gbuf, 0, (size_t)(p->part_stride + PS_COUNT) * 4, p->stream));
    p->b.n = 0;
    p->b.n_mean = n_global > 0 ? n_global : 1;
    return SPP_OK;
}

int spp_ppo_actor_minibatch_grad(spp_ppo* p, const int64_t* perm, int64_t n, int64_t n_global) {
    if (!p || (!perm && n > 0)) return spp_set_error_(SPP_ERR_ARG, "null");
    if (n < 0 || n > p->cap_batch) return spp_set_error_(SPP_ERR_ARG, "minibatch outside [0, max_batch_rows]");
    PCK(cudaSetDevice(p->device));
    if (n == 0) return empty_minibatch(p, n_global);
    for (int64_t i = 0; i < n; ++i) if (perm[i] < 0 || perm[i] >= p->d.N) return spp_set_error_(SPP_ERR_ARG, "permutation index out of range");
    PCK(cudaMemcpyAsync(p->dperm, perm, (size_t)n * 8, cudaMemcpyHostToDevice, p->stream));
    p->b.n = n;
    p->b.n_mean = n_global > 0 ? n_global : n;      // the clipped loss is a mean over the GLOBAL minibatch in data-parallel runs
    PpoArgs a; fill(p, a, n);
    // A2C takes the log-prob of the stored (normalised-space) actions and denormalises only inside its distance term (on_policy.py:106-116)
    a.mode = (p->a2c && !p->cfg.norm_closs && !p->plain_ppo) ? 1 : 0;
    PCK(launch_ppo_gather(a, p->dperm, (p->cfg.norm_closs || p->plain_ppo || p->a2c) ? 0 : 1, p->grid * 4, p->stream)); spp_count_launch_();
    if (p->critic_tc && ppo_critic_tc_supported(p->L.ob, p->L.ldo)) {      // same tile machinery as the critic fit (ppo_critic_tc.cu)
        a.rows_per_cta = rows_per_cta(n, p->sm_use);
        PCK(launch_ppo_actor_grad_tc(a, p->sm_use, p->stream)); spp_count_launch_();
        return reduce_local(p, a, p->sm_use, p->L.actor.size);
    }
    PCK(launch_ppo_actor_grad(a, p->grid, p->stream)); spp_count_launch_();
    return reduce_local(p, a, p->grid, p->L.actor.size);
}

int spp_ppo_actor_minibatch_grad_device(spp_ppo* p, const int64_t* perm_dev, int64_t n, int64_t n_global) {
    if (!p || (!perm_dev && n > 0)) return spp_set_error_(SPP_ERR_ARG, "null");
    if (n < 0 || n > p->cap_batch) return spp_set_error_(SPP_ERR_ARG, "minibatch outside [0, max_batch_rows]");
    PCK(cudaSetDevice(p->device));
    if (n == 0) return empty_minibatch(p, n_global);
    p->b.n = n;
    p->b.n_mean = n_global > 0 ? n_global : n;
    PpoArgs a; fill(p, a, n);
    // A2C takes the log-prob of the stored (normalised-space) actions and denormalises only inside its distance term (on_policy.py:106-116)
    a.mode = (p->a2c && !p->cfg.norm_closs && !p->plain_ppo) ? 1 : 0;
    // the gather reads the caller's device ids in place (they were produced on this stream): no staging copy per minibatch
    PCK(launch_ppo_gather(a, perm_dev, (p->cfg.norm_closs || p->plain_ppo || p->a2c) ? 0 : 1, p->grid * 4, p->stream)); spp_count_launch_();
    if (p->critic_tc && ppo_critic_tc_supported(p->L.ob, p->L.ldo)) {      // same tile machinery as the critic fit (ppo_critic_tc.cu)
        a.rows_per_cta = rows_per_cta(n, p->sm_use);
        PCK(launch_ppo_actor_grad_tc(a, p->sm_use, p->stream)); spp_count_launch_();
        return reduce_local(p, a, p->sm_use, p->L.actor.size);
    }
    PCK(launch_ppo_actor_grad(a, p->grid, p->stream)); spp_count_launch_();
    return reduce_local(p, a, p->grid, p->L.actor.size);
}

int spp_ppo_actor_apply(spp_ppo* p) {
    if (!p) return spp_set_error_(SPP_ERR_ARG, "null");
    PCK(cudaSetDevice(p->device));
    PpoArgs a; fill(p, a, p->b.n);
    PCK(launch_ppo_adam(a, 0, ++p->step_actor, p->cfg.actor_lr, p->stream)); spp_count_launch_();
    return SPP_OK;
}

// One epoch of minibatch steps with the LOCAL row ids already on the device (ids_dev, a device pointer; minibatch k = ids[off[k], off[k + 1])
// of this rank, n_global[k] rows over all ranks).  Per step: gather -> grad -> reduce -> [all-reduce] -> record scalars + log_scale ->
// Adam, all on the policy's stream with no host synchronisation; the log (nb x (8 + ldo) floats) is read once at the end.
// log_host [nb][8 + ldo]: slots 0..7 = the reduced scalars of minibatch k, then log_scale as it was BEFORE that step.
int spp_ppo_actor_epoch_device(spp_ppo* p, const int64_t* ids_dev, const int64_t* off, const int64_t* n_global, int nb, float* log_host) {
    if (!p || !off || !log_host || nb < 1 || nb > kLogSlots) return spp_set_error_(SPP_ERR_ARG, "spp_ppo_actor_epoch_device: bad argument");
    if (p->d.N < 1) return spp_set_error_(SPP_ERR_STATE, "no rollout loaded");
    PCK(cudaSetDevice(p->device));
    struct Defer { spp_ppo* p; Defer(spp_ppo* q) : p(q) { p->defer_reduce = 1; } ~Defer() { p->defer_reduce = 0; p->pend_parts = 0; } } defer(p);
    for (int k = 0; k < nb; ++k) {
        const int64_t n = off[k + 1] - off[k];
        if (n < 0 || n > p->cap_batch) return spp_set_error_(SPP_ERR_ARG, "minibatch outside [0, max_batch_rows]");
        int rc = spp_ppo_actor_minibatch_grad_device(p, n > 0 ? ids_dev + off[k] : nullptr, n, n_global ? n_global[k] : 0); if (rc) return rc;
        rc = finish_step(p, 0, k, true); if (rc) return rc;
    }
    const int w = PS_COUNT + p->L.ldo;
    PCK(read_log(p, log_host, (size_t)nb * w));
    return SPP_OK;
}

int spp_ppo_update_actor(spp_ppo* p, const int64_t* perms, int max_epochs, int batch_size, double kl_threshold, float losses[4],
                         int* epochs_run, float* last_kl) {
    if (!p || !perms) return spp_set_error_(SPP_ERR_ARG, "null");
    if (p->d.N < 1) return spp_set_error_(SPP_ERR_STATE, "no rollout loaded");
    if (p->comm) return spp_set_error_(SPP_ERR_STATE, "data-parallel runs filter the permutation per rank: use spp_ppo_actor_epoch_device");
    if (batch_size < 1 || batch_size > p->cap_batch) return spp_set_error_(SPP_ERR_ARG, "batch_size outside [1, max_batch_rows]");
    const int64_t N = p->d.N;
    const int ob = p->L.ob, w = PS_COUNT + p->L.ldo;
    const int nb = (int)((N + batch_size - 1) / batch_size);
    if (nb > kLogSlots) return spp_set_error_(SPP_ERR_ARG, "more than 4096 minibatches per epoch");
    PCK(cudaSetDevice(p->device));
    for (int64_t i = 0; i < (int64_t)max_epochs * N; ++i)
        if (perms[i] < 0 || perms[i] >= N) return spp_set_error_(SPP_ERR_ARG, "permutation index out of range");
    if (!p->dperm_epoch) PCK(cudaMalloc(&p->dperm_epoch, (size_t)p->cap_rows * 8));
    int64_t* dperm_epoch = p->dperm_epoch;
    std::vector<int64_t> off(nb + 1), ng(nb);
    for (int k = 0; k <= nb; ++k) off[k] = (int64_t)k * batch_size < N ? (int64_t)k * batch_size : N;
    for (int k = 0; k < nb; ++k) ng[k] = off[k + 1] - off[k];
    std::vector<float> log((size_t)nb * w);
    double tot[4] = {0, 0, 0, 0};
    double kl = 0.0;
    int i = 0, ran = 0, rc = SPP_OK;
    for (i = 0; i < max_epochs; ++i) {
        if (kl >= kl_threshold) break;
        cudaError_t e = cudaMemcpyAsync(dperm_epoch, perms + (size_t)i * N, (size_t)N * 8, cudaMemcpyHostToDevice, p->stream);
        if (e != cudaSuccess) return spp_set_error_(SPP_ERR_CUDA, cudaGetErrorString(e));
        rc = spp_ppo_actor_epoch_device(p, dperm_epoch, off.data(), ng.data(), nb, log.data());
        if (rc) return rc;
        for (int k = 0; k < nb; ++k) {
            const float* sc = log.data() + (size_t)k * w;
            const double n = (double)ng[k];
            // state-independent entropy: sum_j 0.5 + 0.5 log(2 pi) + log_scale_j, with log_scale as it was in that minibatch
            float ent = 0.f;
            for (int j = 0; j < ob; ++j) ent += 0.5f + 0.918938533204672741780329736406f + sc[PS_COUNT + j];
            const double actor_loss = (double)sc[PS_LOSS] / n;
            const double dist = (p->h.custom_loss != 0.f && !p->plain_ppo) ? (double)sc[PS_DIST] / (n * ob) : 0.0;
            tot[0] += actor_loss; tot[1] += ent; tot[3] += dist;
            tot[2] += actor_loss - (double)p->h.entropy_coef * ent + (p->plain_ppo ? 0.0 : (double)p->h.custom_loss * dist);
            kl = (double)sc[PS_KL] / n;      // after the loop: the LAST (possibly short) minibatch of the epoch (quirk 16)
        }
        ++ran;
    }
    // the reference divides by (i + 1) with i the loop variable at exit (one more than the epochs run after an early stop)
    // (PPO_AcM.update_actor_acm, on_policy.py:211-214); plain PPO.update_actor reports the raw sums (ppo.py:186-188)
    const double div = p->plain_ppo ? 1.0 : (double)((i < max_epochs ? i : max_epochs - 1) + 1);
    if (losses) for (int k = 0; k < 4; ++k) losses[k] = (float)(tot[k] / div);
    if (epochs_run) *epochs_run = ran;
    if (last_kl) *last_kl = (float)kl;
    return SPP_OK;
}

int spp_ppo_act(spp_ppo* p, int64_t E, const float* obs, const float* noise, int denormalize_actor_out, float* action, float* logp,
                float* target) {
    if (!p || !obs || !noise || !action || !logp || !target) return spp_set_error_(SPP_ERR_ARG, "spp_ppo_act: null argument");
    if (E < 1 || E > p->cap_batch) return spp_set_error_(SPP_ERR_ARG, "E outside [1, max_batch_rows]");
    PCK(cudaSetDevice(p->device));
    cudaStream_t s = p->stream;
    const int ob = p->L.ob, ldo = p->L.ldo;
    PCK(cudaMemcpyAsync(p->raw, obs, (size_t)E * ob * 4, cudaMemcpyHostToDevice, s));
    PCK(launch_ppo_normalize_rows(p->raw, p->b.x, E, ob, ldo, p->norm, p->cfg.min_max_denormalize ? 0 : 1, p->grid * 4, s)); spp_count_launch_();
    PCK(cudaMemcpy2DAsync(p->s.d3, ldo * 4, noise, ob * 4, ob * 4, E, cudaMemcpyHostToDevice, s));
    p->b.n = E; p->b.n_mean = E;
    PpoArgs a; fill(p, a, E); a.mode = denormalize_actor_out ? 1 : 0;
    PCK(launch_ppo_act(a, p->grid, s)); spp_count_launch_();
    PCK(cudaMemcpy2DAsync(action, ob * 4, p->b.act, ldo * 4, ob * 4, E, cudaMemcpyDeviceToHost, s));
    PCK(cudaMemcpy2DAsync(target, ob * 4, p->b.xn, ldo * 4, ob * 4, E, cudaMemcpyDeviceToHost, s));
    PCK(cudaMemcpyAsync(logp, p->b.logp, (size_t)E * 4, cudaMemcpyDeviceToHost, s));
    PCK(cudaStreamSynchronize(s));
    return SPP_OK;
}

int spp_ppo_grad_buffer(spp_ppo* p, void** dev_ptr, int* n_floats, void** scal_ptr) {
    if (!p) return spp_set_error_(SPP_ERR_ARG, "null");
    if (dev_ptr) *dev_ptr = p->gbuf;
    if (n_floats) *n_floats = p->part_stride;
    if (scal_ptr) *scal_ptr = p->gscal;
    return SPP_OK;
}

int spp_ppo_set_actor_mode(spp_ppo* p, int mode) {      // 0: PPO_AcM.update_actor_acm, 1: plain PPO.update_actor, 2 / 3: A2C (with / without the distance term reported)
    if (!p || mode < 0 || mode > 3) return spp_set_error_(SPP_ERR_ARG, "bad argument");
    p->plain_ppo = (mode == 1 || mode == 3) ? 1 : 0;
    p->a2c = mode >= 2 ? 1 : 0;
    return SPP_OK;
}

int spp_ppo_set_global_rows(spp_ppo* p, int64_t global_rows) {
    if (!p || global_rows < 1) return spp_set_error_(SPP_ERR_ARG, "bad argument");
    p->d.Ntot = global_rows;
    return SPP_OK;
}

int spp_ppo_sync(spp_ppo* p) {
    if (!p) return spp_set_error_(SPP_ERR_ARG, "null");
    PCK(cudaSetDevice(p->device));
    PCK(cudaStreamSynchronize(p->stream));
    return SPP_OK;
}

int spp_ppo_stream(spp_ppo* p, void** stream) {
    if (!p || !stream) return spp_set_error_(SPP_ERR_ARG, "null");
    *stream = (void*)p->stream;
    return SPP_OK;
}

}  // extern "C"
