// ramp_policy.cu -- the GNN policy forward on the device (include/ramp_b200.h: ramp_policy_*).
//
// Restates GNNPolicy.forward (ml_models/policies/gnn_policy.py:137-296) for the observation RampJobPartitioningEnvironment
// emits.  Two kernels, both fp32:
//
//   ramp_gnn_embed_kernel   one CTA per job type.  num_rounds x MeanPool (ml_models/models/mean_pool.py:107-150): node module
//                           LN -> Linear -> act per node, edge module per edge, then per destination node the mean over
//                           [own (node | zeros) state, incoming (src node | edge) messages] of reduce module LN -> Linear -> act;
//                           a node with no incoming edge ends a round with zeros (DGL update_all).  Then the mean over the nodes
//                           (gnn_policy.py:262-268).  A job type's node / edge features never change (observation.py:503-567), so
//                           this runs once per weight set, not per decision.
//   ramp_policy_head_kernel persistent CTAs, one warp per episode: graph module LN -> Linear over [graph features | action mask]
//                           (gnn_policy.py:96-109, 271), concat with the model's node-mean embedding, the RLlib fully-connected
//                           read-out (one hidden layer, logits; separate value branch), + max(log(mask), FLT_MIN) on the logits
//                           (gnn_policy.py:283-290), then greedy / categorical action selection written straight into the
//                           environment's action buffer.  All read-out weights are staged once per CTA in shared memory, laid
//                           out so that lanes read consecutive words.
#include <cfloat>
#include <cstdarg>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/ramp_b200.h"

int ramp_internal_set_error(int code, const char* msg);
cudaStream_t ramp_internal_stream(ramp_engine_t* e);
int ramp_internal_device(ramp_engine_t* e);
void ramp_internal_count_launches(ramp_engine_t* e, int n);

namespace ramp {

constexpr int POL_MAX_ROUNDS = 8;
constexpr int POL_MAX_DIM = 128;        // node / message / embedding widths
constexpr int POL_MAX_HPL = 16;         // read-out hidden units per lane (hidden <= 512)
constexpr float LN_EPS = 1e-5f;         // torch.nn.LayerNorm default

static int perr(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    return ramp_internal_set_error(code, buf);
}

#define PCUDA(expr)                                                                                              \
    do {                                                                                                         \
        cudaError_t _e = (expr);                                                                                 \
        if (_e != cudaSuccess) return perr(RAMP_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)

struct RoundW {                          // offsets (in floats) into the weight blob
    int32_t in, out;
    int64_t nln_w, nln_b, nW, nb, eln_w, eln_b, eW, eb, rln_w, rln_b, rW, rb;
};

struct PolicyDev {
    ramp_policy_config_t c;
    RoundW rounds[POL_MAX_ROUNDS];
    int64_t gln_w, gln_b, gW, gb, hW, hb, lW, lb, vhW, vhb, vW, vb;
    const float* w;                      // the blob
};

struct ModelDev {
    int32_t n_nodes, n_edges;
    const float* nf; const float* ef;    // [N][in_node], [E][in_edge]
    const int32_t* in_ptr; const int32_t* in_edge; const int32_t* in_src;   // CSR by destination
    float* z0; float* z1;                // [N][POL_MAX_DIM]
    float* hn; float* he;                // [N][msg/2], [E][msg/2]
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ float act_fn(float x, int kind) {
    if (kind == 0) return fmaxf(x, 0.f);
    if (kind == 1) return x > 0.f ? x : 0.01f * x;
    return tanhf(x);
}

// LayerNorm of the n values a warp holds lane-strided in `buf` (shared, per warp), in place: biased variance, eps inside the root
__device__ __forceinline__ void warp_layer_norm(float* buf, int n, const float* w, const float* b, int lane) {
    float s = 0.f;
    for (int k = lane; k < n; k += 32) s += buf[k];
    const float mean = warp_sum(s) / (float)n;
    float q = 0.f;
    for (int k = lane; k < n; k += 32) { const float d = buf[k] - mean; q += d * d; }
    const float rstd = 1.0f / sqrtf(warp_sum(q) / (float)n + LN_EPS);
    for (int k = lane; k < n; k += 32) buf[k] = (buf[k] - mean) * rstd * w[k] + b[k];
    __syncwarp();
}

__global__ void __launch_bounds__(256) ramp_gnn_embed_kernel(const PolicyDev P, const ModelDev* models, float* emb) {
    __shared__ float sbuf[8][POL_MAX_DIM];
    const ModelDev M = models[blockIdx.x];
    if (M.n_nodes <= 0) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, n_warps = blockDim.x >> 5;
    float* buf = sbuf[warp];
    const int half = P.c.out_features_msg / 2, msg = P.c.out_features_msg, akind = P.c.aggregator_activation;
    const float* zin = M.nf;
    int zin_stride = P.c.in_features_node;
    float* zout = M.z0;
    for (int r = 0; r < P.c.num_rounds; ++r) {
        const RoundW R = P.rounds[r];
        const float* w = P.w;
        // ---- node module on every node, edge module on every edge (mean_pool.py:120-127) ----
        for (int v = warp; v < M.n_nodes; v += n_warps) {
            for (int k = lane; k < R.in; k += 32) buf[k] = zin[(size_t)v * zin_stride + k];
            __syncwarp();
            warp_layer_norm(buf, R.in, w + R.nln_w, w + R.nln_b, lane);
            for (int o = lane; o < half; o += 32) {
                float a = w[R.nb + o];
                const float* row = w + R.nW + (size_t)o * R.in;
                for (int k = 0; k < R.in; ++k) a += row[k] * buf[k];
                M.hn[(size_t)v * half + o] = act_fn(a, akind);
            }
            __syncwarp();
        }
        const int ine = P.c.in_features_edge;
        for (int e = warp; e < M.n_edges; e += n_warps) {
            for (int k = lane; k < ine; k += 32) buf[k] = M.ef[(size_t)e * ine + k];
            __syncwarp();
            warp_layer_norm(buf, ine, w + R.eln_w, w + R.eln_b, lane);
            for (int o = lane; o < half; o += 32) {
                float a = w[R.eb + o];
                const float* row = w + R.eW + (size_t)o * ine;
                for (int k = 0; k < ine; ++k) a += row[k] * buf[k];
                M.he[(size_t)e * half + o] = act_fn(a, akind);
            }
            __syncwarp();
        }
        __syncthreads();
        // ---- per destination node: mean of reduce_module over [own state, messages] (mean_pool.py:129-150) ----
        for (int v = warp; v < M.n_nodes; v += n_warps) {
            const int e0 = M.in_ptr[v], e1 = M.in_ptr[v + 1];
            float acc[POL_MAX_DIM / 32];
#pragma unroll
            for (int i = 0; i < POL_MAX_DIM / 32; ++i) acc[i] = 0.f;
            if (e1 > e0) {                                          // DGL leaves zero-in-degree nodes at zero
                for (int mi = -1; mi < e1 - e0; ++mi) {
                    const int src = mi < 0 ? v : M.in_src[e0 + mi];
                    for (int k = lane; k < half; k += 32) {
                        buf[k] = M.hn[(size_t)src * half + k];
                        buf[half + k] = mi < 0 ? 0.f : M.he[(size_t)M.in_edge[e0 + mi] * half + k];
                    }
                    __syncwarp();
                    warp_layer_norm(buf, msg, w + R.rln_w, w + R.rln_b, lane);
#pragma unroll
                    for (int i = 0; i < POL_MAX_DIM / 32; ++i) {
                        const int o = lane + 32 * i;
                        if (o < R.out) {
                            float a = w[R.rb + o];
                            const float* row = w + R.rW + (size_t)o * msg;
                            for (int k = 0; k < msg; ++k) a += row[k] * buf[k];
                            acc[i] += act_fn(a, akind);
                        }
                    }
                    __syncwarp();
                }
            }
            const float inv = 1.0f / (float)(e1 - e0 + 1);
#pragma unroll
            for (int i = 0; i < POL_MAX_DIM / 32; ++i) {
                const int o = lane + 32 * i;
                if (o < R.out) zout[(size_t)v * POL_MAX_DIM + o] = acc[i] * inv;
            }
        }
        __syncthreads();
        zin = zout; zin_stride = POL_MAX_DIM;
        zout = (zout == M.z0) ? M.z1 : M.z0;
    }
    // ---- mean over the job's nodes (gnn_policy.py:262-268) ----
    const int od = P.c.out_features_node;
    for (int o = threadIdx.x; o < od; o += blockDim.x) {
        float s = 0.f;
        for (int v = 0; v < M.n_nodes; ++v) s += zin[(size_t)v * POL_MAX_DIM + o];
        emb[(size_t)blockIdx.x * od + o] = s / (float)M.n_nodes;
    }
}

struct HeadArgs {
    int32_t n;                            // decisions
    // inputs: either full graph features (host-style forward) or the environment's buffers
    const float* graph_features;          // [n][in_graph] or nullptr
    const float* obs_dyn;                 // [n][11]
    const float* graph_static;            // [n_models][6]
    const int32_t* model;                 // [n]
    const uint8_t* done;                  // [n] or nullptr
    const uint8_t* mask;                  // [n][A]
    const float* emb;                     // [n_models][out_node]
    float* logits; float* value; float* logp; int32_t* actions;   // outputs (logits / value / logp may be nullptr)
    int32_t sample;
    unsigned long long seed;
};

__host__ __device__ inline size_t head_smem_floats(const ramp_policy_config_t& c, int warps) {
    const int gin = c.in_features_graph + c.n_actions, fin = c.out_features_node + c.out_features_graph, H = c.fcnet_hidden, A = c.n_actions;
    return (size_t)2 * gin + (size_t)c.out_features_graph * gin + c.out_features_graph      // graph module
           + (size_t)2 * fin * H + 2 * (size_t)H                                           // hidden layers (policy, value), transposed
           + (size_t)A * H + A + H + 1                                                      // logits, value
           + (size_t)warps * 2 * POL_MAX_DIM;                                               // per-warp staging
}

__device__ __forceinline__ unsigned long long splitmix64(unsigned long long x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

__global__ void __launch_bounds__(256) ramp_policy_head_kernel(const PolicyDev P, const HeadArgs a) {
    extern __shared__ float sm[];
    const ramp_policy_config_t& c = P.c;
    const int gin = c.in_features_graph + c.n_actions, og = c.out_features_graph, on = c.out_features_node, fin = on + og;
    const int H = c.fcnet_hidden, A = c.n_actions, hpl = H / 32;
    float* s_gln_w = sm;                 float* s_gln_b = s_gln_w + gin;
    float* s_gW = s_gln_b + gin;         float* s_gb = s_gW + (size_t)og * gin;
    float* s_hWt = s_gb + og;            float* s_hb = s_hWt + (size_t)fin * H;      // [fin][H]: lanes read consecutive hidden units
    float* s_vhWt = s_hb + H;            float* s_vhb = s_vhWt + (size_t)fin * H;
    float* s_lW = s_vhb + H;             float* s_lb = s_lW + (size_t)A * H;         // [A][H]
    float* s_vW = s_lb + A;              float* s_vb = s_vW + H;
    float* s_warp = s_vb + 1;
    const float* w = P.w;
    for (int i = threadIdx.x; i < gin; i += blockDim.x) { s_gln_w[i] = w[P.gln_w + i]; s_gln_b[i] = w[P.gln_b + i]; }
    for (int i = threadIdx.x; i < og * gin; i += blockDim.x) s_gW[i] = w[P.gW + i];
    for (int i = threadIdx.x; i < og; i += blockDim.x) s_gb[i] = w[P.gb + i];
    for (int i = threadIdx.x; i < fin * H; i += blockDim.x) {
        const int j = i / fin, k = i - j * fin;                       // blob is [H][fin]
        s_hWt[(size_t)k * H + j] = w[P.hW + i];
        s_vhWt[(size_t)k * H + j] = w[P.vhW + i];
    }
    for (int i = threadIdx.x; i < H; i += blockDim.x) { s_hb[i] = w[P.hb + i]; s_vhb[i] = w[P.vhb + i]; s_vW[i] = w[P.vW + i]; }
    for (int i = threadIdx.x; i < A * H; i += blockDim.x) s_lW[i] = w[P.lW + i];
    for (int i = threadIdx.x; i < A; i += blockDim.x) s_lb[i] = w[P.lb + i];
    if (threadIdx.x == 0) s_vb[0] = w[P.vb];
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpc = blockDim.x >> 5;
    float* xb = s_warp + (size_t)warp * 2 * POL_MAX_DIM;                // graph-module input, then the read-out input
    float* fb = xb + POL_MAX_DIM;
    for (int b = blockIdx.x * wpc + warp; b < a.n; b += gridDim.x * wpc) {
        const int m = a.model[b];
        const bool live = m >= 0 && m < c.n_models && !(a.done && a.done[b]);
        if (!live) {                                                    // nothing queued: the action is ignored by the environment
            if (lane < A && a.logits) a.logits[(size_t)b * A + lane] = 0.f;
            if (lane == 0) { a.actions[b] = 0; if (a.value) a.value[b] = 0.f; if (a.logp) a.logp[b] = 0.f; }
            continue;
        }
        // ---- [graph features | action mask] (observation.py:298-300) -> LN -> Linear (gnn_policy.py:96-109) ----
        for (int k = lane; k < gin; k += 32) {
            float x;
            if (k >= c.in_features_graph) x = a.mask[(size_t)b * A + (k - c.in_features_graph)] ? 1.f : 0.f;
            else if (a.graph_features) x = a.graph_features[(size_t)b * c.in_features_graph + k];
            else if (k < 9) x = a.obs_dyn[(size_t)b * 11 + k];
            else if (k < 15) x = a.graph_static[(size_t)m * 6 + (k - 9)];
            else x = a.obs_dyn[(size_t)b * 11 + (k - 6)];
            xb[k] = x;
        }
        __syncwarp();
        warp_layer_norm(xb, gin, s_gln_w, s_gln_b, lane);
        for (int k = lane; k < on; k += 32) fb[k] = a.emb[(size_t)m * on + k];
        if (lane < og) {
            float g = s_gb[lane];
            const float* row = s_gW + (size_t)lane * gin;
            for (int k = 0; k < gin; ++k) g += row[k] * xb[k];
            fb[on + lane] = g;
        }
        __syncwarp();
        // ---- read-out: hidden layer of the policy and of the value branch, each lane owns units lane + 32 i ----
        float h[POL_MAX_HPL], hv[POL_MAX_HPL];
#pragma unroll
        for (int i = 0; i < POL_MAX_HPL; ++i) {
            if (i < hpl) {
                const int j = lane + 32 * i;
                float p = s_hb[j], q = s_vhb[j];
                for (int k = 0; k < fin; ++k) { const float f = fb[k]; p += s_hWt[(size_t)k * H + j] * f; q += s_vhWt[(size_t)k * H + j] * f; }
                h[i] = act_fn(p, c.fcnet_activation); hv[i] = act_fn(q, c.fcnet_activation);
            }
        }
        float my_logit = -FLT_MAX;
        for (int o = 0; o < A; ++o) {
            float p = 0.f;
#pragma unroll
            for (int i = 0; i < POL_MAX_HPL; ++i) if (i < hpl) p += h[i] * s_lW[(size_t)o * H + lane + 32 * i];
            p = warp_sum(p) + s_lb[o];
            if (c.apply_action_mask && !a.mask[(size_t)b * A + o]) p += -FLT_MAX;   // + max(log 0, finfo.min) (gnn_policy.py:285-290)
            if (lane == o) my_logit = p;
        }
        float val = 0.f;
#pragma unroll
        for (int i = 0; i < POL_MAX_HPL; ++i) if (i < hpl) val += hv[i] * s_vW[lane + 32 * i];
        val = warp_sum(val) + s_vb[0];
        // ---- action: first maximal logit, or a categorical draw over softmax(logits) ----
        float best = my_logit; int arg = lane < A ? lane : 0x7fffffff;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
            if (ob > best || (ob == best && oa < arg)) { best = ob; arg = oa; }
        }
        const float ex = lane < A ? expf(my_logit - best) : 0.f;
        const float denom = warp_sum(ex);
        int action = arg;
        if (a.sample) {
            float cum = ex;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const float t = __shfl_up_sync(0xffffffffu, cum, o); if (lane >= o) cum += t; }
            const unsigned long long r = splitmix64(a.seed ^ ((unsigned long long)b * 0xD1342543DE82EF95ull));
            const float u = (float)(r >> 40) * (1.0f / 16777216.0f) * denom;
            const unsigned ok = __ballot_sync(0xffffffffu, lane < A && ex > 0.f && cum > u);
            action = ok ? __ffs(ok) - 1 : arg;
        }
        const float chosen = __shfl_sync(0xffffffffu, my_logit, action);
        if (lane < A && a.logits) a.logits[(size_t)b * A + lane] = my_logit;
        if (lane == 0) {
            a.actions[b] = action;
            if (a.value) a.value[b] = val;
            if (a.logp) a.logp[b] = chosen - best - logf(denom);
        }
        __syncwarp();
    }
}

// one launch per phase instead of a handful of small device-to-device copies
struct TrajArgs {
    int32_t B, A, phase;
    const float* obs; const int32_t* model; const uint8_t* mask; const int32_t* action; const float* logp; const float* value;
    const double* reward; const uint8_t* done;
    float* t_obs; int32_t* t_model; uint8_t* t_mask; int32_t* t_action; float* t_logp; float* t_value; double* t_reward; uint8_t* t_done;
};

__global__ void ramp_trajectory_record_kernel(const TrajArgs a) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= a.B) return;
    if (a.phase == 0) {
        for (int k = 0; k < 11; ++k) a.t_obs[(size_t)b * 11 + k] = a.obs[(size_t)b * 11 + k];
        for (int k = 0; k < a.A; ++k) a.t_mask[(size_t)b * a.A + k] = a.mask[(size_t)b * a.A + k];
        a.t_model[b] = a.model[b]; a.t_action[b] = a.action[b]; a.t_logp[b] = a.logp[b]; a.t_value[b] = a.value[b];
    } else {
        a.t_reward[b] = a.reward[b]; a.t_done[b] = a.done[b];
    }
}

}  // namespace ramp

using namespace ramp;

struct HostModel {
    ModelDev d{};
    std::vector<void*> allocs;
    bool set = false;
};

struct ramp_policy {
    int device = 0;
    PolicyDev P{};
    int64_t n_weights = 0;
    float* d_w = nullptr;
    std::vector<HostModel> models;
    ModelDev* d_models = nullptr;
    float* d_emb = nullptr;               // [n_models][out_node]
    float* d_gstatic = nullptr;           // [n_models][6]
    bool weights_set = false, emb_valid = false;
    int sm_count = 148;
    size_t head_smem = 0;
    // outputs of the last act / forward
    int32_t cap = 0;
    float* d_logits = nullptr; float* d_value = nullptr; float* d_logp = nullptr;
    // staging for ramp_policy_forward
    int32_t fcap = 0;
    int32_t* f_model = nullptr; float* f_gf = nullptr; uint8_t* f_mask = nullptr; int32_t* f_actions = nullptr;
    unsigned long long act_calls = 0;
    // trajectory of a rollout segment (ramp_policy_trajectory_*): [horizon][B] per field, on the device until read
    int32_t traj_h = 0, traj_b = 0, traj_a = 0;
    float* t_obs = nullptr; int32_t* t_model = nullptr; uint8_t* t_mask = nullptr; int32_t* t_action = nullptr;
    float* t_logp = nullptr; float* t_value = nullptr; double* t_reward = nullptr; uint8_t* t_done = nullptr;
};

namespace {

int64_t layout(const ramp_policy_config_t& c, PolicyDev* P) {
    int64_t o = 0;
    auto take = [&](int64_t n) { const int64_t at = o; o += n; return at; };
    const int half = c.out_features_msg / 2;
    for (int r = 0; r < c.num_rounds; ++r) {
        const int in = r == 0 ? c.in_features_node : c.out_features_hidden;
        const int out = r == c.num_rounds - 1 ? c.out_features_node : c.out_features_hidden;
        RoundW R{};
        R.in = in; R.out = out;
        R.nln_w = take(in); R.nln_b = take(in); R.nW = take((int64_t)half * in); R.nb = take(half);
        R.eln_w = take(c.in_features_edge); R.eln_b = take(c.in_features_edge); R.eW = take((int64_t)half * c.in_features_edge); R.eb = take(half);
        R.rln_w = take(c.out_features_msg); R.rln_b = take(c.out_features_msg); R.rW = take((int64_t)out * c.out_features_msg); R.rb = take(out);
        if (P) P->rounds[r] = R;
    }
    const int gin = c.in_features_graph + c.n_actions, fin = c.out_features_node + c.out_features_graph, H = c.fcnet_hidden;
    const int64_t gln_w = take(gin), gln_b = take(gin), gW = take((int64_t)c.out_features_graph * gin), gb = take(c.out_features_graph);
    const int64_t hW = take((int64_t)H * fin), hb = take(H), lW = take((int64_t)c.n_actions * H), lb = take(c.n_actions);
    const int64_t vhW = take((int64_t)H * fin), vhb = take(H), vW = take(H), vb = take(1);
    if (P) { P->gln_w = gln_w; P->gln_b = gln_b; P->gW = gW; P->gb = gb; P->hW = hW; P->hb = hb; P->lW = lW; P->lb = lb;
             P->vhW = vhW; P->vhb = vhb; P->vW = vW; P->vb = vb; }
    return o;
}

int check_config(const ramp_policy_config_t& c) {
    auto in = [](int v, int lo, int hi) { return v >= lo && v <= hi; };
    if (!in(c.in_features_node, 1, POL_MAX_DIM) || !in(c.in_features_edge, 1, POL_MAX_DIM) || !in(c.in_features_graph, 1, POL_MAX_DIM - 32))
        return perr(RAMP_ERR_BAD_ARG, "policy: feature widths must be in [1, %d]", POL_MAX_DIM);
    if (!in(c.out_features_msg, 2, POL_MAX_DIM) || (c.out_features_msg & 1) || !in(c.out_features_hidden, 1, POL_MAX_DIM) || !in(c.out_features_node, 1, POL_MAX_DIM - 32))
        return perr(RAMP_ERR_BAD_ARG, "policy: out_features_msg must be even and every width <= %d", POL_MAX_DIM);
    if (!in(c.out_features_graph, 1, 32) || !in(c.n_actions, 1, 32) || c.in_features_graph + c.n_actions > POL_MAX_DIM)
        return perr(RAMP_ERR_BAD_ARG, "policy: out_features_graph and n_actions must be <= 32");
    if (c.num_rounds < 2 || c.num_rounds > POL_MAX_ROUNDS) return perr(RAMP_ERR_BAD_ARG, "policy: num_rounds must be in [2, %d] (gnn.py:40-41)", POL_MAX_ROUNDS);
    if (c.fcnet_hidden < 32 || c.fcnet_hidden % 32 || c.fcnet_hidden > 32 * POL_MAX_HPL)
        return perr(RAMP_ERR_BAD_ARG, "policy: fcnet_hidden must be a multiple of 32 in [32, %d]", 32 * POL_MAX_HPL);
    if (!in(c.aggregator_activation, 0, 1) || !(c.fcnet_activation == 0 || c.fcnet_activation == 2))
        return perr(RAMP_ERR_BAD_ARG, "policy: unsupported activation");
    if (c.n_models < 1) return perr(RAMP_ERR_BAD_ARG, "policy: n_models must be >= 1");
    return RAMP_OK;
}

int ensure_outputs(ramp_policy* p, int32_t n) {
    if (n <= p->cap) return RAMP_OK;
    cudaFree(p->d_logits); cudaFree(p->d_value); cudaFree(p->d_logp);
    p->d_logits = p->d_value = p->d_logp = nullptr; p->cap = 0;
    PCUDA(cudaMalloc(&p->d_logits, sizeof(float) * (size_t)n * p->P.c.n_actions));
    PCUDA(cudaMalloc(&p->d_value, sizeof(float) * (size_t)n));
    PCUDA(cudaMalloc(&p->d_logp, sizeof(float) * (size_t)n));
    p->cap = n;
    return RAMP_OK;
}

int launch_embed(ramp_policy* p, cudaStream_t st) {
    for (size_t m = 0; m < p->models.size(); ++m)
        if (!p->models[m].set) return perr(RAMP_ERR_BAD_ARG, "policy: model %zu was never registered (ramp_policy_set_model)", m);
    if (!p->weights_set) return perr(RAMP_ERR_BAD_ARG, "policy: no weights (ramp_policy_set_weights)");
    std::vector<ModelDev> h(p->models.size());
    for (size_t m = 0; m < h.size(); ++m) h[m] = p->models[m].d;
    PCUDA(cudaMemcpyAsync(p->d_models, h.data(), sizeof(ModelDev) * h.size(), cudaMemcpyHostToDevice, st));
    PCUDA(cudaStreamSynchronize(st));                                  // `h` is pageable
    ramp_gnn_embed_kernel<<<(unsigned)h.size(), 256, 0, st>>>(p->P, p->d_models, p->d_emb);
    PCUDA(cudaGetLastError());
    p->emb_valid = true;
    return RAMP_OK;
}

int launch_head(ramp_policy* p, const HeadArgs& a, cudaStream_t st) {
    const int wpc = 8;
    int grid = (a.n + wpc - 1) / wpc;
    if (grid > p->sm_count * 2) grid = p->sm_count * 2;
    if (grid < 1) grid = 1;
    ramp_policy_head_kernel<<<grid, wpc * 32, p->head_smem, st>>>(p->P, a);
    PCUDA(cudaGetLastError());
    return RAMP_OK;
}

}  // namespace

extern "C" {

int64_t ramp_policy_weight_count(const ramp_policy_config_t* cfg) {
    if (!cfg || check_config(*cfg) != RAMP_OK) return -1;
    return layout(*cfg, nullptr);
}

int ramp_policy_create(int device, const ramp_policy_config_t* cfg, ramp_policy_t** out) {
    if (!cfg || !out) return perr(RAMP_ERR_BAD_ARG, "null argument");
    int rc = check_config(*cfg);
    if (rc != RAMP_OK) return rc;
    PCUDA(cudaSetDevice(device));
    ramp_policy* p = new ramp_policy();
    p->device = device;
    p->P.c = *cfg;
    p->n_weights = layout(*cfg, &p->P);
    p->models.resize(cfg->n_models);
    cudaDeviceProp prop{};
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) p->sm_count = prop.multiProcessorCount;
    p->head_smem = head_smem_floats(*cfg, 8) * sizeof(float);
    auto fail = [&](int code) { ramp_policy_destroy(p); return code; };
    if (p->head_smem > 200 * 1024) return fail(perr(RAMP_ERR_CAPACITY, "policy: the read-out needs %zu B of shared memory (max 200 KiB)", p->head_smem));
    if (cudaFuncSetAttribute(ramp_policy_head_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->head_smem) != cudaSuccess)
        return fail(perr(RAMP_ERR_CUDA, "policy: cannot reserve %zu B of shared memory", p->head_smem));
    if (cudaMalloc(&p->d_w, sizeof(float) * p->n_weights) != cudaSuccess ||
        cudaMalloc(&p->d_models, sizeof(ModelDev) * cfg->n_models) != cudaSuccess ||
        cudaMalloc(&p->d_emb, sizeof(float) * (size_t)cfg->n_models * cfg->out_features_node) != cudaSuccess ||
        cudaMalloc(&p->d_gstatic, sizeof(float) * (size_t)cfg->n_models * 6) != cudaSuccess)
        return fail(perr(RAMP_ERR_CUDA, "policy: out of device memory"));
    cudaMemset(p->d_emb, 0, sizeof(float) * (size_t)cfg->n_models * cfg->out_features_node);
    cudaMemset(p->d_gstatic, 0, sizeof(float) * (size_t)cfg->n_models * 6);
    p->P.w = p->d_w;
    *out = p;
    return RAMP_OK;
}

void ramp_policy_destroy(ramp_policy_t* p) {
    if (!p) return;
    cudaSetDevice(p->device);
    for (auto& m : p->models) for (void* a : m.allocs) cudaFree(a);
    cudaFree(p->d_w); cudaFree(p->d_models); cudaFree(p->d_emb); cudaFree(p->d_gstatic);
    cudaFree(p->d_logits); cudaFree(p->d_value); cudaFree(p->d_logp);
    cudaFree(p->f_model); cudaFree(p->f_gf); cudaFree(p->f_mask); cudaFree(p->f_actions);
    cudaFree(p->t_obs); cudaFree(p->t_model); cudaFree(p->t_mask); cudaFree(p->t_action); cudaFree(p->t_logp); cudaFree(p->t_value);
    cudaFree(p->t_reward); cudaFree(p->t_done);
    delete p;
}

int ramp_policy_set_weights(ramp_policy_t* p, const float* weights, int64_t n) {
    if (!p || !weights) return perr(RAMP_ERR_BAD_ARG, "null argument");
    if (n != p->n_weights) return perr(RAMP_ERR_BAD_ARG, "policy: %lld weights given, the configuration has %lld", (long long)n, (long long)p->n_weights);
    PCUDA(cudaSetDevice(p->device));
    PCUDA(cudaDeviceSynchronize());                                    // a running rollout may still read the old set
    PCUDA(cudaMemcpy(p->d_w, weights, sizeof(float) * n, cudaMemcpyHostToDevice));
    p->weights_set = true;
    p->emb_valid = false;
    return RAMP_OK;
}

int ramp_policy_set_model(ramp_policy_t* p, int32_t model, int32_t n_nodes, int32_t n_edges, const float* node_features,
                          const float* edge_features, const int32_t* edges_src, const int32_t* edges_dst, const float* graph_static) {
    if (!p || !node_features || !graph_static || (n_edges > 0 && (!edge_features || !edges_src || !edges_dst)))
        return perr(RAMP_ERR_BAD_ARG, "null argument");
    const ramp_policy_config_t& c = p->P.c;
    if (model < 0 || model >= c.n_models || n_nodes < 1 || n_edges < 0) return perr(RAMP_ERR_BAD_ARG, "policy: bad model %d (%d nodes, %d edges)", model, n_nodes, n_edges);
    for (int e = 0; e < n_edges; ++e)
        if (edges_src[e] < 0 || edges_src[e] >= n_nodes || edges_dst[e] < 0 || edges_dst[e] >= n_nodes)
            return perr(RAMP_ERR_BAD_ARG, "policy: edge %d of model %d names node %d -> %d of %d", e, model, edges_src[e], edges_dst[e], n_nodes);
    PCUDA(cudaSetDevice(p->device));
    HostModel& hm = p->models[model];
    for (void* a : hm.allocs) cudaFree(a);
    hm.allocs.clear();
    hm.set = false;
    // incoming-edge lists by destination, in edge order (the order DGL delivers a node's mailbox is not observable through a mean)
    std::vector<int32_t> ptr(n_nodes + 1, 0), ine(n_edges), ins(n_edges);
    for (int e = 0; e < n_edges; ++e) ptr[edges_dst[e] + 1]++;
    for (int v = 0; v < n_nodes; ++v) ptr[v + 1] += ptr[v];
    std::vector<int32_t> cur(ptr.begin(), ptr.end() - 1);
    for (int e = 0; e < n_edges; ++e) { const int at = cur[edges_dst[e]]++; ine[at] = e; ins[at] = edges_src[e]; }
    const int half = c.out_features_msg / 2;
    auto up = [&](const void* src, size_t bytes, const void** dst) -> int {
        void* d = nullptr;
        PCUDA(cudaMalloc(&d, bytes ? bytes : 4));
        hm.allocs.push_back(d);
        if (src && bytes) PCUDA(cudaMemcpy(d, src, bytes, cudaMemcpyHostToDevice));
        *dst = d;
        return RAMP_OK;
    };
    int rc;
    ModelDev d{};
    d.n_nodes = n_nodes; d.n_edges = n_edges;
    if ((rc = up(node_features, sizeof(float) * (size_t)n_nodes * c.in_features_node, (const void**)&d.nf))) return rc;
    if ((rc = up(edge_features, sizeof(float) * (size_t)n_edges * c.in_features_edge, (const void**)&d.ef))) return rc;
    if ((rc = up(ptr.data(), sizeof(int32_t) * ptr.size(), (const void**)&d.in_ptr))) return rc;
    if ((rc = up(ine.data(), sizeof(int32_t) * ine.size(), (const void**)&d.in_edge))) return rc;
    if ((rc = up(ins.data(), sizeof(int32_t) * ins.size(), (const void**)&d.in_src))) return rc;
    if ((rc = up(nullptr, sizeof(float) * (size_t)n_nodes * POL_MAX_DIM, (const void**)&d.z0))) return rc;
    if ((rc = up(nullptr, sizeof(float) * (size_t)n_nodes * POL_MAX_DIM, (const void**)&d.z1))) return rc;
    if ((rc = up(nullptr, sizeof(float) * (size_t)n_nodes * half, (const void**)&d.hn))) return rc;
    if ((rc = up(nullptr, sizeof(float) * (size_t)n_edges * half, (const void**)&d.he))) return rc;
    PCUDA(cudaMemcpy(p->d_gstatic + (size_t)model * 6, graph_static, sizeof(float) * 6, cudaMemcpyHostToDevice));
    hm.d = d;
    hm.set = true;
    p->emb_valid = false;
    return RAMP_OK;
}

int ramp_policy_embed(ramp_policy_t* p, float* embeddings_out) {
    if (!p) return perr(RAMP_ERR_BAD_ARG, "null policy");
    PCUDA(cudaSetDevice(p->device));
    int rc = launch_embed(p, 0);
    if (rc != RAMP_OK) return rc;
    PCUDA(cudaStreamSynchronize(0));
    if (embeddings_out)
        PCUDA(cudaMemcpy(embeddings_out, p->d_emb, sizeof(float) * (size_t)p->P.c.n_models * p->P.c.out_features_node, cudaMemcpyDeviceToHost));
    return RAMP_OK;
}

int ramp_policy_forward(ramp_policy_t* p, int32_t n, const int32_t* model, const float* graph_features, const uint8_t* action_mask,
                        float* logits_out, float* value_out) {
    if (!p || !model || !graph_features || !action_mask) return perr(RAMP_ERR_BAD_ARG, "null argument");
    if (n < 1) return RAMP_OK;
    const ramp_policy_config_t& c = p->P.c;
    PCUDA(cudaSetDevice(p->device));
    int rc;
    if (!p->emb_valid && (rc = launch_embed(p, 0)) != RAMP_OK) return rc;
    if ((rc = ensure_outputs(p, n)) != RAMP_OK) return rc;
    if (n > p->fcap) {
        cudaFree(p->f_model); cudaFree(p->f_gf); cudaFree(p->f_mask); cudaFree(p->f_actions);
        p->f_model = nullptr; p->f_gf = nullptr; p->f_mask = nullptr; p->f_actions = nullptr; p->fcap = 0;
        PCUDA(cudaMalloc(&p->f_model, sizeof(int32_t) * (size_t)n));
        PCUDA(cudaMalloc(&p->f_gf, sizeof(float) * (size_t)n * c.in_features_graph));
        PCUDA(cudaMalloc(&p->f_mask, (size_t)n * c.n_actions));
        PCUDA(cudaMalloc(&p->f_actions, sizeof(int32_t) * (size_t)n));
        p->fcap = n;
    }
    PCUDA(cudaMemcpy(p->f_model, model, sizeof(int32_t) * (size_t)n, cudaMemcpyHostToDevice));
    PCUDA(cudaMemcpy(p->f_gf, graph_features, sizeof(float) * (size_t)n * c.in_features_graph, cudaMemcpyHostToDevice));
    PCUDA(cudaMemcpy(p->f_mask, action_mask, (size_t)n * c.n_actions, cudaMemcpyHostToDevice));
    HeadArgs a{};
    a.n = n; a.graph_features = p->f_gf; a.model = p->f_model; a.mask = p->f_mask; a.emb = p->d_emb; a.graph_static = p->d_gstatic;
    a.logits = p->d_logits; a.value = p->d_value; a.logp = p->d_logp; a.actions = p->f_actions;
    if ((rc = launch_head(p, a, 0)) != RAMP_OK) return rc;
    PCUDA(cudaStreamSynchronize(0));
    if (logits_out) PCUDA(cudaMemcpy(logits_out, p->d_logits, sizeof(float) * (size_t)n * c.n_actions, cudaMemcpyDeviceToHost));
    if (value_out) PCUDA(cudaMemcpy(value_out, p->d_value, sizeof(float) * (size_t)n, cudaMemcpyDeviceToHost));
    return RAMP_OK;
}

int ramp_policy_act(ramp_policy_t* p, ramp_engine_t* eng, int32_t sample, uint64_t seed) {
    if (!p || !eng) return perr(RAMP_ERR_BAD_ARG, "null argument");
    ramp_env_buffers_t eb{};
    int rc = ramp_env_buffers(eng, &eb);
    if (rc != RAMP_OK) return rc;
    const ramp_policy_config_t& c = p->P.c;
    if (ramp_internal_device(eng) != p->device) return perr(RAMP_ERR_BAD_ARG, "policy and engine live on different devices");
    if (eb.n_actions != c.n_actions) return perr(RAMP_ERR_BAD_ARG, "policy has %d actions, the environment %d", c.n_actions, eb.n_actions);
    if (c.in_features_graph != 17) return perr(RAMP_ERR_BAD_ARG, "the environment emits 17 graph features, the policy expects %d", c.in_features_graph);
    if (eb.n_models > c.n_models) return perr(RAMP_ERR_BAD_ARG, "the environment has %d job types, the policy %d", eb.n_models, c.n_models);
    PCUDA(cudaSetDevice(p->device));
    cudaStream_t st = ramp_internal_stream(eng);
    int launches = 1;
    if (!p->emb_valid) { if ((rc = launch_embed(p, st)) != RAMP_OK) return rc; ++launches; }
    if ((rc = ensure_outputs(p, eb.n_episodes)) != RAMP_OK) return rc;
    HeadArgs a{};
    a.n = eb.n_episodes; a.obs_dyn = eb.obs_dynamic; a.graph_static = p->d_gstatic; a.model = eb.queued_model; a.done = eb.done;
    a.mask = eb.action_mask; a.emb = p->d_emb; a.logits = p->d_logits; a.value = p->d_value; a.logp = p->d_logp; a.actions = eb.actions;
    a.sample = sample; a.seed = seed ^ (0x9E3779B97F4A7C15ull * (++p->act_calls));
    if ((rc = launch_head(p, a, st)) != RAMP_OK) return rc;
    ramp_internal_count_launches(eng, launches);
    return RAMP_OK;
}

int ramp_policy_read(ramp_policy_t* p, ramp_engine_t* eng, float* logits_out, float* value_out, float* logp_out, int32_t* actions_out) {
    if (!p || !eng) return perr(RAMP_ERR_BAD_ARG, "null argument");
    ramp_env_buffers_t eb{};
    int rc = ramp_env_buffers(eng, &eb);
    if (rc != RAMP_OK) return rc;
    if (eb.n_episodes > p->cap) return perr(RAMP_ERR_BAD_ARG, "policy: nothing to read (ramp_policy_act was not called)");
    cudaStream_t st = ramp_internal_stream(eng);
    const size_t B = (size_t)eb.n_episodes;
    if (logits_out) PCUDA(cudaMemcpyAsync(logits_out, p->d_logits, sizeof(float) * B * p->P.c.n_actions, cudaMemcpyDeviceToHost, st));
    if (value_out) PCUDA(cudaMemcpyAsync(value_out, p->d_value, sizeof(float) * B, cudaMemcpyDeviceToHost, st));
    if (logp_out) PCUDA(cudaMemcpyAsync(logp_out, p->d_logp, sizeof(float) * B, cudaMemcpyDeviceToHost, st));
    if (actions_out) PCUDA(cudaMemcpyAsync(actions_out, eb.actions, sizeof(int32_t) * B, cudaMemcpyDeviceToHost, st));
    PCUDA(cudaStreamSynchronize(st));
    return RAMP_OK;
}

void* ramp_pinned_alloc(size_t bytes) {
    void* ptr = nullptr;
    if (cudaMallocHost(&ptr, bytes ? bytes : 1) != cudaSuccess) { perr(RAMP_ERR_CUDA, "cudaMallocHost(%zu) failed", bytes); return nullptr; }
    return ptr;
}

void ramp_pinned_free(void* ptr) { if (ptr) cudaFreeHost(ptr); }

int ramp_policy_trajectory_begin(ramp_policy_t* p, ramp_engine_t* eng, int32_t horizon) {
    if (!p || !eng || horizon < 1) return perr(RAMP_ERR_BAD_ARG, "policy: bad trajectory horizon");
    ramp_env_buffers_t eb{};
    int rc = ramp_env_buffers(eng, &eb);
    if (rc != RAMP_OK) return rc;
    PCUDA(cudaSetDevice(p->device));
    if (horizon == p->traj_h && eb.n_episodes == p->traj_b && eb.n_actions == p->traj_a) return RAMP_OK;
    PCUDA(cudaStreamSynchronize(ramp_internal_stream(eng)));
    cudaFree(p->t_obs); cudaFree(p->t_model); cudaFree(p->t_mask); cudaFree(p->t_action); cudaFree(p->t_logp); cudaFree(p->t_value);
    cudaFree(p->t_reward); cudaFree(p->t_done);
    p->t_obs = nullptr; p->t_model = nullptr; p->t_mask = nullptr; p->t_action = nullptr; p->t_logp = nullptr; p->t_value = nullptr;
    p->t_reward = nullptr; p->t_done = nullptr; p->traj_h = 0;
    const size_t n = (size_t)horizon * (size_t)eb.n_episodes;
    PCUDA(cudaMalloc(&p->t_obs, sizeof(float) * 11 * n));
    PCUDA(cudaMalloc(&p->t_model, sizeof(int32_t) * n));
    PCUDA(cudaMalloc(&p->t_mask, (size_t)eb.n_actions * n));
    PCUDA(cudaMalloc(&p->t_action, sizeof(int32_t) * n));
    PCUDA(cudaMalloc(&p->t_logp, sizeof(float) * n));
    PCUDA(cudaMalloc(&p->t_value, sizeof(float) * n));
    PCUDA(cudaMalloc(&p->t_reward, sizeof(double) * n));
    PCUDA(cudaMalloc(&p->t_done, n));
    p->traj_h = horizon; p->traj_b = eb.n_episodes; p->traj_a = eb.n_actions;
    return RAMP_OK;
}

int ramp_policy_trajectory_record(ramp_policy_t* p, ramp_engine_t* eng, int32_t t, int32_t phase) {
    if (!p || !eng) return perr(RAMP_ERR_BAD_ARG, "null argument");
    if (p->traj_h < 1 || t < 0 || t >= p->traj_h) return perr(RAMP_ERR_BAD_ARG, "policy: trajectory slot %d outside [0, %d)", t, p->traj_h);
    ramp_env_buffers_t eb{};
    int rc = ramp_env_buffers(eng, &eb);
    if (rc != RAMP_OK) return rc;
    if (eb.n_episodes != p->traj_b || eb.n_actions != p->traj_a || eb.n_episodes > p->cap)
        return perr(RAMP_ERR_BAD_ARG, "policy: the trajectory was set up for another environment, or ramp_policy_act was not called");
    cudaStream_t st = ramp_internal_stream(eng);
    const size_t B = (size_t)eb.n_episodes, o = (size_t)t * B;
    // phase 0: after ramp_policy_act, before the environment steps -- what the policy saw and decided;
    // phase 1: after the environment stepped -- what came back
    TrajArgs a{};
    a.B = eb.n_episodes; a.A = p->traj_a; a.phase = phase;
    a.obs = eb.obs_dynamic; a.model = eb.queued_model; a.mask = eb.action_mask; a.action = eb.actions; a.logp = p->d_logp; a.value = p->d_value;
    a.reward = eb.reward; a.done = eb.done;
    a.t_obs = p->t_obs + o * 11; a.t_model = p->t_model + o; a.t_mask = p->t_mask + o * p->traj_a; a.t_action = p->t_action + o;
    a.t_logp = p->t_logp + o; a.t_value = p->t_value + o; a.t_reward = p->t_reward + o; a.t_done = p->t_done + o;
    ramp_trajectory_record_kernel<<<(unsigned)((B + 127) / 128), 128, 0, st>>>(a);
    PCUDA(cudaGetLastError());
    ramp_internal_count_launches(eng, 1);
    return RAMP_OK;
}

int ramp_policy_trajectory_read(ramp_policy_t* p, ramp_engine_t* eng, int32_t n_steps, float* obs_dynamic_out, int32_t* model_out,
                                uint8_t* action_mask_out, int32_t* action_out, float* logp_out, float* value_out, double* reward_out,
                                uint8_t* done_out) {
    if (!p || !eng) return perr(RAMP_ERR_BAD_ARG, "null argument");
    if (n_steps < 1 || n_steps > p->traj_h) return perr(RAMP_ERR_BAD_ARG, "policy: %d steps asked of a trajectory of %d", n_steps, p->traj_h);
    cudaStream_t st = ramp_internal_stream(eng);
    const size_t n = (size_t)n_steps * (size_t)p->traj_b;
    if (obs_dynamic_out) PCUDA(cudaMemcpyAsync(obs_dynamic_out, p->t_obs, sizeof(float) * 11 * n, cudaMemcpyDeviceToHost, st));
    if (model_out) PCUDA(cudaMemcpyAsync(model_out, p->t_model, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, st));
    if (action_mask_out) PCUDA(cudaMemcpyAsync(action_mask_out, p->t_mask, (size_t)p->traj_a * n, cudaMemcpyDeviceToHost, st));
    if (action_out) PCUDA(cudaMemcpyAsync(action_out, p->t_action, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, st));
    if (logp_out) PCUDA(cudaMemcpyAsync(logp_out, p->t_logp, sizeof(float) * n, cudaMemcpyDeviceToHost, st));
    if (value_out) PCUDA(cudaMemcpyAsync(value_out, p->t_value, sizeof(float) * n, cudaMemcpyDeviceToHost, st));
    if (reward_out) PCUDA(cudaMemcpyAsync(reward_out, p->t_reward, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
    if (done_out) PCUDA(cudaMemcpyAsync(done_out, p->t_done, n, cudaMemcpyDeviceToHost, st));
    PCUDA(cudaStreamSynchronize(st));
    return RAMP_OK;
}

}  // extern "C"
