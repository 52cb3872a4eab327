// ck_net.cuh -- internal interface of the policy/value network (K3).
#pragma once
#include "ck_common.cuh"

namespace ck {

// offsets (in floats) into the Keras-ordered weight blob described in include/ckb200.h
struct ConvParams { int64_t kernel, bias, gamma, beta, mean, var; int cin, cout; };
struct NetLayout {
    ConvParams conv[8];      // conv0..conv6 + policy conv3x3  (training_pipeline.py:57-88)
    ConvParams pol1x1;       // 128 -> 8                        (:89-96)
    int64_t pol_dense_k, pol_dense_b;                        // 512 -> 512 softmax (:97-100)
    ConvParams val1x1;       // 128 -> 1                        (:102-105)
    int64_t val_d1_k, val_d1_b, val_d1_gamma, val_d1_beta, val_d1_mean, val_d1_var;   // 64 -> 64 (:107-110)
    int64_t val_d2_k, val_d2_b;                              // 64 -> 1 tanh (:111-112)
    int64_t total;
};
NetLayout net_layout();

constexpr float kBnEps = 1e-3f;      // Keras BatchNormalization default epsilon
constexpr int kTowerConvs = 8;       // 3x3 convolutions evaluated by the tower kernels
constexpr int kC = 128;              // trunk width

}  // namespace ck

struct ck_net {
    int device = 0;
    int impl = CK_NET_IMPL_TC;
    bool have_weights = false;
    uint64_t weights_gen = 0;        // bumped by every ck_net_set_weights*: consumers that cache evaluations compare it
    float *d_blob = nullptr;         // Keras-ordered fp32 parameters
    float *d_stage = nullptr;        // ck_net_set_weights: the upload is staged here and compared with d_blob first
    float *d_scale = nullptr;        // folded BN: per layer 128 scale + 128 shift (tower), heads after
    // tcgen05 tower operands (built by ck_net_tc.cu)
    void *d_wpack = nullptr;         // split-fp16 weights in UMMA core-matrix order
    size_t wpack_bytes = 0;
    void *d_hpack = nullptr;         // policy Dense(512) weights, split fp16, UMMA layout (ck_heads_tc.cu)
    void *d_wts = nullptr;           // split-fp16 weights in k-step order for the weights-in-TMEM tower (ck_net_ts.cu)
    // activation scratch, grown on demand
    // tower -> heads scratch (net_reserve); two sets ("lanes") so that two batches of one network can be in flight at once
    size_t act0_floats = 0, act1_floats = 0;      // lane 0 (also what ck_net_last_features reads)
    float *d_act0 = nullptr, *d_act1 = nullptr;
    size_t act0b_floats = 0, act1b_floats = 0;    // lane 1 (the second slot group of an overlapped engine)
    float *d_act0b = nullptr, *d_act1b = nullptr;
    ck_leaf *d_leaves = nullptr;     // staging for host entry points
    float *d_policy = nullptr, *d_value = nullptr;
    int64_t io_cap = 0;
    cudaEvent_t ev_after_tower = nullptr;   // profiling hook: recorded between tower and heads
    int32_t *d_range_flag = nullptr; // set by the tensor-core kernels when an activation leaves the split-fp16 range
    float *d_value_pre = nullptr;    // ck_net_forward_logits: receives the value head's pre-tanh output (else nullptr)
    float *d_logits_out = nullptr;   // ck_net_forward_logits on the CUDA-core heads: receives the policy logits
};

namespace ck {
// folded BN table layout in ck_net::d_scale (floats)
constexpr int kScaleTower = 0;                        // 8 layers x (128 scale, 128 shift)
constexpr int kScalePol1x1 = kTowerConvs * 2 * kC;    // 8 scale, 8 shift
constexpr int kScaleVal1x1 = kScalePol1x1 + 16;       // 1 scale, 1 shift
constexpr int kScaleValD1 = kScaleVal1x1 + 2;         // 64 scale, 64 shift
constexpr int kScaleTotal = kScaleValD1 + 128;

int net_reserve(ck_net *net, int64_t n, bool full_maps, int lane = 0);
// CK_ERR_NET_RANGE (and the flag cleared) if a tensor-core kernel of this net saw an activation beyond the split-fp16
// range since the last check; synchronises the default stream's view of the flag with a blocking 4-byte copy
int net_check_range(ck_net *net);
// n_dev (optional): device pointer to the live row count; rows >= *n_dev are skipped
int net_forward_rows(ck_net *net, const ck_leaf *d_leaves, int64_t max_n, const int32_t *n_dev,
                     float *d_policy, float *d_value, cudaStream_t stream, int *launches, int lane = 0);
// tcgen05 tower (ck_net_tc.cu): leaves -> trunk (conv6 output) and policy-conv output,
// both fp32 [n][128][64]
int net_tc_prepare(ck_net *net);
int net_tc_tower(ck_net *net, const ck_leaf *d_leaves, int64_t max_n, const int32_t *n_dev,
                 float *d_trunk, float *d_pconv, cudaStream_t stream, int *launches);
// weights-in-TMEM tower (ck_net_ts.cu), same contract; the default tensor-core path
int net_ts_prepare(ck_net *net);
int net_ts_tower(ck_net *net, const ck_leaf *d_leaves, int64_t max_n, const int32_t *n_dev,
                 float *d_trunk, float *d_pconv, cudaStream_t stream, int *launches);
// tensor-core policy Dense + fp32 tails of both heads on the fused tower outputs (ck_heads_tc.cu)
int net_heads_tc_prepare(ck_net *net);
int net_heads_tc(ck_net *net, const float *d_pflat, const float *d_vconv, float *d_logits, int64_t max_n, const int32_t *n_dev,
                 float *d_policy, float *d_value, cudaStream_t stream, int *launches);
}  // namespace ck
