// ck_net.cu -- K3 policy/value network: weight handling, fp32 CUDA-core tower (cross-check
// implementation), heads kernel (shared with the tcgen05 tower), C ABI.
//
// Architecture = create_nn (reference training_pipeline.py:44-120): 7 x [Conv3x3(128) -> ReLU
// -> BN], policy head Conv3x3(128)->ReLU->BN->Conv1x1(8)->ReLU->BN->Flatten(x,y,c)->
// Dense(512, softmax); value head Conv1x1(1)->ReLU->BN->Flatten->Dense(64, ReLU)->BN->
// Dense(1, tanh).  Input planes per Checkers.py:37-48.
#include <math.h>
#include <vector>
#include "ck_net.cuh"

namespace ck {

NetLayout net_layout() {
    NetLayout L;
    int64_t o = 0;
    auto conv = [&](ConvParams &c, int k, int cin, int cout, bool bn) {
        c.cin = cin; c.cout = cout;
        c.kernel = o; o += (int64_t)k * k * cin * cout;
        c.bias = o; o += cout;
        if (bn) { c.gamma = o; o += cout; c.beta = o; o += cout; c.mean = o; o += cout; c.var = o; o += cout; }
    };
    conv(L.conv[0], 3, 14, kC, true);
    for (int i = 1; i < 8; ++i) conv(L.conv[i], 3, kC, kC, true);
    conv(L.pol1x1, 1, kC, 8, true);
    L.pol_dense_k = o; o += 512 * 512; L.pol_dense_b = o; o += 512;
    conv(L.val1x1, 1, kC, 1, true);
    L.val_d1_k = o; o += 64 * 64; L.val_d1_b = o; o += 64;
    L.val_d1_gamma = o; o += 64; L.val_d1_beta = o; o += 64; L.val_d1_mean = o; o += 64; L.val_d1_var = o; o += 64;
    L.val_d2_k = o; o += 64; L.val_d2_b = o; o += 1;
    L.total = o;
    return L;
}

__constant__ float c_plane5[81];     // float32(n / 80.0) as numpy produces it (Checkers.py:346-361 + :432)
static bool g_plane5_ready[64] = {false};

static int ensure_constants(int device) {
    if (device < 64 && g_plane5_ready[device]) return CK_OK;
    float t[81];
    for (int i = 0; i <= 80; ++i) t[i] = (float)((double)i / 80.0);
    CK_CUDA(cudaMemcpyToSymbol(c_plane5, t, sizeof(t)));
    if (device < 64) g_plane5_ready[device] = true;
    return CK_OK;
}

// ---- folded batch-norm table ------------------------------------------------------------
struct FoldJob { int64_t gamma, beta, mean, var; int n; int dst; };
__global__ void fold_bn_kernel(const float *__restrict__ blob, float *__restrict__ out, FoldJob j) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= j.n) return;
    const float sc = __fdiv_rn(blob[j.gamma + i], __fsqrt_rn(__fadd_rn(blob[j.var + i], kBnEps)));
    out[j.dst + i] = sc;
    out[j.dst + j.n + i] = __fsub_rn(blob[j.beta + i], __fmul_rn(blob[j.mean + i], sc));
}

// ---- network input ------------------------------------------------------------------------
__device__ __forceinline__ float leaf_plane(const ck_leaf &L, int ci, int x, int y) {
    if (ci == 4) return (float)(L.info & 1u);
    if (ci == 5) return c_plane5[(L.info >> 8) & 0xFFu];
    if (((x ^ y) & 1) == 0) return 0.f;
    const uint32_t bit = 1u << (4 * x + (y >> 1));
    uint32_t set;
    switch (ci) {
        case 0: set = L.p1 & ~L.k; break;
        case 1: set = L.p1 & L.k; break;
        case 2: set = L.p2 & ~L.k; break;
        case 3: set = L.p2 & L.k; break;
        default: set = L.mask[ci - 6]; break;
    }
    return (set & bit) ? 1.f : 0.f;
}

__global__ void planes_to_leaf_kernel(const float *__restrict__ x, int64_t n, ck_leaf *__restrict__ out) {
    // inverse of leaf_plane for the Keras-signature entry point: x is [n,8,8,14] channels-last
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float *p = x + i * 896;
    uint32_t b[14];
    for (int c = 0; c < 14; ++c) b[c] = 0;
    for (int s = 0; s < 32; ++s) {
        const int xx = s >> 2, yy = ((s & 3) << 1) | ((xx & 1) ^ 1);
        for (int c = 0; c < 14; ++c)
            if (c != 4 && c != 5 && p[(xx * 8 + yy) * 14 + c] != 0.f) b[c] |= 1u << s;
    }
    ck_leaf L;
    L.p1 = b[0] | b[1]; L.p2 = b[2] | b[3]; L.k = b[1] | b[3];
    const int player = p[4] != 0.f;
    const int p5 = (int)lrintf(p[5] * 80.f);
    L.info = (uint32_t)player | ((uint32_t)p5 << 8);
    for (int c = 0; c < 8; ++c) L.mask[c] = b[6 + c];
    out[i] = L;
}

// ---- fp32 CUDA-core tower (cross-check implementation) -------------------------------------
// One CTA per position; thread = (output channel, half board); the padded input planes sit
// in shared memory and are read as warp-wide broadcasts; weights stream from L2 in the
// Keras [kh,kw,Cin,Cout] layout, coalesced over Cout.
constexpr int kRowStride = 12;                 // padded row of 10 floats, 16-byte aligned
constexpr int kPlaneStride = 10 * kRowStride;  // 10 padded rows

template <int CIN, bool FROM_LEAF>
__global__ void __launch_bounds__(256)
conv3x3_simt_kernel(const ck_leaf *__restrict__ leaves, const float *__restrict__ in, const int32_t *__restrict__ n_dev,
                    const float *__restrict__ w, const float *__restrict__ bias, const float *__restrict__ scale,
                    float *__restrict__ out) {
    const int pos = blockIdx.x;
    if (n_dev != nullptr && pos >= *n_dev) return;
    extern __shared__ __align__(16) float s_in[];
    const int tid = threadIdx.x;
    for (int i = tid; i < CIN * kPlaneStride; i += 256) s_in[i] = 0.f;
    __syncthreads();
    if (FROM_LEAF) {
        const ck_leaf L = leaves[pos];
        for (int i = tid; i < CIN * 64; i += 256) {
            const int ci = i >> 6, x = (i >> 3) & 7, y = i & 7;
            s_in[ci * kPlaneStride + (x + 1) * kRowStride + (y + 1)] = leaf_plane(L, ci, x, y);
        }
    } else {
        const float *src = in + (int64_t)pos * CIN * 64;
        for (int i = tid; i < CIN * 64; i += 256) {
            const int ci = i >> 6, x = (i >> 3) & 7, y = i & 7;
            s_in[ci * kPlaneStride + (x + 1) * kRowStride + (y + 1)] = src[i];
        }
    }
    __syncthreads();
    const int co = tid & 127, half = tid >> 7;
    float acc[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) acc[i] = 0.f;
    for (int ci = 0; ci < CIN; ++ci) {
        float wk[9];
#pragma unroll
        for (int t = 0; t < 9; ++t) wk[t] = __ldg(w + ((int64_t)t * CIN + ci) * kC + co);
        const float *ip = s_in + ci * kPlaneStride + (4 * half) * kRowStride;
#pragma unroll
        for (int r6 = 0; r6 < 6; ++r6) {
            float v[12];
            const float4 a = *reinterpret_cast<const float4 *>(ip + r6 * kRowStride);
            const float4 b = *reinterpret_cast<const float4 *>(ip + r6 * kRowStride + 4);
            const float4 c = *reinterpret_cast<const float4 *>(ip + r6 * kRowStride + 8);
            v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
            v[8] = c.x; v[9] = c.y; v[10] = c.z; v[11] = c.w;
#pragma unroll
            for (int kh = 0; kh < 3; ++kh) {
                const int r = r6 - kh;
                if (r < 0 || r > 3) continue;
#pragma unroll
                for (int cx = 0; cx < 8; ++cx)
#pragma unroll
                    for (int kw = 0; kw < 3; ++kw) acc[r * 8 + cx] = fmaf(v[cx + kw], wk[kh * 3 + kw], acc[r * 8 + cx]);
            }
        }
    }
    const float b = bias[co], sc = scale[co], sh = scale[kC + co];
    float *dst = out + ((int64_t)pos * kC + co) * 64 + half * 32;
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
        float4 o;
        o.x = fmaf(fmaxf(acc[i] + b, 0.f), sc, sh);
        o.y = fmaf(fmaxf(acc[i + 1] + b, 0.f), sc, sh);
        o.z = fmaf(fmaxf(acc[i + 2] + b, 0.f), sc, sh);
        o.w = fmaf(fmaxf(acc[i + 3] + b, 0.f), sc, sh);
        *reinterpret_cast<float4 *>(dst + i) = o;
    }
}

// ---- heads --------------------------------------------------------------------------------
// Policy head tail (conv1x1 128->8, ReLU, BN, Flatten over (x,y,c), Dense 512->512, softmax;
// training_pipeline.py:89-100) and value head (conv1x1 128->1, ReLU, BN, Dense 64 ReLU BN, Dense 1
// tanh; :102-112) in fp32 on the CUDA cores.  kHeadPB = 32 positions per CTA: the 1 MB dense matrix
// is the dominant traffic (L2 -> SM), read once per 32 positions (134 MB per 4096-position batch;
// at 8 positions per CTA it was 537 MB and the kernel was L2-bandwidth bound).
// Inputs: trunk = conv6 output, pconv = policy conv3x3 output, fp32 [n][128][64].
constexpr int kHeadPB = 32;
constexpr int kHeadStride = kHeadPB + 4;            // row stride of s_flat[i][p]: float4-aligned, spreads banks
constexpr int kHeadThreads = 512;
constexpr int kHeadSmem = (512 * kHeadStride + 128 * 8 + kHeadPB * 64) * (int)sizeof(float);

struct HeadParams {
    int64_t pol1x1_k, pol1x1_b, pol_dense_k, pol_dense_b;
    int64_t val1x1_k, val1x1_b, val_d1_k, val_d1_b, val_d2_k, val_d2_b;
};

// kFused: the tower already applied both conv1x1 (ck_net_ts.cu); `trunk` then holds the value conv output
// [n][64] and `pconv` the flattened policy features [n][512].
template <bool kFused>
__global__ void __launch_bounds__(kHeadThreads, 1)
heads_kernel(const float *__restrict__ trunk, const float *__restrict__ pconv, int64_t max_n,
             const int32_t *__restrict__ n_dev, const float *__restrict__ blob, const float *__restrict__ fold,
             HeadParams hp, float *__restrict__ policy, float *__restrict__ value,
             float *__restrict__ logits_out, float *__restrict__ vpre_out) {
    int64_t n = max_n;
    if (n_dev != nullptr) n = min((int64_t)*n_dev, max_n);
    const int64_t base = (int64_t)blockIdx.x * kHeadPB;
    if (base >= n) return;
    const int npos = (int)min((int64_t)kHeadPB, n - base);
    extern __shared__ __align__(16) float s_head[];
    float *s_flat = s_head;                         // [i][p] during the dense, [p][512] for the softmax
    float *s_wp = s_head + 512 * kHeadStride;       // policy conv1x1 kernel [c][o]
    float *s_v1 = s_wp + 128 * 8;                   // value conv1x1 output [p][64]
    const int tid = threadIdx.x, lane = tid & 31, wp = tid >> 5;
    if (kFused) {
        for (int idx = tid; idx < kHeadPB * 512; idx += kHeadThreads) {
            const int p = idx >> 9, i = idx & 511;
            s_flat[i * kHeadStride + p] = p < npos ? pconv[(base + p) * 512 + i] : 0.f;
        }
        for (int idx = tid; idx < kHeadPB * 64; idx += kHeadThreads) s_v1[idx] = (idx >> 6) < npos ? trunk[base * 64 + idx] : 0.f;
    } else {
        for (int i = tid; i < 128 * 8; i += kHeadThreads) s_wp[i] = blob[hp.pol1x1_k + i];
        __syncthreads();
        // conv1x1 of both heads: one warp per position (two rounds), lanes over squares
        for (int p = wp; p < kHeadPB; p += kHeadThreads / 32) {
            float acc[2][8], va0 = 0.f, va1 = 0.f;
    #pragma unroll
            for (int h = 0; h < 2; ++h)
    #pragma unroll
                for (int o = 0; o < 8; ++o) acc[h][o] = 0.f;
            if (p < npos) {
                const float *src = pconv + (base + p) * kC * 64;
                const float *tsrc = trunk + (base + p) * kC * 64;
    #pragma unroll 8
                for (int c = 0; c < kC; ++c) {
                    const float a0 = src[c * 64 + lane], a1 = src[c * 64 + 32 + lane];
                    const float wv = blob[hp.val1x1_k + c];
                    va0 = fmaf(tsrc[c * 64 + lane], wv, va0);
                    va1 = fmaf(tsrc[c * 64 + 32 + lane], wv, va1);
    #pragma unroll
                    for (int o = 0; o < 8; ++o) {
                        const float w = s_wp[c * 8 + o];
                        acc[0][o] = fmaf(a0, w, acc[0][o]);
                        acc[1][o] = fmaf(a1, w, acc[1][o]);
                    }
                }
            }
    #pragma unroll
            for (int h = 0; h < 2; ++h)
    #pragma unroll
                for (int o = 0; o < 8; ++o) {
                    const float v = fmaf(fmaxf(acc[h][o] + blob[hp.pol1x1_b + o], 0.f), fold[kScalePol1x1 + o], fold[kScalePol1x1 + 8 + o]);
                    s_flat[((h * 32 + lane) * 8 + o) * kHeadStride + p] = (p < npos) ? v : 0.f;       // flatten in (x, y, c) order
                }
            const float b = blob[hp.val1x1_b], sc = fold[kScaleVal1x1], sh = fold[kScaleVal1x1 + 1];
            s_v1[p * 64 + lane] = fmaf(fmaxf(va0 + b, 0.f), sc, sh);
            s_v1[p * 64 + lane + 32] = fmaf(fmaxf(va1 + b, 0.f), sc, sh);
        }
    }
    __syncthreads();
    // policy dense 512 -> 512 (Keras kernel [in, out]): thread = outputs (o, o + 256) x 16 positions
    {
        const int o = tid & 255, pg = (tid >> 8) * 16;
        float l0[16], l1[16];
#pragma unroll
        for (int p = 0; p < 16; ++p) { l0[p] = 0.f; l1[p] = 0.f; }
        const float *wd = blob + hp.pol_dense_k;
#pragma unroll 8
        for (int i = 0; i < 512; ++i) {
            const float w0 = __ldg(wd + i * 512 + o), w1 = __ldg(wd + i * 512 + 256 + o);
            float f[16];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float4 v = *reinterpret_cast<const float4 *>(&s_flat[i * kHeadStride + pg + 4 * q]);
                f[4 * q] = v.x; f[4 * q + 1] = v.y; f[4 * q + 2] = v.z; f[4 * q + 3] = v.w;
            }
#pragma unroll
            for (int p = 0; p < 16; ++p) { l0[p] = fmaf(f[p], w0, l0[p]); l1[p] = fmaf(f[p], w1, l1[p]); }
        }
        __syncthreads();
        const float b0 = blob[hp.pol_dense_b + o], b1 = blob[hp.pol_dense_b + 256 + o];
#pragma unroll
        for (int p = 0; p < 16; ++p) {
            s_flat[(pg + p) * 512 + o] = l0[p] + b0;
            s_flat[(pg + p) * 512 + 256 + o] = l1[p] + b1;
        }
    }
    __syncthreads();
    for (int p = wp; p < npos; p += kHeadThreads / 32) {
        // softmax, one warp per position
        float *row = s_flat + p * 512;
        float m = -INFINITY;
        for (int i = lane; i < 512; i += 32) m = fmaxf(m, row[i]);
#pragma unroll
        for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xFFFFFFFFu, m, o));
        float e[16], s = 0.f;
        if (logits_out != nullptr)
            for (int i = lane; i < 512; i += 32) logits_out[(base + p) * 512 + i] = row[i];
#pragma unroll
        for (int i = 0; i < 16; ++i) { e[i] = expf(row[lane + 32 * i] - m); s += e[i]; }
#pragma unroll
        for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xFFFFFFFFu, s, o);
        float *dst = policy + (base + p) * CK_POLICY_SIZE;
#pragma unroll
        for (int i = 0; i < 16; ++i) dst[lane + 32 * i] = e[i] / s;
        // value head: dense 64 -> 64, ReLU, BN; dense 64 -> 1, tanh
        float h0 = 0.f, h1 = 0.f;
        for (int q = 0; q < 64; ++q) {
            const float v = s_v1[p * 64 + q];
            h0 = fmaf(v, blob[hp.val_d1_k + q * 64 + lane], h0);
            h1 = fmaf(v, blob[hp.val_d1_k + q * 64 + 32 + lane], h1);
        }
        h0 = fmaf(fmaxf(h0 + blob[hp.val_d1_b + lane], 0.f), fold[kScaleValD1 + lane], fold[kScaleValD1 + 64 + lane]);
        h1 = fmaf(fmaxf(h1 + blob[hp.val_d1_b + 32 + lane], 0.f), fold[kScaleValD1 + 32 + lane], fold[kScaleValD1 + 96 + lane]);
        float t = h0 * blob[hp.val_d2_k + lane] + h1 * blob[hp.val_d2_k + 32 + lane];
#pragma unroll
        for (int o = 16; o; o >>= 1) t += __shfl_xor_sync(0xFFFFFFFFu, t, o);
        if (lane == 0) {
            value[base + p] = tanhf(t + blob[hp.val_d2_b]);
            if (vpre_out != nullptr) vpre_out[base + p] = t + blob[hp.val_d2_b];
        }
    }
}

// ---- host side -----------------------------------------------------------------------------
// scratch between the tower and the heads: per position 64 + 1024 floats on the fused path (value conv
// output; policy features + logits), two full fp32 feature maps [128][64] on the cross-check paths
int net_reserve(ck_net *net, int64_t n, bool full_maps, int lane) {
    const size_t per0 = full_maps ? (size_t)kC * 64 : 64, per1 = full_maps ? (size_t)kC * 64 : 1024;
    float *&a0 = lane ? net->d_act0b : net->d_act0, *&a1 = lane ? net->d_act1b : net->d_act1;
    size_t &f0 = lane ? net->act0b_floats : net->act0_floats, &f1 = lane ? net->act1b_floats : net->act1_floats;
    if ((size_t)n * per0 <= f0 && (size_t)n * per1 <= f1) return CK_OK;
    cudaFree(a0); cudaFree(a1);
    a0 = a1 = nullptr; f0 = f1 = 0;
    CK_CUDA(cudaMalloc(&a0, (size_t)n * per0 * sizeof(float)));
    CK_CUDA(cudaMalloc(&a1, (size_t)n * per1 * sizeof(float)));
    f0 = (size_t)n * per0; f1 = (size_t)n * per1;
    return CK_OK;
}

static int net_reserve_io(ck_net *net, int64_t n) {
    if (n <= net->io_cap) return CK_OK;
    cudaFree(net->d_leaves); cudaFree(net->d_policy); cudaFree(net->d_value);
    net->d_leaves = nullptr; net->d_policy = nullptr; net->d_value = nullptr; net->io_cap = 0;
    CK_CUDA(cudaMalloc(&net->d_leaves, (size_t)n * sizeof(ck_leaf)));
    CK_CUDA(cudaMalloc(&net->d_policy, (size_t)n * CK_POLICY_SIZE * sizeof(float)));
    CK_CUDA(cudaMalloc(&net->d_value, (size_t)n * sizeof(float)));
    net->io_cap = n;
    return CK_OK;
}

static int net_finish_weights(ck_net *net) {
    const NetLayout L = net_layout();
    if (!net->d_scale) CK_CUDA(cudaMalloc(&net->d_scale, kScaleTotal * sizeof(float)));
    std::vector<FoldJob> jobs;
    for (int i = 0; i < 8; ++i) jobs.push_back({L.conv[i].gamma, L.conv[i].beta, L.conv[i].mean, L.conv[i].var, kC, kScaleTower + i * 2 * kC});
    jobs.push_back({L.pol1x1.gamma, L.pol1x1.beta, L.pol1x1.mean, L.pol1x1.var, 8, kScalePol1x1});
    jobs.push_back({L.val1x1.gamma, L.val1x1.beta, L.val1x1.mean, L.val1x1.var, 1, kScaleVal1x1});
    jobs.push_back({L.val_d1_gamma, L.val_d1_beta, L.val_d1_mean, L.val_d1_var, 64, kScaleValD1});
    for (const FoldJob &j : jobs) fold_bn_kernel<<<1, 128>>>(net->d_blob, net->d_scale, j);
    CK_CUDA(cudaGetLastError());
    int rc = net_tc_prepare(net);
    if (rc != CK_OK) return rc;
    rc = net_ts_prepare(net);
    if (rc != CK_OK) return rc;
    rc = net_heads_tc_prepare(net);
    if (rc != CK_OK) return rc;
    CK_CUDA(cudaDeviceSynchronize());
    net->have_weights = true;
    net->weights_gen += 1;
    return CK_OK;
}

int net_forward_rows(ck_net *net, const ck_leaf *d_leaves, int64_t max_n, const int32_t *n_dev,
                     float *d_policy, float *d_value, cudaStream_t stream, int *launches, int lane) {
    if (!net->have_weights) return fail(CK_ERR_NO_NET, "ck_net: weights were never set");
    if (max_n <= 0) return CK_OK;
    static const bool tower_ss = [] { const char *v = getenv("CK_TOWER"); return v && v[0] == 's'; }();
    int rc = net_reserve(net, max_n, net->impl == CK_NET_IMPL_SIMT || tower_ss, lane);
    if (rc != CK_OK) return rc;
    rc = ensure_constants(net->device);
    if (rc != CK_OK) return rc;
    const NetLayout L = net_layout();
    const float *blob = net->d_blob;
    float *trunk = lane ? net->d_act0b : net->d_act0, *pconv = lane ? net->d_act1b : net->d_act1;
    int nl = 0;
    bool fused = false;
    if (net->impl == CK_NET_IMPL_SIMT) {
        static bool attr_done[64] = {false};      // per device: function attributes belong to the context
        const int smem128 = kC * kPlaneStride * sizeof(float), smem14 = 14 * kPlaneStride * sizeof(float);
        if (!attr_done[net->device & 63]) {
            CK_CUDA(cudaFuncSetAttribute(conv3x3_simt_kernel<kC, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem128));
            attr_done[net->device & 63] = true;
        }
        float *a = trunk, *b = pconv;
        conv3x3_simt_kernel<14, true><<<(unsigned)max_n, 256, smem14, stream>>>(
            d_leaves, nullptr, n_dev, blob + L.conv[0].kernel, blob + L.conv[0].bias, net->d_scale + kScaleTower, a);
        ++nl;
        for (int i = 1; i < 8; ++i) {
            conv3x3_simt_kernel<kC, false><<<(unsigned)max_n, 256, smem128, stream>>>(
                nullptr, a, n_dev, blob + L.conv[i].kernel, blob + L.conv[i].bias, net->d_scale + kScaleTower + i * 2 * kC, b);
            ++nl;
            if (i < 7) { float *t = a; a = b; b = t; }   // after conv6 keep a = trunk, b = policy conv output
        }
        trunk = a; pconv = b;
        CK_CUDA(cudaGetLastError());
    } else {
        // CK_TOWER=ss selects the earlier all-shared-memory tcgen05 kernel (cross-check)
        const bool use_ss = tower_ss;
        fused = !use_ss;
        // fused path: d_act0 receives the value conv1x1 output [n][64], d_act1 the policy features [n][512]
        rc = use_ss ? net_tc_tower(net, d_leaves, max_n, n_dev, trunk, pconv, stream, &nl)
                    : net_ts_tower(net, d_leaves, max_n, n_dev, pconv, trunk, stream, &nl);
        if (rc != CK_OK) return rc;
    }
    if (net->ev_after_tower) CK_CUDA(cudaEventRecord(net->ev_after_tower, stream));
    HeadParams hp{L.pol1x1.kernel, L.pol1x1.bias, L.pol_dense_k, L.pol_dense_b, L.val1x1.kernel, L.val1x1.bias,
                  L.val_d1_k, L.val_d1_b, L.val_d2_k, L.val_d2_b};
    static bool heads_attr_done[64] = {false};
    if (!heads_attr_done[net->device & 63]) {
        CK_CUDA(cudaFuncSetAttribute(heads_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kHeadSmem));
        CK_CUDA(cudaFuncSetAttribute(heads_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kHeadSmem));
        heads_attr_done[net->device & 63] = true;
    }
    const unsigned hgrid = (unsigned)((max_n + kHeadPB - 1) / kHeadPB);
    // CK_HEADS=simt keeps the fp32 CUDA-core Dense for the fused path too (cross-check)
    static const bool heads_simt = [] { const char *v = getenv("CK_HEADS"); return v && v[0] == 's'; }();
    if (fused && !heads_simt) {
        // d_act1 = [pflat n x 512 | logits n x 512 | ...], d_act0 = vconv
        rc = net_heads_tc(net, pconv, trunk, pconv + max_n * 512, max_n, n_dev, d_policy, d_value, stream, &nl);
        if (rc != CK_OK) return rc;
        if (net->d_logits_out)
            CK_CUDA(cudaMemcpyAsync(net->d_logits_out, pconv + max_n * 512, (size_t)max_n * 512 * sizeof(float), cudaMemcpyDeviceToDevice, stream));
    } else if (fused) {
        heads_kernel<true><<<hgrid, kHeadThreads, kHeadSmem, stream>>>(trunk, pconv, max_n, n_dev, blob, net->d_scale, hp, d_policy, d_value,
                                                                       net->d_logits_out, net->d_value_pre);
        ++nl;
    } else {
        heads_kernel<false><<<hgrid, kHeadThreads, kHeadSmem, stream>>>(trunk, pconv, max_n, n_dev, blob, net->d_scale, hp, d_policy, d_value,
                                                                        net->d_logits_out, net->d_value_pre);
        ++nl;
    }
    CK_CUDA(cudaGetLastError());
    if (launches) *launches += nl;
    return CK_OK;
}

// ck_net_set_weights* with the values the net already holds is a no-op: the packed tensor-core operands stay, and so
// does weights_gen, i.e. evaluations cached by the engines that use this net remain valid
__global__ void blob_differs_kernel(const uint32_t *__restrict__ a, const uint32_t *__restrict__ b, int64_t n, int32_t *flag) {
    bool diff = false;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) diff |= a[i] != b[i];
    if (diff) atomicExch(flag, 1);
}

static int net_same_weights(ck_net *net, const float *d_new, bool *same) {
    *same = false;
    if (!net->have_weights) return CK_OK;
    int32_t *flag = net->d_range_flag + 1;
    CK_CUDA(cudaMemset(flag, 0, sizeof(int32_t)));
    blob_differs_kernel<<<296, 256>>>((const uint32_t *)net->d_blob, (const uint32_t *)d_new, CK_NET_PARAM_COUNT, flag);
    int32_t h = 1;
    CK_CUDA(cudaMemcpy(&h, flag, sizeof(h), cudaMemcpyDeviceToHost));
    *same = h == 0;
    return CK_OK;
}

int net_check_range(ck_net *net) {
    int32_t flag = 0;
    CK_CUDA(cudaMemcpy(&flag, net->d_range_flag, sizeof(flag), cudaMemcpyDeviceToHost));
    if (flag == 0) return CK_OK;
    CK_CUDA(cudaMemset(net->d_range_flag, 0, sizeof(flag)));
    return fail(CK_ERR_NET_RANGE, "network activation beyond the split-fp16 range of the tensor-core path (|a| >= 4094 after a "
                                  "BatchNorm, or non-finite): outputs of this call are invalid; use CK_NET_IMPL_SIMT for these weights");
}

}  // namespace ck

using namespace ck;

extern "C" {

ck_net *ck_net_create(int device) {
    DeviceGuard g(device);
    if (!g.ok) { fail(CK_ERR_CUDA, "ck_net_create: cannot select CUDA device " + std::to_string(device)); return nullptr; }
    ck_net *net = new ck_net();
    net->device = device;
    CK_CUDA_PTR(cudaMalloc(&net->d_blob, CK_NET_PARAM_COUNT * sizeof(float)));
    CK_CUDA_PTR(cudaMalloc(&net->d_range_flag, 2 * sizeof(int32_t)));      // [0] range flag, [1] scratch of net_same_weights
    CK_CUDA_PTR(cudaMemset(net->d_range_flag, 0, 2 * sizeof(int32_t)));
    return net;
}

void ck_net_destroy(ck_net *net) {
    if (!net) return;
    DeviceGuard g(net->device);
    cudaFree(net->d_blob); cudaFree(net->d_scale); cudaFree(net->d_wpack); cudaFree(net->d_wts); cudaFree(net->d_hpack);
    cudaFree(net->d_act0); cudaFree(net->d_act1); cudaFree(net->d_act0b); cudaFree(net->d_act1b);
    cudaFree(net->d_leaves); cudaFree(net->d_policy); cudaFree(net->d_value); cudaFree(net->d_range_flag); cudaFree(net->d_stage);
    delete net;
}

int ck_net_set_impl(ck_net *net, int impl) {
    if (!net || (impl != CK_NET_IMPL_TC && impl != CK_NET_IMPL_SIMT)) return fail(CK_ERR_ARG, "ck_net_set_impl: bad arguments");
    net->impl = impl;
    return CK_OK;
}

int ck_net_set_weights(ck_net *net, const float *blob, int64_t count) {
    if (!net || !blob || count != net_layout().total || count != CK_NET_PARAM_COUNT)
        return fail(CK_ERR_ARG, "ck_net_set_weights: expected " + std::to_string(CK_NET_PARAM_COUNT) + " floats");
    DeviceGuard g(net->device);
    if (net->have_weights) {
        // stage the upload and compare on the device: unchanged weights keep the packed operands (and cached evaluations)
        if (!net->d_stage) CK_CUDA(cudaMalloc(&net->d_stage, CK_NET_PARAM_COUNT * sizeof(float)));
        CK_CUDA(cudaMemcpy(net->d_stage, blob, count * sizeof(float), cudaMemcpyHostToDevice));
        bool same = false;
        int rc = net_same_weights(net, net->d_stage, &same);
        if (rc != CK_OK) return rc;
        if (same) return CK_OK;
        CK_CUDA(cudaMemcpy(net->d_blob, net->d_stage, count * sizeof(float), cudaMemcpyDeviceToDevice));
        return net_finish_weights(net);
    }
    CK_CUDA(cudaMemcpy(net->d_blob, blob, count * sizeof(float), cudaMemcpyHostToDevice));
    return net_finish_weights(net);
}

int ck_net_set_weights_device(ck_net *net, const float *d_blob, int64_t count) {
    if (!net || !d_blob || count != CK_NET_PARAM_COUNT)
        return fail(CK_ERR_ARG, "ck_net_set_weights_device: expected " + std::to_string(CK_NET_PARAM_COUNT) + " floats");
    DeviceGuard g(net->device);
    bool same = false;
    int rc = net_same_weights(net, d_blob, &same);
    if (rc != CK_OK) return rc;
    if (same) return CK_OK;
    CK_CUDA(cudaMemcpy(net->d_blob, d_blob, count * sizeof(float), cudaMemcpyDeviceToDevice));
    return net_finish_weights(net);
}

int ck_net_forward_device(ck_net *net, const ck_leaf *d_leaves, int64_t n, float *d_policy, float *d_value, void *stream) {
    if (!net || n < 0) return fail(CK_ERR_ARG, "ck_net_forward_device: bad arguments");
    DeviceGuard g(net->device);
    return net_forward_rows(net, d_leaves, n, nullptr, d_policy, d_value, (cudaStream_t)stream, nullptr);
}

int ck_net_forward(ck_net *net, const ck_leaf *leaves, int64_t n, float *policy, float *value) {
    if (!net || !leaves || n < 0) return fail(CK_ERR_ARG, "ck_net_forward: bad arguments");
    if (n == 0) return CK_OK;
    DeviceGuard g(net->device);
    int rc = net_reserve_io(net, n);
    if (rc != CK_OK) return rc;
    CK_CUDA(cudaMemcpy(net->d_leaves, leaves, n * sizeof(ck_leaf), cudaMemcpyHostToDevice));
    rc = net_forward_rows(net, net->d_leaves, n, nullptr, net->d_policy, net->d_value, nullptr, nullptr);
    if (rc != CK_OK) return rc;
    CK_CUDA(cudaDeviceSynchronize());
    rc = net_check_range(net);
    if (rc != CK_OK) return rc;
    if (policy) CK_CUDA(cudaMemcpy(policy, net->d_policy, n * CK_POLICY_SIZE * sizeof(float), cudaMemcpyDeviceToHost));
    if (value) CK_CUDA(cudaMemcpy(value, net->d_value, n * sizeof(float), cudaMemcpyDeviceToHost));
    return CK_OK;
}

int ck_net_forward_logits(ck_net *net, const ck_leaf *leaves, int64_t n, float *policy, float *value, float *logits, float *value_pre) {
    if (!net || !leaves || n < 0) return fail(CK_ERR_ARG, "ck_net_forward_logits: bad arguments");
    if (n == 0) return CK_OK;
    DeviceGuard g(net->device);
    int rc = net_reserve_io(net, n);
    if (rc != CK_OK) return rc;
    float *d_dbg = nullptr;                                   // [n][512] logits + [n] pre-tanh values
    CK_CUDA(cudaMalloc(&d_dbg, (size_t)n * 513 * sizeof(float)));
    net->d_logits_out = d_dbg; net->d_value_pre = d_dbg + n * 512;
    cudaError_t ce = cudaMemcpy(net->d_leaves, leaves, n * sizeof(ck_leaf), cudaMemcpyHostToDevice);
    rc = ce == cudaSuccess ? net_forward_rows(net, net->d_leaves, n, nullptr, net->d_policy, net->d_value, nullptr, nullptr) : CK_ERR_CUDA;
    net->d_logits_out = nullptr; net->d_value_pre = nullptr;
    if (rc == CK_OK && cudaDeviceSynchronize() != cudaSuccess) rc = CK_ERR_CUDA;
    if (rc == CK_OK) rc = net_check_range(net);
    if (rc == CK_OK) {
        if (policy) ce = cudaMemcpy(policy, net->d_policy, n * CK_POLICY_SIZE * sizeof(float), cudaMemcpyDeviceToHost);
        if (value && ce == cudaSuccess) ce = cudaMemcpy(value, net->d_value, n * sizeof(float), cudaMemcpyDeviceToHost);
        if (logits && ce == cudaSuccess) ce = cudaMemcpy(logits, d_dbg, (size_t)n * 512 * sizeof(float), cudaMemcpyDeviceToHost);
        if (value_pre && ce == cudaSuccess) ce = cudaMemcpy(value_pre, d_dbg + n * 512, n * sizeof(float), cudaMemcpyDeviceToHost);
        if (ce != cudaSuccess) rc = CK_ERR_CUDA;
    }
    cudaFree(d_dbg);
    if (rc == CK_ERR_CUDA && ce != cudaSuccess) return fail(CK_ERR_CUDA, std::string("ck_net_forward_logits: ") + cudaGetErrorString(ce));
    return rc;
}

// debug / parity: what the tower handed to the heads in the most recent forward call of `n` positions on the
// tensor-core path -- pflat [n,512] (policy conv1x1 + ReLU + BN, flattened (x,y,c)) and vconv [n,64]
int ck_net_last_features(ck_net *net, int64_t n, float *pflat, float *vconv) {
    if (!net || n <= 0) return fail(CK_ERR_ARG, "ck_net_last_features: bad arguments");
    if (net->impl != CK_NET_IMPL_TC || (size_t)n * 1024 > net->act1_floats || (size_t)n * 64 > net->act0_floats)
        return fail(CK_ERR_STATE, "ck_net_last_features: no tensor-core forward of that size has run on this net");
    DeviceGuard g(net->device);
    CK_CUDA(cudaDeviceSynchronize());
    if (pflat) CK_CUDA(cudaMemcpy(pflat, net->d_act1, (size_t)n * 512 * sizeof(float), cudaMemcpyDeviceToHost));
    if (vconv) CK_CUDA(cudaMemcpy(vconv, net->d_act0, (size_t)n * 64 * sizeof(float), cudaMemcpyDeviceToHost));
    return CK_OK;
}

int ck_net_range_status(ck_net *net) {
    if (!net) return fail(CK_ERR_ARG, "ck_net_range_status: null net");
    DeviceGuard g(net->device);
    if (cudaDeviceSynchronize() != cudaSuccess) return fail(CK_ERR_CUDA, "ck_net_range_status: device error");
    return net_check_range(net);
}

int ck_net_forward_planes(ck_net *net, const float *x, int64_t n, float *policy, float *value) {
    if (!net || !x || n < 0) return fail(CK_ERR_ARG, "ck_net_forward_planes: bad arguments");
    if (n == 0) return CK_OK;
    DeviceGuard g(net->device);
    int rc = net_reserve_io(net, n);
    if (rc != CK_OK) return rc;
    float *d_x = nullptr;
    CK_CUDA(cudaMalloc(&d_x, n * 896 * sizeof(float)));
    cudaError_t e = cudaMemcpy(d_x, x, n * 896 * sizeof(float), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        planes_to_leaf_kernel<<<(unsigned)((n + 127) / 128), 128>>>(d_x, n, net->d_leaves);
        e = cudaDeviceSynchronize();
    }
    cudaFree(d_x);
    if (e != cudaSuccess) return fail(CK_ERR_CUDA, std::string("ck_net_forward_planes: ") + cudaGetErrorString(e));
    rc = net_forward_rows(net, net->d_leaves, n, nullptr, net->d_policy, net->d_value, nullptr, nullptr);
    if (rc != CK_OK) return rc;
    CK_CUDA(cudaDeviceSynchronize());
    rc = net_check_range(net);
    if (rc != CK_OK) return rc;
    if (policy) CK_CUDA(cudaMemcpy(policy, net->d_policy, n * CK_POLICY_SIZE * sizeof(float), cudaMemcpyDeviceToHost));
    if (value) CK_CUDA(cudaMemcpy(value, net->d_value, n * sizeof(float), cudaMemcpyDeviceToHost));
    return CK_OK;
}

}  // extern "C"
