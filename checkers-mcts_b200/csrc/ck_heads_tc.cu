// ck_heads_tc.cu -- policy head Dense(512 -> 512) on the tensor cores + the small fp32 tail of both
// heads (softmax; value Dense(64) + BN + Dense(1) + tanh).  sm_100a.
//
// Reference: training_pipeline.py:97-100 (policy Dense softmax) and :106-112 (value head tail); the
// inputs are the tower's outputs (ck_net_ts.cu): pflat [n][512] = policy conv1x1 + ReLU + BN flattened in
// (x, y, c) order, vconv [n][64] = value conv1x1 + ReLU + BN.
//
// heads_dense_tc_kernel: logits[pos][out] = sum_in pflat[pos][in] * W[in][out] + b[out] as a GEMM with
// M = 128 positions (TMEM lanes), N = 128 outputs per CTA (grid.y = 4 output quarters), K = 512 in four
// chunks of 128.  Same split-fp16 scheme as the tower (Ahi*Bhi + Ahi*Blo + Alo*Bhi, fp32 accumulation
// in TMEM): the activation chunk is converted from fp32 by the CTA's threads straight into the UMMA
// K-major core-matrix layout, the pre-packed weight chunk (64 KB, hi + lo) arrives by one bulk copy.
// The fp32 CUDA-core version (heads_kernel) needs ~35 us of FMA issue for a 4096-position batch and
// measured 86 us; this one takes 25 us and is latency-bound (tensor pipe 10 % active: each CTA waits for its
// activation rows chunk by chunk; a two-stage version that converts chunk k+1 while chunk k multiplies
// measured the same, the loads would have to be prefetched deeper).
// heads_finish_kernel: one warp per position, softmax over the 512 logits and the value head tail.
#include <cuda_fp16.h>
#include "ck_net.cuh"
#include "ck_tc_ptx.cuh"

namespace ck {
namespace htc {

using namespace ck::ptx;

constexpr int kM = 128, kN = 128, kKC = 128;        // positions per CTA, outputs per CTA, K per chunk
constexpr int kSplit = kM * kKC * 2;                // bytes of one operand half (hi or lo) of a chunk = 32 KB
constexpr int kChunkBytes = 2 * kSplit;             // 64 KB
constexpr int kThreads = 256;
constexpr int kSmem = 2 * kChunkBytes + 64;
constexpr float kActScale = 16.0f;
constexpr size_t kPackBytes = (size_t)4 * 4 * kChunkBytes;   // [out quarter][k chunk][hi|lo][k/8][out 128][8]
constexpr uint32_t kIdesc = make_idesc_f16(kM, kN);
// The tensor pipe rounds its fp32 accumulator toward zero after every MMA, so a long accumulation chain loses about
// half an ulp per instruction in one direction (measured on the reference's trained networks: logits off by 1e-5
// relative with one 96-instruction chain).  Three accumulators keep the chains short where the values are large: the
// hi*hi products of K chunks {0,1} and {2,3} (16 instructions each) and all hi*lo + lo*hi cross terms (2^-11 of the
// magnitude, so their 64 roundings do not matter); the epilogue adds them in fp32 round-to-nearest.
constexpr int kTmemCols = 512;                      // 3 x 128 used (allocations are powers of two)
constexpr uint32_t kColMain0 = 0, kColMain1 = kN, kColCross = 2 * kN;

__global__ void __launch_bounds__(kThreads, 1)
heads_dense_tc_kernel(const float *__restrict__ pflat, int64_t max_n, const int32_t *__restrict__ n_dev,
                      const uint8_t *__restrict__ wpack, const float *__restrict__ bias, const float *__restrict__ inv_scale,
                      float *__restrict__ logits, int32_t *__restrict__ range_flag) {
    extern __shared__ __align__(1024) uint8_t smem[];
    int64_t n = max_n;
    if (n_dev != nullptr) n = min((int64_t)*n_dev, max_n);
    const int64_t base = (int64_t)blockIdx.x * kM;
    if (base >= n) return;
    const int nq = blockIdx.y;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint8_t *s_a = smem, *s_b = smem + kChunkBytes;
    const uint32_t bar_b = smem_u32(smem + 2 * kChunkBytes), bar_mma = bar_b + 8;
    uint32_t *s_tmem = reinterpret_cast<uint32_t *>(smem + 2 * kChunkBytes + 16);
    if (warp == 0) {
        if (lane == 0) { mbar_init(bar_b, 1); mbar_init(bar_mma, 1); mbar_init_fence(); }
        __syncwarp();
        tmem_alloc<kTmemCols>(smem_u32(s_tmem));
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *s_tmem;

    for (int kc = 0; kc < 4; ++kc) {
        if (kc > 0) { mbar_wait(bar_mma, (uint32_t)((kc - 1) & 1)); tc_fence_after(); }     // chunk buffers are free again
        if (tid == 0) {
            mbar_expect_tx(bar_b, kChunkBytes);
            bulk_g2s(smem_u32(s_b), wpack + ((size_t)nq * 4 + kc) * kChunkBytes, kChunkBytes, bar_b);
        }
        // activations: (position m, 8-element K unit u) -> 16 B hi + 16 B lo in the K-major core-matrix layout
        for (int idx = tid; idx < kM * (kKC / 8); idx += kThreads) {
            const int m = idx & (kM - 1), u = idx >> 7;
            float v[8];
            if (base + m < n) {
                const float4 *src = reinterpret_cast<const float4 *>(pflat + (base + m) * 512 + kc * kKC + u * 8);
                const float4 a = __ldg(src), b = __ldg(src + 1);
                v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
            } else {
#pragma unroll
                for (int e = 0; e < 8; ++e) v[e] = 0.f;
            }
            __half hi[8], lo[8];
            bool bad = false;
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const float s = v[e] * kActScale;
                bad |= !(fabsf(s) < 65504.f);
                hi[e] = __float2half_rn(s);
                lo[e] = __float2half_rn(s - __half2float(hi[e]));
            }
            if (bad) atomicExch(range_flag, 1);                     // beyond fp16 (or NaN from the tower): never silent
            uint8_t *dst = s_a + u * (kM * 16) + m * 16;
            *reinterpret_cast<uint4 *>(dst) = *reinterpret_cast<const uint4 *>(hi);
            *reinterpret_cast<uint4 *>(dst + kSplit) = *reinterpret_cast<const uint4 *>(lo);
        }
        fence_proxy_async();
        __syncthreads();
        if (warp == 0) {
            mbar_wait(bar_b, (uint32_t)(kc & 1));
            tc_fence_after();
            if (lane == 0) {
                const uint32_t a0 = smem_u32(s_a), b0 = smem_u32(s_b);
#pragma unroll
                for (int ks = 0; ks < kKC / 16; ++ks) {
                    const uint64_t a_hi = make_desc(a0 + ks * 2 * (kM * 16), kM * 16, 128);
                    const uint64_t a_lo = make_desc(a0 + kSplit + ks * 2 * (kM * 16), kM * 16, 128);
                    const uint64_t b_hi = make_desc(b0 + ks * 2 * (kN * 16), kN * 16, 128);
                    const uint64_t b_lo = make_desc(b0 + kSplit + ks * 2 * (kN * 16), kN * 16, 128);
                    tc_mma_ss(tmem + ((kc >> 1) ? kColMain1 : kColMain0), a_hi, b_hi, kIdesc, ((kc & 1) | ks) != 0 ? 1u : 0u);
                    tc_mma_ss(tmem + kColCross, a_hi, b_lo, kIdesc, (kc | ks) != 0 ? 1u : 0u);
                    tc_mma_ss(tmem + kColCross, a_lo, b_hi, kIdesc, 1u);
                }
                tc_commit(bar_mma);
            }
            __syncwarp();
        }
    }
    mbar_wait(bar_mma, 1u);                           // fourth completion: phases 0,1,0,1
    tc_fence_after();
    {
        // epilogue: lane quadrant = warp % 4 (positions), column half = warp / 4 (64 outputs)
        const int quad = warp & 3, half = warp >> 2;
        const int64_t pos = base + quad * 32 + lane;
        const float inv = inv_scale[0];
        const float *bptr = bias + nq * kN + half * 64;
        float *out = logits + pos * 512 + nq * kN + half * 64;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            float v[16], v1[16], vc[16];
            const uint32_t taddr = tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)(half * 64 + q * 16);
            tmem_ld16(taddr + kColMain0, v);
            tmem_ld16(taddr + kColMain1, v1);
            tmem_ld16(taddr + kColCross, vc);
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = (v[i] + v1[i]) + vc[i];
            if (pos < n) {
#pragma unroll
                for (int i = 0; i < 16; i += 4)
                    *reinterpret_cast<float4 *>(out + q * 16 + i) =
                        make_float4(fmaf(v[i], inv, bptr[q * 16 + i]), fmaf(v[i + 1], inv, bptr[q * 16 + i + 1]),
                                    fmaf(v[i + 2], inv, bptr[q * 16 + i + 2]), fmaf(v[i + 3], inv, bptr[q * 16 + i + 3]));
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<kTmemCols>(tmem);
}

struct FinishParams { int64_t val_d1_k, val_d1_b, val_d2_k, val_d2_b; };

__global__ void __launch_bounds__(256)
heads_finish_kernel(const float *__restrict__ logits, const float *__restrict__ vconv, int64_t max_n,
                    const int32_t *__restrict__ n_dev, const float *__restrict__ blob, const float *__restrict__ fold,
                    FinishParams hp, float *__restrict__ policy, float *__restrict__ value, float *__restrict__ value_pre) {
    int64_t n = max_n;
    if (n_dev != nullptr) n = min((int64_t)*n_dev, max_n);
    const int lane = threadIdx.x & 31;
    const int64_t p = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (p >= n) return;
    // softmax (Dense activation, training_pipeline.py:98)
    const float *row = logits + p * 512;
    float x[16], m = -INFINITY;
#pragma unroll
    for (int i = 0; i < 16; ++i) { x[i] = row[lane + 32 * i]; m = fmaxf(m, x[i]); }
#pragma unroll
    for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xFFFFFFFFu, m, o));
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) { x[i] = expf(x[i] - m); s += x[i]; }
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xFFFFFFFFu, s, o);
    float *dst = policy + p * CK_POLICY_SIZE;
#pragma unroll
    for (int i = 0; i < 16; ++i) dst[lane + 32 * i] = x[i] / s;
    // value head: Dense 64 -> 64, ReLU, BN; Dense 64 -> 1, tanh (:106-112)
    const float v0 = vconv[p * 64 + lane], v1 = vconv[p * 64 + 32 + lane];
    float h0 = 0.f, h1 = 0.f;
#pragma unroll 8
    for (int q = 0; q < 64; ++q) {
        const float v = __shfl_sync(0xFFFFFFFFu, q < 32 ? v0 : v1, q & 31);
        h0 = fmaf(v, blob[hp.val_d1_k + q * 64 + lane], h0);
        h1 = fmaf(v, blob[hp.val_d1_k + q * 64 + 32 + lane], h1);
    }
    h0 = fmaf(fmaxf(h0 + blob[hp.val_d1_b + lane], 0.f), fold[kScaleValD1 + lane], fold[kScaleValD1 + 64 + lane]);
    h1 = fmaf(fmaxf(h1 + blob[hp.val_d1_b + 32 + lane], 0.f), fold[kScaleValD1 + 32 + lane], fold[kScaleValD1 + 96 + lane]);
    float t = h0 * blob[hp.val_d2_k + lane] + h1 * blob[hp.val_d2_k + 32 + lane];
#pragma unroll
    for (int o = 16; o; o >>= 1) t += __shfl_xor_sync(0xFFFFFFFFu, t, o);
    if (lane == 0) {
        value[p] = tanhf(t + blob[hp.val_d2_b]);
        if (value_pre != nullptr) value_pre[p] = t + blob[hp.val_d2_b];
    }
}

// ---- weight packing ----------------------------------------------------------------------------
__global__ void dense_scale_kernel(const float *__restrict__ w, float *__restrict__ scale_out) {
    float m = 0.f;
    for (int i = threadIdx.x; i < 512 * 512; i += blockDim.x) m = fmaxf(m, fabsf(w[i]));
    __shared__ float s_m[256];
    s_m[threadIdx.x] = m;
    __syncthreads();
    for (int o = 128; o; o >>= 1) { if ((int)threadIdx.x < o) s_m[threadIdx.x] = fmaxf(s_m[threadIdx.x], s_m[threadIdx.x + o]); __syncthreads(); }
    if (threadIdx.x == 0) {
        m = s_m[0];
        int e = 0;
        if (m > 0.f && isfinite(m)) { int ex; frexpf(m, &ex); e = 14 - ex; }
        if (e > 40) e = 40;
        if (e < -40) e = -40;
        const float S = ldexpf(1.0f, e);
        scale_out[0] = 1.0f / (S * kActScale);
        scale_out[1] = S;
    }
}

// Keras Dense kernel [in][out] -> [out quarter][k chunk][hi|lo][k/8][out 128][8 halfs]
__global__ void dense_pack_kernel(const float *__restrict__ w, const float *__restrict__ scale, __half *__restrict__ out) {
    const float S = scale[1];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 512 * 512; i += gridDim.x * blockDim.x) {
        const int e = i & 7, o = (i >> 3) & 127, u = (i >> 10) & 15, kc = (i >> 14) & 3, nq = i >> 16;
        const int in = kc * kKC + u * 8 + e, outc = nq * kN + o;
        const float v = w[in * 512 + outc] * S;
        const __half hi = __float2half_rn(v);
        __half *blk = out + ((size_t)nq * 4 + kc) * (kChunkBytes / 2);
        blk[u * (kN * 8) + o * 8 + e] = hi;
        blk[kSplit / 2 + u * (kN * 8) + o * 8 + e] = __float2half_rn(v - __half2float(hi));
    }
}

}  // namespace htc

int net_heads_tc_prepare(ck_net *net) {
    const NetLayout L = net_layout();
    if (!net->d_hpack) CK_CUDA(cudaMalloc(&net->d_hpack, htc::kPackBytes + 64));
    float *aux = (float *)((uint8_t *)net->d_hpack + htc::kPackBytes);
    htc::dense_scale_kernel<<<1, 256>>>(net->d_blob + L.pol_dense_k, aux);
    htc::dense_pack_kernel<<<512, 256>>>(net->d_blob + L.pol_dense_k, aux, (__half *)net->d_hpack);
    CK_CUDA(cudaGetLastError());
    return CK_OK;
}

// pflat [n][512], vconv [n][64] -> policy [n][512] (softmax), value [n]; logits = scratch [n][512]
int net_heads_tc(ck_net *net, const float *d_pflat, const float *d_vconv, float *d_logits, int64_t max_n, const int32_t *n_dev,
                 float *d_policy, float *d_value, cudaStream_t stream, int *launches) {
    if (!net->d_hpack) return fail(CK_ERR_NO_NET, "tensor-core heads: weights were never packed");
    const NetLayout L = net_layout();
    static bool attr_done[64] = {false};          // per device: function attributes belong to the context
    if (!attr_done[net->device & 63]) {
        CK_CUDA(cudaFuncSetAttribute(htc::heads_dense_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, htc::kSmem));
        attr_done[net->device & 63] = true;
    }
    const float *aux = (const float *)((const uint8_t *)net->d_hpack + htc::kPackBytes);
    const dim3 grid((unsigned)((max_n + htc::kM - 1) / htc::kM), 4);
    htc::heads_dense_tc_kernel<<<grid, htc::kThreads, htc::kSmem, stream>>>(
        d_pflat, max_n, n_dev, (const uint8_t *)net->d_hpack, net->d_blob + L.pol_dense_b, aux, d_logits, net->d_range_flag);
    htc::FinishParams hp{L.val_d1_k, L.val_d1_b, L.val_d2_k, L.val_d2_b};
    htc::heads_finish_kernel<<<(unsigned)((max_n + 7) / 8), 256, 0, stream>>>(
        d_logits, d_vconv, max_n, n_dev, net->d_blob, net->d_scale, hp, d_policy, d_value, net->d_value_pre);
    CK_CUDA(cudaGetLastError());
    if (launches) *launches += 2;
    return CK_OK;
}

}  // namespace ck
