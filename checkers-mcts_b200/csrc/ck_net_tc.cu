// ck_net_tc.cu -- K3 tower on 5th-gen tensor cores (tcgen05 + TMEM), sm_100a: the first, all-shared-memory
// kernel (both MMA operands from shared memory, one tile per CTA).  Superseded as the product path by
// ck_net_ts.cu (weights in TMEM, two tiles per CTA); kept as an independent cross-check (CK_TOWER=ss).
//
// All eight 3x3 convolutions of create_nn (reference training_pipeline.py:57-88: conv0..conv6
// and the policy head's 3x3 conv; 98 % of the network's 134.87 MFLOP/position) run inside ONE
// persistent kernel.  A CTA owns a tile of P positions and carries their activations through
// all layers without leaving the SM:
//
//   * Orientation: D[co][square] = sum_k W[co][k] * act[k][square] -- the weights are the
//     M = 128 operand (A), one position's 64 squares are the N = 64 operand (B).
//   * Activations live in shared memory as split fp16 (hi + lo, 2 x 11 significand bits) in
//     the UMMA K-major no-swizzle core-matrix layout over ZERO-PADDED 10 x 10 boards whose
//     rows are INTERLEAVED across the P positions of the tile:
//     byte offset = chunk(ci/8) * LBO + ((P*row + p)*10 + col) * 16 + (ci%8)*2.
//     Eight board columns are the eight 16-byte rows of a core matrix; the next 8-row group
//     (SBO = 160 B) is the same board row of the next position, so ONE MMA covers all P
//     positions (N = 64*P) and a 3x3 tap is just a different descriptor start address
//     ((P*kh*10 + kw) * 16 bytes).  No im2col, no halo logic, and the weight operand is read
//     from shared memory once per P positions (at N = 64 the SS-mode MMA was shared-memory
//     bound: profiles/r1a_*).
//   * Weights stream from L2 through a ring of 32 KB stages with 1-D bulk async copies
//     (cp.async.bulk + mbarrier complete_tx); they are pre-packed on the device, once per
//     ck_net_set_weights, as split fp16 core matrices in exactly the stage order.
//   * fp32-grade accuracy (north_star: 1e-5) from three kind::f16 MMAs per k-step:
//     Whi*Ahi + Whi*Alo + Wlo*Ahi, accumulated in fp32 in TMEM (the dropped Wlo*Alo term is
//     2^-22 relative).  Weights are scaled per layer by a power of two into fp16's sweet
//     spot, activations by 2^4; the epilogue undoes both exactly.
//   * Epilogue (8 warps, one TMEM lane = one output channel per thread): tcgen05.ld,
//     bias + ReLU + folded BatchNorm, re-split to fp16 hi/lo and written IN PLACE into the
//     activation buffer (the accumulators in TMEM are the second buffer), fp32 copies of
//     conv6 / policy-conv outputs go to HBM for the heads kernel.
//   * Warp roles: warp 0 producer (bulk copies), warp 1 MMA issuer + TMEM allocator,
//     warps 2-9 epilogue (two warps per TMEM lane quadrant).  mbarriers: full/empty per weight stage, acc_full (MMA->epilogue),
//     act_ready (epilogue->MMA).
#include <cuda_fp16.h>
#include <stdlib.h>
#include "ck_net.cuh"
#include "ck_tc_ptx.cuh"

namespace ck {
namespace tc {

using namespace ck::ptx;

constexpr int kStageBytes = 32768;                 // 128 co x 64 ci x (hi + lo) fp16
constexpr int kEpiWarps = 8;                        // two per TMEM lane quadrant, each takes half of the columns
constexpr int kThreads = 64 + 32 * kEpiWarps;
constexpr float kActScale = 16.0f;                 // activations are stored as a * 2^4
constexpr int kLayer0Stages = 9, kLayerStages = 18;
constexpr int kLayer0StageBytes = 8192;            // one tap: 128 co x 16 ci x (hi + lo)
constexpr size_t kLayer0Bytes = (size_t)kLayer0Stages * kLayer0StageBytes;
constexpr size_t kLayerBytes = (size_t)kLayerStages * kStageBytes;
constexpr size_t kPackBytes = kLayer0Bytes + 7 * kLayerBytes;

template <int P, int S> struct Cfg {
    static constexpr int kStages = S;                           // weight ring depth
    static constexpr int kChunkStride = P * 1600 + 16;          // P x 100 padded squares x 16 B + 16 B pad (bank spread)
    static constexpr int kSplitBytes = 16 * kChunkStride;       // all 128 channels, hi or lo
    static constexpr int kActBytes = 2 * kSplitBytes;
    static constexpr int kN = 64 * P;                           // MMA N: all positions of the tile
    // kind::f16 instruction descriptor (cute::UMMA::InstrDescriptor): c_format F32 (bit 4), A and B
    // F16 (0), both K-major, N>>3 at [17,23), M>>4 at [24,29)
    static constexpr uint32_t kIdesc = (1u << 4) | ((uint32_t)(kN >> 3) << 17) | ((128u >> 4) << 24);
    // byte offset of (position p, padded row r, padded column c) inside one chunk
    __host__ __device__ static constexpr int sq_off(int p, int r, int c) { return ((P * r + p) * 10 + c) * 16; }
    static constexpr int kTmemCols = P * 64 <= 32 ? 32 : P * 64 <= 64 ? 64 : P * 64 <= 128 ? 128 : P * 64 <= 256 ? 256 : 512;
    static constexpr int kBarOff = kActBytes + kStages * kStageBytes;
    static constexpr int kSmem = kBarOff + 256;
};

struct TowerParams {
    const ck_leaf *leaves;
    const int32_t *n_dev;
    int32_t max_n;
    const uint8_t *wpack;        // split-fp16 weights in stage order
    const float *blob;           // Keras-ordered fp32 parameters (biases)
    const float *fold;           // folded BN scale/shift table
    const float *inv_scale;      // per layer 1 / (weight scale * activation scale)
    int64_t bias_off[8];
    float *trunk, *pconv;        // fp32 [n][128][64]
    const float *plane5;         // 81-entry float32(n/80) table
};

__device__ __forceinline__ void stage_info(int layer, int st, size_t &off, uint32_t &bytes, int &ksteps, int &tap, int &chunk0) {
    if (layer == 0) {
        off = (size_t)st * kLayer0StageBytes; bytes = kLayer0StageBytes; ksteps = 1; tap = st; chunk0 = 0;
    } else {
        off = kLayer0Bytes + (size_t)(layer - 1) * kLayerBytes + (size_t)st * kStageBytes;
        bytes = kStageBytes; ksteps = 4; tap = st >> 1; chunk0 = (st & 1) * 8;
    }
}

template <int P, int S>
__global__ void __launch_bounds__(kThreads, 1)
tower_tc_kernel(const TowerParams prm) {
    using C = Cfg<P, S>;
    constexpr int kStages = S;
    extern __shared__ __align__(1024) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    int n = prm.max_n;
    if (prm.n_dev != nullptr) n = min(n, *prm.n_dev);
    const int n_tiles = (n + P - 1) / P;
    if ((int)blockIdx.x >= n_tiles) return;

    uint8_t *s_act = smem;
    uint8_t *s_w = smem + C::kActBytes;
    const uint32_t bar0 = smem_u32(smem + C::kBarOff);
    auto bar_full = [&](int s) { return bar0 + 8u * s; };
    auto bar_empty = [&](int s) { return bar0 + 8u * (kStages + s); };
    const uint32_t bar_acc_full = bar0 + 8u * (2 * kStages), bar_act_ready = bar0 + 8u * (2 * kStages + 1);
    uint32_t *s_tmem = reinterpret_cast<uint32_t *>(smem + C::kBarOff + 8 * (2 * kStages + 2));

    for (int i = tid * 16; i < C::kActBytes; i += kThreads * 16) *reinterpret_cast<uint4 *>(s_act + i) = make_uint4(0, 0, 0, 0);
    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < kStages; ++s) { mbar_init(bar_full(s), 1); mbar_init(bar_empty(s), 1); }
            mbar_init(bar_acc_full, 1);
            mbar_init(bar_act_ready, 32 * kEpiWarps);
            mbar_init_fence();
        }
        __syncwarp();
        tmem_alloc<C::kTmemCols>(smem_u32(s_tmem));
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *s_tmem;

    if (warp == 0) {
        // ===== producer: stream every weight stage of every layer through the ring =====
        if (lane == 0) {
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                for (int layer = 0; layer < kTowerConvs; ++layer) {
                    const int ns = layer == 0 ? kLayer0Stages : kLayerStages;
                    for (int st = 0; st < ns; ++st, ++it) {
                        const int slot = it % kStages;
                        mbar_wait(bar_empty(slot), ((it / kStages) & 1u) ^ 1u);
                        size_t off; uint32_t bytes; int ks, tap, c0;
                        stage_info(layer, st, off, bytes, ks, tap, c0);
                        mbar_expect_tx(bar_full(slot), bytes);
                        bulk_g2s(smem_u32(s_w + slot * kStageBytes), prm.wpack + off, bytes, bar_full(slot));
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        uint32_t it = 0, ar_phase = 0;
        const uint32_t act_base = smem_u32(s_act);
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            for (int layer = 0; layer < kTowerConvs; ++layer) {
                mbar_wait(bar_act_ready, ar_phase); ar_phase ^= 1u;
                tc_fence_after();
                const int ns = layer == 0 ? kLayer0Stages : kLayerStages;
                for (int st = 0; st < ns; ++st, ++it) {
                    const int slot = it % kStages;
                    mbar_wait(bar_full(slot), (it / kStages) & 1u);
                    tc_fence_after();
                    if (lane == 0) {
                        size_t off; uint32_t bytes; int ksteps, tap, chunk0;
                        stage_info(layer, st, off, bytes, ksteps, tap, chunk0);
                        const uint32_t a_base = smem_u32(s_w + slot * kStageBytes);
                        const uint32_t a_split = bytes >> 1;                       // hi block, then lo block
                        const uint32_t tap_off = (uint32_t)C::sq_off(0, tap / 3, tap % 3);
                        for (int ks = 0; ks < ksteps; ++ks) {
                            const uint64_t a_hi = make_desc(a_base + ks * 4096, 2048, 128);
                            const uint64_t a_lo = make_desc(a_base + a_split + ks * 4096, 2048, 128);
                            const uint32_t b_addr = act_base + (chunk0 + 2 * ks) * C::kChunkStride + tap_off;
                            const uint64_t b_hi = make_desc(b_addr, C::kChunkStride, 160);
                            const uint64_t b_lo = make_desc(b_addr + C::kSplitBytes, C::kChunkStride, 160);
                            tc_mma_ss(tmem_base, a_hi, b_hi, C::kIdesc, (st | ks) != 0 ? 1u : 0u);
                            tc_mma_ss(tmem_base, a_hi, b_lo, C::kIdesc, 1u);
                            tc_mma_ss(tmem_base, a_lo, b_hi, C::kIdesc, 1u);
                        }
                        tc_commit(bar_empty(slot));                                // frees the weight stage when the MMAs retire
                        if (st == ns - 1) tc_commit(bar_acc_full);                 // layer complete -> epilogue
                    }
                    __syncwarp();
                }
            }
        }
    } else {
        // ===== epilogue warps: TMEM lane quadrant = warp % 4, one output channel per thread =====
        const int et = tid - 64;                   // 0..255
        const int quad = warp & 3;                 // TMEM lanes 32*quad .. 32*quad+31 (hardware: warp id % 4)
        const int half = (warp - 2) >> 2;          // which half of the accumulator columns this warp drains
        const int co = quad * 32 + lane;
        const uint32_t t_lane = tmem_base + ((uint32_t)(quad * 32) << 16);
        uint32_t af_phase = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            // layer-0 input: planes 0..13 (+2 zero channels) of P positions, thread = (square, chunk)
            {
                const int sq = et & 63, ch = (et >> 6) & 1, x = sq >> 3, y = sq & 7;
                const bool dark = ((x ^ y) & 1) != 0;
                const uint32_t bit = 1u << (4 * x + (y >> 1));
                for (int p = et >> 7; p < P; p += 2) {
                    const int pos = tile * P + p;
                    float v[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) v[e] = 0.f;
                    if (pos < n) {
                        const ck_leaf L = prm.leaves[pos];
                        if (ch == 0) {
                            if (dark) {
                                v[0] = (L.p1 & ~L.k & bit) ? 1.f : 0.f; v[1] = (L.p1 & L.k & bit) ? 1.f : 0.f;
                                v[2] = (L.p2 & ~L.k & bit) ? 1.f : 0.f; v[3] = (L.p2 & L.k & bit) ? 1.f : 0.f;
                                v[6] = (L.mask[0] & bit) ? 1.f : 0.f; v[7] = (L.mask[1] & bit) ? 1.f : 0.f;
                            }
                            v[4] = (float)(L.info & 1u);
                            v[5] = prm.plane5[(L.info >> 8) & 0xFFu];
                        } else if (dark) {
#pragma unroll
                            for (int e = 0; e < 6; ++e) v[e] = (L.mask[2 + e] & bit) ? 1.f : 0.f;
                        }
                    }
                    __half hi[8], lo[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const float a = v[e] * kActScale;
                        hi[e] = __float2half_rn(a);
                        lo[e] = __float2half_rn(a - __half2float(hi[e]));
                    }
                    uint8_t *dst = s_act + ch * C::kChunkStride + C::sq_off(p, x + 1, y + 1);
                    *reinterpret_cast<uint4 *>(dst) = *reinterpret_cast<const uint4 *>(hi);
                    *reinterpret_cast<uint4 *>(dst + C::kSplitBytes) = *reinterpret_cast<const uint4 *>(lo);
                }
            }
            fence_proxy_async();
            tc_fence_before();
            mbar_arrive(bar_act_ready);
            for (int layer = 0; layer < kTowerConvs; ++layer) {
                const float bias = prm.blob[prm.bias_off[layer] + co];
                const float sc = prm.fold[kScaleTower + layer * 2 * kC + co], sh = prm.fold[kScaleTower + layer * 2 * kC + kC + co];
                const float inv = prm.inv_scale[layer];
                mbar_wait(bar_acc_full, af_phase); af_phase ^= 1u;
                tc_fence_after();
                // accumulator column j = (P*x + p)*8 + y: 16 columns = two (board row, position) groups
                uint8_t *abase = s_act + (co >> 3) * C::kChunkStride + (co & 7) * 2;
                float *gbase = nullptr;
                if (layer >= 6) gbase = (layer == 6 ? prm.trunk : prm.pconv) + ((int64_t)tile * P * kC + co) * 64;
#pragma unroll
                for (int q = half; q < C::kN / 16; q += 2) {
                    float v[16];
                    tmem_ld16(t_lane + (uint32_t)(q * 16), v);
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] = fmaf(fmaxf(fmaf(v[i], inv, bias), 0.f), sc, sh);
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int g = 2 * q + h, x = g / P, p = g % P;
                        if (gbase != nullptr && tile * P + p < n) {
                            float *gout = gbase + (int64_t)p * kC * 64 + x * 8;
                            *reinterpret_cast<float4 *>(gout) = make_float4(v[8 * h], v[8 * h + 1], v[8 * h + 2], v[8 * h + 3]);
                            *reinterpret_cast<float4 *>(gout + 4) = make_float4(v[8 * h + 4], v[8 * h + 5], v[8 * h + 6], v[8 * h + 7]);
                        }
                        if (layer < 7) {
#pragma unroll
                            for (int y = 0; y < 8; ++y) {
                                const float a = v[8 * h + y] * kActScale;
                                const __half hh = __float2half_rn(a);
                                const __half ll = __float2half_rn(a - __half2float(hh));
                                uint8_t *dst = abase + C::sq_off(p, x + 1, y + 1);
                                *reinterpret_cast<__half *>(dst) = hh;
                                *reinterpret_cast<__half *>(dst + C::kSplitBytes) = ll;
                            }
                        }
                    }
                }
                if (layer < 7) {
                    fence_proxy_async();
                    tc_fence_before();
                    mbar_arrive(bar_act_ready);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tmem_dealloc<C::kTmemCols>(tmem_base);
    }
}

// ---- weight packing (once per ck_net_set_weights) ---------------------------------------------
// per layer: power-of-two scale S with max|w| * S in [8192, 16384)
__global__ void wscale_kernel(const float *__restrict__ blob, const int64_t *__restrict__ koff, const int *__restrict__ cin,
                              float *__restrict__ wscale, float *__restrict__ inv_scale) {
    const int layer = blockIdx.x;
    const int64_t n = (int64_t)9 * cin[layer] * kC;
    const float *w = blob + koff[layer];
    float m = 0.f;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) m = fmaxf(m, fabsf(w[i]));
    __shared__ float s_m[256];
    s_m[threadIdx.x] = m;
    __syncthreads();
    for (int o = 128; o; o >>= 1) { if ((int)threadIdx.x < o) s_m[threadIdx.x] = fmaxf(s_m[threadIdx.x], s_m[threadIdx.x + o]); __syncthreads(); }
    if (threadIdx.x == 0) {
        m = s_m[0];
        int e = 0;
        if (m > 0.f && isfinite(m)) { int ex; frexpf(m, &ex); e = 14 - ex; }    // m = f * 2^ex, f in [0.5,1) -> m*2^e in [8192,16384)
        if (e > 40) e = 40;
        if (e < -40) e = -40;
        const float S = ldexpf(1.0f, e);
        wscale[layer] = S;
        inv_scale[layer] = 1.0f / (S * kActScale);
    }
}

// one thread per packed (hi, lo) pair: [stage][split][chunk][co][8]
__global__ void wpack_kernel(const float *__restrict__ blob, const int64_t *__restrict__ koff, const float *__restrict__ wscale,
                             __half *__restrict__ out) {
    const int64_t total = (int64_t)(kPackBytes / 4);          // hi/lo pairs
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int layer, st, cin, ksteps, tap, chunk0;
        int64_t r, stage_pairs, stage_off_bytes;
        const int64_t l0_pairs = (int64_t)(kLayer0Bytes / 4);
        if (i < l0_pairs) {
            layer = 0; cin = 14; ksteps = 1; stage_pairs = kLayer0StageBytes / 4;
            st = (int)(i / stage_pairs); r = i % stage_pairs; tap = st; chunk0 = 0;
            stage_off_bytes = (int64_t)st * kLayer0StageBytes;
        } else {
            const int64_t j = i - l0_pairs, lp = (int64_t)(kLayerBytes / 4);
            layer = 1 + (int)(j / lp); cin = kC; ksteps = 4; stage_pairs = kStageBytes / 4;
            const int64_t jj = j % lp;
            st = (int)(jj / stage_pairs); r = jj % stage_pairs; tap = st >> 1; chunk0 = (st & 1) * 8;
            stage_off_bytes = (int64_t)kLayer0Bytes + (int64_t)(layer - 1) * kLayerBytes + (int64_t)st * kStageBytes;
        }
        // r indexes [chunk (2*ksteps)][co (128)][e (8)]
        const int e = (int)(r & 7), co = (int)((r >> 3) & 127), ch = (int)(r >> 10);
        const int ci = (chunk0 + ch) * 8 + e;
        float w = 0.f;
        if (ci < cin) w = blob[koff[layer] + ((int64_t)tap * cin + ci) * kC + co] * wscale[layer];
        const __half hi = __float2half_rn(w);
        const __half lo = __float2half_rn(w - __half2float(hi));
        __half *stage = out + stage_off_bytes / 2;
        stage[r] = hi;
        stage[(int64_t)ksteps * 2 * 128 * 8 + r] = lo;
    }
}

}  // namespace tc

struct TcAux { int64_t koff[8]; int cin[8]; };

int net_tc_prepare(ck_net *net) {
    const NetLayout L = net_layout();
    if (!net->d_wpack) {
        // [packed weights][wscale 8 f][inv 8 f][plane5 81 f][koff 8 i64][cin 8 i32]
        net->wpack_bytes = tc::kPackBytes + 4096;
        CK_CUDA(cudaMalloc(&net->d_wpack, net->wpack_bytes));
    }
    uint8_t *aux = (uint8_t *)net->d_wpack + tc::kPackBytes;
    float *d_wscale = (float *)aux, *d_inv = d_wscale + 8, *d_p5 = d_inv + 8;
    int64_t *d_koff = (int64_t *)(aux + 1024);
    int *d_cin = (int *)(aux + 1024 + 64);
    TcAux h;
    for (int i = 0; i < 8; ++i) { h.koff[i] = L.conv[i].kernel; h.cin[i] = L.conv[i].cin; }
    float p5[81];
    for (int i = 0; i <= 80; ++i) p5[i] = (float)((double)i / 80.0);
    CK_CUDA(cudaMemcpy(d_koff, h.koff, sizeof(h.koff), cudaMemcpyHostToDevice));
    CK_CUDA(cudaMemcpy(d_cin, h.cin, sizeof(h.cin), cudaMemcpyHostToDevice));
    CK_CUDA(cudaMemcpy(d_p5, p5, sizeof(p5), cudaMemcpyHostToDevice));
    tc::wscale_kernel<<<8, 256>>>(net->d_blob, d_koff, d_cin, d_wscale, d_inv);
    tc::wpack_kernel<<<1024, 256>>>(net->d_blob, d_koff, d_wscale, (__half *)net->d_wpack);
    CK_CUDA(cudaGetLastError());
    return CK_OK;
}

template <int P, int S>
static int launch_tower(ck_net *net, const tc::TowerParams &prm, int64_t max_n, cudaStream_t stream) {
    static bool attr_done[64] = {false};          // per device (and per template instantiation)
    static_assert(tc::Cfg<P, S>::kSmem <= 232448, "tower tile does not fit in shared memory");
    if (!attr_done[net->device & 63]) {
        CK_CUDA(cudaFuncSetAttribute(tc::tower_tc_kernel<P, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::Cfg<P, S>::kSmem));
        attr_done[net->device & 63] = true;
    }
    const int64_t tiles = (max_n + P - 1) / P;
    const int grid = (int)std::min<int64_t>(tiles, num_sms(net->device));
    tc::tower_tc_kernel<P, S><<<grid, tc::kThreads, tc::Cfg<P, S>::kSmem, stream>>>(prm);
    CK_CUDA(cudaGetLastError());
    return CK_OK;
}

int net_tc_tower(ck_net *net, const ck_leaf *d_leaves, int64_t max_n, const int32_t *n_dev, float *d_trunk, float *d_pconv,
                 cudaStream_t stream, int *launches) {
    if (!net->d_wpack) return fail(CK_ERR_NO_NET, "tcgen05 tower: weights were never packed");
    const NetLayout L = net_layout();
    uint8_t *aux = (uint8_t *)net->d_wpack + tc::kPackBytes;
    tc::TowerParams prm;
    prm.leaves = d_leaves; prm.n_dev = n_dev; prm.max_n = (int32_t)max_n;
    prm.wpack = (const uint8_t *)net->d_wpack; prm.blob = net->d_blob; prm.fold = net->d_scale;
    prm.inv_scale = (const float *)aux + 8;
    prm.plane5 = (const float *)aux + 16;
    for (int i = 0; i < 8; ++i) prm.bias_off[i] = L.conv[i].bias;
    prm.trunk = d_trunk; prm.pconv = d_pconv;
    static const int variant = [] { const char *v = getenv("CK_TC_TILE"); return v ? atoi(v) : 3; }();
    int rc = variant == 2 ? launch_tower<2, 3>(net, prm, max_n, stream) : launch_tower<3, 2>(net, prm, max_n, stream);
    if (rc != CK_OK) return rc;
    if (launches) *launches += 1;
    return CK_OK;
}

}  // namespace ck
