// ck_net_ts.cu -- K3 tower, weights-in-TMEM kernel (tcgen05.mma with the A operand in tensor memory),
// sm_100a.  The product path of ck_net_forward* and of the engine; ck_net_tc.cu keeps the earlier
// all-shared-memory kernel as a cross-check (CK_TOWER=ss).
//
// Nine GEMM layers per position without leaving the SM: the eight 3x3 convolutions of create_nn (reference
// training_pipeline.py:57-88) and the policy head's 1x1 convolution (:89-96), each as
// D[co][square] = sum_k W[co][k] * act[k][square] with split-fp16 operands (Whi*Ahi + Whi*Alo + Wlo*Ahi,
// fp32 accumulation in TMEM) and the epilogue bias + ReLU + folded BatchNorm (conv -> bias -> ReLU -> BN,
// :60-63); the value head's 1x1 convolution (:102-105) is reduced in the conv6 epilogue.  Outputs per
// position: pflat[512] (policy features, flattened in (x, y, c) order) and vconv[64].
//
//   * WEIGHTS never touch shared memory.  Eight loader warps (two sets of four, one warp per TMEM lane
//     quadrant) stream the pre-packed split-fp16 weights from L2 with coalesced 16-byte loads (one output
//     channel = one TMEM lane per thread) and write them with tcgen05.st into a 16-slot ring of TMEM
//     columns [256,512) (one slot = one k-step = 16 input channels of one tap: 8 columns hi + 8 lo).  The
//     MMAs take A from TMEM, so shared-memory bandwidth is spent on the activation (B) operand only
//     (64 B/clk instead of 128 B/clk for an N = 128 tile with both operands in shared memory).
//   * TWO position tiles (X, Y; 2 positions = N 128 each, accumulators in TMEM columns [0,128) and
//     [128,256)) share every weight slot, each with its own MMA-issuing warp.  The issuers are not ordered
//     against each other: while one tile is in its epilogue the other owns the tensor pipe, and the ring
//     bounds how far they drift apart.  A single tile leaves the pipe idle during every epilogue (17 % of the
//     earlier kernel, profiles/r1c_tower_ncu_summary.json).  Small batches run with one tile per CTA
//     (TowerParams::tiles) to fill twice as many SMs.
//   * Activations: shared memory, split fp16 (hi, lo), zero-padded 10 x 10 boards in the UMMA K-major
//     no-swizzle core-matrix layout with rows interleaved over the 2 positions of a tile
//     (byte offset = chunk(ci/8)*kChunkStride + ((2*row + p)*10 + col)*16 + (ci%8)*2), so a 3x3 tap is a
//     descriptor start offset and one MMA covers the whole tile.  2 tiles x 100.5 KB.
//   * Epilogue: 8 warps (quadrant = warp % 4, two warps share a quadrant's columns) poll both accumulators;
//     tcgen05.ld.16x256b gives the mma-fragment layout, stmatrix.trans writes whole 16-byte operand units.
//   * Warp roles: warps 0-7 loaders, 8-15 epilogue, 16-17 MMA issuers (16 also allocates TMEM).
//     mbarriers: full/empty per weight slot, acc_full / act_ready per tile.
// Tried and measured neutral (kept out): replicating the weights per CTA group, longer epilogue poll sleeps,
// 16 epilogue warps, splitting act_ready so the next layer starts on the first half of the channels.
#include <cuda_fp16.h>
#include <stdlib.h>
#include <type_traits>
#include "ck_net.cuh"
#include "ck_tc_ptx.cuh"

namespace ck {
namespace ts {

using namespace ck::ptx;

constexpr int kP = 2;                               // positions per tile
constexpr int kN = 64 * kP;                         // MMA N
constexpr int kChunkStride = kP * 1600 + 16;        // 8 channels x (kP x 100 padded squares) x 2 B + 16 B bank spread
constexpr int kSplitBytes = 16 * kChunkStride;      // all 128 channels, hi or lo
constexpr int kTileBytes = 2 * kSplitBytes;
constexpr int kKSteps0 = 9, kKStepsL = 72;          // k-steps (16 input channels of one tap) per layer
constexpr int kKSteps8 = 8;                          // "layer 8": the policy head's 1x1 conv 128 -> 8, centre tap only
constexpr int kLayers = kTowerConvs + 1;            // 8 3x3 convolutions + the fused policy conv1x1
constexpr int kG = kKSteps0 + 7 * kKStepsL + kKSteps8;   // 521 k-steps = the whole weight stream (interleaved order)
// "Cross terms first" order (the product path).  The tensor pipe rounds its fp32 accumulator toward zero after every MMA,
// so a chain of 216 MMAs into one accumulator loses ~1e-5 relative in one direction.  Two thirds of
// those MMAs add the 2^-11-sized cross terms Whi*Alo + Wlo*Ahi: issued FIRST, while the accumulator is still tiny, their
// roundings cost nothing, and only the 72 Whi*Ahi MMAs round at full magnitude.  Price: a layer's Whi is streamed twice.
// A weight slot (16 TMEM columns) then holds either (Whi, Wlo) of one k-step [phase A: 2 cross MMAs] or Whi of two
// consecutive k-steps [phase B: 2 main MMAs], so every slot still feeds exactly two MMAs per tile.
__host__ __device__ constexpr int cf_slots(int ksteps) { return ksteps + (ksteps + 1) / 2; }
constexpr int kGcf = cf_slots(kKSteps0) + 7 * cf_slots(kKStepsL) + cf_slots(kKSteps8);   // 14 + 7 * 108 + 12 = 782 slot images
#ifndef CK_TS_RING_LOG2
#define CK_TS_RING_LOG2 4
#endif
constexpr int kNSLog2 = CK_TS_RING_LOG2;            // (3 = an 8-slot ring: the experiment that shows how much the ring's depth matters)
constexpr int kNS = 1 << kNSLog2;                   // TMEM weight ring slots (16 columns each)
constexpr int kWCol0 = 256;                         // first weight column; accumulators: X [0,128), Y [128,256)
// A layer's epilogue writes its output channels in two time halves (TMEM lanes 0-15 of every quadrant, then 16-31), i.e.
// first the input channels of the next layer's even k-steps, then those of its odd ones.  With CK_TS_HALF_SPLIT the next
// layer's cross-term MMAs of the even k-steps start after the first half (act_ready_a) and only the odd ones wait for the
// second (act_ready_b): half of an otherwise exposed epilogue hides behind 36 weight slots of tensor work.
#ifndef CK_TS_HALF_SPLIT
#define CK_TS_HALF_SPLIT 1
#endif
#ifndef CK_TS_FAKE_EPI
#define CK_TS_FAKE_EPI 0
#endif
#ifndef CK_TS_ISSUE2
#define CK_TS_ISSUE2 4
#endif
#ifndef CK_TS_ARRIVE_EACH
#define CK_TS_ARRIVE_EACH 0
#endif
#ifndef CK_TS_LOADER_SETS
#define CK_TS_LOADER_SETS 2
#endif
constexpr int kLoaderSets = CK_TS_LOADER_SETS;      // sets of four warps (one warp per TMEM lane quadrant) that take turns on the slot groups
constexpr int kLoaderWarps = 4 * kLoaderSets;
constexpr int kLoadGroup = 4;                       // k-steps a loader warp fetches per batch
constexpr float kActScale = 16.0f;                  // activations are stored as a * 2^4
#ifndef CK_POLL_NS
#define CK_POLL_NS 32
#endif
constexpr unsigned kPollNs = CK_POLL_NS;            // epilogue warps sleep this long between polls of the two accumulator barriers
constexpr size_t kWtsBytes = (size_t)kG * 8192;     // [k-step][unit 4][co 128][16 B]
constexpr size_t kWtsBytesCf = (size_t)kGcf * 8192; // [slot image][unit 4][co 128][16 B]
// All CTAs stream the same 4.3 MB (L2-resident after the first tile pair).  kCopies > 1 replicates the
// packed weights so that CTA b reads copy b % kCopies; measured no difference on B200 (the L2 serves the
// lock-step readers from one copy at the same rate), so one copy is kept.
constexpr int kCopies = 1;
constexpr uint32_t kIdesc = make_idesc_f16(128, kN);
constexpr int kBarOff = 2 * kTileBytes;
constexpr int kVredOff = kBarOff + 512;               // value conv1x1 partial sums [tile 2][quadrant 4][128 columns] fp32
constexpr int kSmem = kVredOff + 2 * 4 * kN * 4 + 16;       // + the two arrival counters

__host__ __device__ constexpr int sq_off(int p, int r, int c) { return ((kP * r + p) * 10 + c) * 16; }

struct TowerParams {
    const ck_leaf *leaves;
    const int32_t *n_dev;
    int32_t max_n;
    int32_t tiles;               // 2: tiles X and Y per CTA; 1: X only (small batches: half the latency, twice the CTAs)
    const uint4 *wts;            // split-fp16 weights, k-step order
    const float *blob;           // Keras-ordered fp32 parameters (biases)
    const float *fold;           // folded BN scale/shift table
    const float *inv_scale;      // per layer 1 / (weight scale * activation scale), 9 entries
    int64_t bias_off[9];         // conv0..6, policy conv3x3, policy conv1x1
    int64_t val1x1_k, val1x1_b;  // value head conv1x1 128 -> 1 (fp32, reduced in the conv6 epilogue)
    float *pflat;                // fp32 [n][512]: policy conv1x1 + ReLU + BN, flattened in (x, y, c) order
    float *vconv;                // fp32 [n][64]: value conv1x1 + ReLU + BN
    const float *plane5;         // 81-entry float32(n/80) table
    const uint4 *wts_cf;         // the same weights as slot images in cross-terms-first order
    int32_t *range_flag;         // set to 1 when an activation leaves the split-fp16 range (|a| * 2^4 >= 65504)
};

__device__ __forceinline__ uint4 ld_stream(const uint4 *p) {
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

template <int kEpiWarps, bool kCrossFirst>
__global__ void __launch_bounds__((kLoaderWarps + 2 + kEpiWarps) * 32, 1)
tower_ts_kernel(const TowerParams prm) {
    static_assert(kEpiWarps == 8, "epilogue: two warps per TMEM lane quadrant, 64 accumulator columns each");
    constexpr int kGs = kCrossFirst ? kGcf : kG;                    // weight slots one tile pair consumes
    constexpr int kMmaWarp0 = kLoaderWarps + kEpiWarps;            // warps kMmaWarp0 (tile X) and kMmaWarp0 + 1 (tile Y)
    extern __shared__ __align__(1024) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    int n = prm.max_n;
    if (prm.n_dev != nullptr) n = min(n, *prm.n_dev);
    const int ppi = prm.tiles * kP;                   // positions per CTA iteration
    const int n_pairs = (n + ppi - 1) / ppi;
    if ((int)blockIdx.x >= n_pairs) return;
    const uint32_t np = (uint32_t)(n_pairs - 1 - (int)blockIdx.x) / gridDim.x + 1;   // tile pairs of this CTA
    const uint32_t total = np * (uint32_t)kGs;                                       // weight slots this CTA streams

    const uint32_t bar0 = smem_u32(smem + kBarOff);
    auto bar_full = [&](uint32_t s) { return bar0 + 8u * s; };
    auto bar_empty = [&](uint32_t s) { return bar0 + 8u * (kNS + s); };
    auto bar_acc_full = [&](int t) { return bar0 + 8u * (2 * kNS + t); };
    auto bar_act_ready = [&](int t) { return bar0 + 8u * (2 * kNS + 2 + t); };        // first half of the channels (all of them for layer 0's input)
    auto bar_act_ready_b = [&](int t) { return bar0 + 8u * (2 * kNS + 4 + t); };      // second half
    uint32_t *s_tmem = reinterpret_cast<uint32_t *>(smem + kBarOff + 8 * (2 * kNS + 6));

    for (int i = tid * 16; i < 2 * kTileBytes; i += (int)blockDim.x * 16) *reinterpret_cast<uint4 *>(smem + i) = make_uint4(0, 0, 0, 0);
    if (warp == kMmaWarp0) {
        if (lane == 0) {
            for (int s = 0; s < kNS; ++s) { mbar_init(bar_full(s), 4); mbar_init(bar_empty(s), (uint32_t)prm.tiles); }
            for (int t = 0; t < 2; ++t) {
                mbar_init(bar_acc_full(t), 1); mbar_init(bar_act_ready(t), 32 * kEpiWarps); mbar_init(bar_act_ready_b(t), 32 * kEpiWarps);
            }
            mbar_init_fence();
            reinterpret_cast<int *>(smem + kVredOff + 2 * 4 * kN * 4)[0] = 0;
            reinterpret_cast<int *>(smem + kVredOff + 2 * 4 * kN * 4)[1] = 0;
        }
        __syncwarp();
        tmem_alloc<512>(smem_u32(s_tmem));
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    // the CTA owns all 512 columns, so the allocation starts at column 0, lane 0; the role code below
    // uses literal TMEM addresses (keeps the MMA operands provably warp-uniform)
    if (*s_tmem != 0u) __trap();

    if (warp < kLoaderWarps) {
        // ===== weight loaders: L2 -> registers -> TMEM ring (lane = output channel).  Two sets of four
        // warps (one warp per TMEM lane quadrant) alternate groups of kLoadGroup k-steps: a warp has
        // ONE batch of loads in flight at a time (ptxas puts every LDG of a warp on the same
        // scoreboard, so deeper per-warp prefetch does not overlap), and the other set's batch covers
        // its latency. =====
        const int set = warp >> 2, quad = warp & 3;
        const uint4 *src = (kCrossFirst ? prm.wts_cf : prm.wts + (size_t)(blockIdx.x % kCopies) * (kWtsBytes / 16)) + (quad * 32 + lane);
        const uint32_t t_w = ((uint32_t)(quad * 32) << 16) + kWCol0;
        uint32_t r[kLoadGroup][16];
        for (uint32_t base = (uint32_t)(set * kLoadGroup); base < total; base += kLoaderSets * kLoadGroup) {
#pragma unroll
            for (int j = 0; j < kLoadGroup; ++j) {
                const uint32_t g = (base + j) % (uint32_t)kGs;           // reads past `total` stay inside the stream
                const uint4 *p = src + (size_t)g * 512;
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const uint4 v = ld_stream(p + u * 128);
                    r[j][4 * u] = v.x; r[j][4 * u + 1] = v.y; r[j][4 * u + 2] = v.z; r[j][4 * u + 3] = v.w;
                }
            }
#if CK_TS_ARRIVE_EACH
            // every slot is announced as soon as its own store has landed: a stored slot that waits for the rest of its
            // batch is ring capacity the MMA issuers cannot use (up to 3 of the 16 slots)
#pragma unroll
            for (int j = 0; j < kLoadGroup; ++j) {
                const uint32_t ss = base + j;
                if (ss < total) {
                    const uint32_t slot = ss % kNS;
                    mbar_wait(bar_empty(slot), ((ss / kNS) & 1u) ^ 1u);
                    tc_fence_after();
                    tmem_st16(t_w + slot * 16, r[j]);
                    tmem_st_wait();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_full(slot));
                }
            }
#else
#pragma unroll
            for (int j = 0; j < kLoadGroup; ++j) {
                const uint32_t ss = base + j;
                if (ss < total) {
                    const uint32_t slot = ss % kNS;
                    mbar_wait(bar_empty(slot), ((ss / kNS) & 1u) ^ 1u);
                    tc_fence_after();
                    tmem_st16(t_w + slot * 16, r[j]);
                }
            }
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
#pragma unroll
                for (int j = 0; j < kLoadGroup; ++j)
                    if (base + j < total) mbar_arrive(bar_full((base + j) % kNS));
            }
#endif
        }
    } else if (warp >= kMmaWarp0) {
        // ===== MMA issuers, one warp per tile.  The whole warp runs the loop converged; elect.sync
        // inside the wrappers picks the issuing lane.  The two warps are not ordered against each
        // other: whichever tile is in its epilogue leaves the tensor pipe to the other one, and the
        // 16-slot weight ring bounds how far they drift apart. =====
        const int t = warp - kMmaWarp0;
        // trip count from kernel parameters and special registers only, so that the loop state (k-step
        // counter, descriptors, TMEM addresses) stays in uniform registers; the device-side batch count
        // only shortens the loop through the break below
        const uint32_t npu = t < prm.tiles ? (uint32_t)((prm.max_n + ppi - 1) / ppi - 1 - (int)blockIdx.x) / gridDim.x + 1 : 0u;
        const uint32_t tile16 = (smem_u32(smem) + (uint32_t)(t * kTileBytes)) >> 4;      // tile base in 16-byte units
        constexpr uint32_t kDescLo = (uint32_t)(kChunkStride >> 4) << 16;                   // LBO
        constexpr uint64_t kDescHi = ((uint64_t)((160u >> 4) | (1u << 14))) << 32;          // SBO, descriptor version
        const uint32_t d = (uint32_t)(t * kN);
        uint32_t s = 0, ar_phase = 0;
        for (uint32_t pair = 0; pair < npu; ++pair) {
            if (pair >= np) break;
            for (int layer = 0; layer < kLayers; ++layer) {
                constexpr bool kSplit = kCrossFirst && CK_TS_HALF_SPLIT && CK_TS_ISSUE2 == 4;
                mbar_wait(bar_act_ready(t), ar_phase);
                if (!kSplit || layer == 0) mbar_wait(bar_act_ready_b(t), ar_phase);     // (both halves flip together, one phase bit serves)
                ar_phase ^= 1u;
                tc_fence_after();
                const int nk = layer == 0 ? 1 : 8;
                const int tap_lo = layer == kLayers - 1 ? 4 : 0, tap_hi = layer == kLayers - 1 ? 5 : 9;   // conv1x1 = centre tap
                uint32_t acc = 0u;
                if (kCrossFirst) {
                    // phase A: the cross terms of every k-step of the layer (slot = Whi | Wlo of one k-step)
                    for (int tap = tap_lo; tap < tap_hi; ++tap) {
                        uint32_t b = tile16 + (uint32_t)(kP * (tap / 3) * 10 + tap % 3);
#if CK_TS_ISSUE2 == 4
                        if (nk == 8) {
                            if (kSplit) break;                // handled below: even k-steps of every tap, then the odd ones
                            // four slots per trip of the issuing loop
#pragma unroll 1
                            for (int kc = 0; kc < 8; kc += 4, s += 4) {
                                uint32_t aa[8], bars[4];
                                uint64_t bb[8];
#pragma unroll
                                for (int i = 0; i < 4; ++i) {
                                    const uint32_t sl = (s + i) & (kNS - 1);
                                    mbar_wait(bar_full(sl), ((s + i) >> kNSLog2) & 1u);
                                    const uint64_t bh = kDescHi | (uint64_t)(kDescLo + b + (uint32_t)((2 * i * kChunkStride) >> 4));
                                    aa[2 * i] = kWCol0 + sl * 16; bb[2 * i] = bh + (uint64_t)(kSplitBytes >> 4);
                                    aa[2 * i + 1] = kWCol0 + sl * 16 + 8; bb[2 * i + 1] = bh;
                                    bars[i] = bar_empty(sl);
                                }
                                tc_fence_after();
                                tc_oct_ts_elect(d, aa, bb, kIdesc, acc, bars);
                                acc = 1u;
                                b += (uint32_t)(8 * kChunkStride) >> 4;
                            }
                            continue;
                        }
#elif CK_TS_ISSUE2
                        if (nk == 8) {
                            // two slots per trip of the issuing loop (the per-slot overhead of waiting, fencing and electing
                            // is what the 1.5x more slots of this order cost)
#pragma unroll 1
                            for (int kc = 0; kc < 8; kc += 2, s += 2) {
                                const uint32_t s0 = s & (kNS - 1), s1 = (s + 1) & (kNS - 1);
                                mbar_wait(bar_full(s0), (s >> kNSLog2) & 1u);
                                mbar_wait(bar_full(s1), ((s + 1) >> kNSLog2) & 1u);
                                tc_fence_after();
                                const uint64_t bh0 = kDescHi | (uint64_t)(kDescLo + b), bh1 = bh0 + (uint64_t)((2 * kChunkStride) >> 4);
                                tc_quad_ts_elect(d, kWCol0 + s0 * 16, bh0 + (uint64_t)(kSplitBytes >> 4), kWCol0 + s0 * 16 + 8, bh0,
                                                 kWCol0 + s1 * 16, bh1 + (uint64_t)(kSplitBytes >> 4), kWCol0 + s1 * 16 + 8, bh1, kIdesc, acc,
                                                 bar_empty(s0), bar_empty(s1));
                                acc = 1u;
                                b += (uint32_t)(4 * kChunkStride) >> 4;
                            }
                            continue;
                        }
#endif
#pragma unroll 1
                        for (int kc = 0; kc < nk; ++kc, ++s) {
                            const uint32_t slot = s & (kNS - 1);
                            mbar_wait(bar_full(slot), (s >> kNSLog2) & 1u);
                            tc_fence_after();
                            const uint64_t bhi = kDescHi | (uint64_t)(kDescLo + b);
                            tc_pair_ts_elect(d, kWCol0 + slot * 16, bhi + (uint64_t)(kSplitBytes >> 4),      // Whi * Alo
                                             kWCol0 + slot * 16 + 8, bhi, kIdesc, acc, bar_empty(slot));     // Wlo * Ahi
                            acc = 1u;
                            b += (uint32_t)(2 * kChunkStride) >> 4;
                        }
                    }
#if CK_TS_ISSUE2 == 4
                    if (kSplit && nk == 8) {
#pragma unroll 1
                        for (int half = 0; half < 2; ++half) {
                            if (half == 1) { mbar_wait(bar_act_ready_b(t), ar_phase ^ 1u); tc_fence_after(); }   // the phase bit was flipped above
#pragma unroll 1
                            for (int tap = tap_lo; tap < tap_hi; ++tap, s += 4) {
                                // k-steps half, half + 2, half + 4, half + 6 of this tap: 16 input channels each, two 8-channel chunks apart
                                const uint32_t b = tile16 + (uint32_t)(kP * (tap / 3) * 10 + tap % 3) + (uint32_t)((2 * half * kChunkStride) >> 4);
                                uint32_t aa[8], bars[4];
                                uint64_t bb[8];
#pragma unroll
                                for (int i = 0; i < 4; ++i) {
                                    const uint32_t sl = (s + i) & (kNS - 1);
                                    mbar_wait(bar_full(sl), ((s + i) >> kNSLog2) & 1u);
                                    const uint64_t bh = kDescHi | (uint64_t)(kDescLo + b + (uint32_t)((4 * i * kChunkStride) >> 4));
                                    aa[2 * i] = kWCol0 + sl * 16; bb[2 * i] = bh + (uint64_t)(kSplitBytes >> 4);
                                    aa[2 * i + 1] = kWCol0 + sl * 16 + 8; bb[2 * i + 1] = bh;
                                    bars[i] = bar_empty(sl);
                                }
                                tc_fence_after();
                                tc_oct_ts_elect(d, aa, bb, kIdesc, acc, bars);
                                acc = 1u;
                            }
                        }
                    }
#endif
                    // phase B: Whi * Ahi, two consecutive k-steps per slot
                    if (layer == 0) {
                        // nine k-steps (one per tap, 16 input channels): pairs of taps, the tenth half-slot holds zeros
#pragma unroll 1
                        for (int j = 0; j < 5; ++j, ++s) {
                            const uint32_t slot = s & (kNS - 1);
                            mbar_wait(bar_full(slot), (s >> kNSLog2) & 1u);
                            tc_fence_after();
                            const int t0 = 2 * j, t1 = 2 * j + 1 < 9 ? 2 * j + 1 : 2 * j;
                            const uint32_t b0 = tile16 + (uint32_t)(kP * (t0 / 3) * 10 + t0 % 3), b1 = tile16 + (uint32_t)(kP * (t1 / 3) * 10 + t1 % 3);
                            tc_pair_ts_elect(d, kWCol0 + slot * 16, kDescHi | (uint64_t)(kDescLo + b0),
                                             kWCol0 + slot * 16 + 8, kDescHi | (uint64_t)(kDescLo + b1), kIdesc, 1u, bar_empty(slot));
                        }
                    } else {
                        for (int tap = tap_lo; tap < tap_hi; ++tap) {
                            uint32_t b = tile16 + (uint32_t)(kP * (tap / 3) * 10 + tap % 3);
#if CK_TS_ISSUE2 == 4
                            {
                                uint32_t aa[8], bars[4];
                                uint64_t bb[8];
                                const uint64_t st = (uint64_t)((2 * kChunkStride) >> 4);
#pragma unroll
                                for (int i = 0; i < 4; ++i) {
                                    const uint32_t sl = (s + i) & (kNS - 1);
                                    mbar_wait(bar_full(sl), ((s + i) >> kNSLog2) & 1u);
                                    const uint64_t b0 = kDescHi | (uint64_t)(kDescLo + b);
                                    aa[2 * i] = kWCol0 + sl * 16; bb[2 * i] = b0 + (uint64_t)(2 * i) * st;
                                    aa[2 * i + 1] = kWCol0 + sl * 16 + 8; bb[2 * i + 1] = b0 + (uint64_t)(2 * i + 1) * st;
                                    bars[i] = bar_empty(sl);
                                }
                                tc_fence_after();
                                tc_oct_ts_elect(d, aa, bb, kIdesc, 1u, bars);
                                s += 4;
                            }
#elif CK_TS_ISSUE2
#pragma unroll 1
                            for (int kc = 0; kc < 8; kc += 4, s += 2) {
                                const uint32_t s0 = s & (kNS - 1), s1 = (s + 1) & (kNS - 1);
                                mbar_wait(bar_full(s0), (s >> kNSLog2) & 1u);
                                mbar_wait(bar_full(s1), ((s + 1) >> kNSLog2) & 1u);
                                tc_fence_after();
                                const uint64_t b0 = kDescHi | (uint64_t)(kDescLo + b), st = (uint64_t)((2 * kChunkStride) >> 4);
                                tc_quad_ts_elect(d, kWCol0 + s0 * 16, b0, kWCol0 + s0 * 16 + 8, b0 + st, kWCol0 + s1 * 16, b0 + 2 * st,
                                                 kWCol0 + s1 * 16 + 8, b0 + 3 * st, kIdesc, 1u, bar_empty(s0), bar_empty(s1));
                                b += (uint32_t)(8 * kChunkStride) >> 4;
                            }
#else
#pragma unroll 1
                            for (int kc = 0; kc < 8; kc += 2, ++s) {
                                const uint32_t slot = s & (kNS - 1);
                                mbar_wait(bar_full(slot), (s >> kNSLog2) & 1u);
                                tc_fence_after();
                                const uint64_t b0 = kDescHi | (uint64_t)(kDescLo + b);
                                tc_pair_ts_elect(d, kWCol0 + slot * 16, b0, kWCol0 + slot * 16 + 8, b0 + (uint64_t)((2 * kChunkStride) >> 4),
                                                 kIdesc, 1u, bar_empty(slot));
                                b += (uint32_t)(4 * kChunkStride) >> 4;
                            }
#endif
                        }
                    }
                } else
                for (int tap = tap_lo; tap < tap_hi; ++tap) {
                    uint32_t b = tile16 + (uint32_t)(kP * (tap / 3) * 10 + tap % 3);     // sq_off(0, tap / 3, tap % 3) / 16
#pragma unroll 1
                    for (int kc = 0; kc < nk; ++kc, ++s) {
                        const uint32_t slot = s & (kNS - 1);
                        mbar_wait(bar_full(slot), (s >> kNSLog2) & 1u);
                        tc_fence_after();
                        tc_kstep_ts_elect(d, kWCol0 + slot * 16, kDescHi | (uint64_t)(kDescLo + b), (uint32_t)(kSplitBytes >> 4), kIdesc, acc,
                                          bar_empty(slot));            // 2 arrivals (X and Y) free the slot
                        acc = 1u;
                        b += (uint32_t)(2 * kChunkStride) >> 4;
                    }
                }
                tc_commit_elect(bar_acc_full(t));
            }
        }
    } else {
        // ===== epilogue warps: TMEM lane quadrant = warp % 4, one output channel per thread.  Every
        // warp serves both tiles and takes whichever accumulator is complete first. =====
        const int ew = warp - kLoaderWarps, et = tid - kLoaderWarps * 32;
        const int quad = warp & 3, sub = ew >> 2;

        auto build_input = [&](int t, uint32_t k) {
            // planes 0..13 (+2 zero channels) of the tile's positions, thread = (square, chunk, position)
            if (et < 64 * 2 * kP) {
                const int sq = et & 63, ch = (et >> 6) & 1, p = et >> 7, x = sq >> 3, y = sq & 7;
                const bool dark = ((x ^ y) & 1) != 0;
                const uint32_t bit = 1u << (4 * x + (y >> 1));
                const int64_t pos = ((int64_t)blockIdx.x + (int64_t)k * gridDim.x) * ppi + t * kP + p;
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
                uint8_t *dst = smem + t * kTileBytes + ch * kChunkStride + sq_off(p, x + 1, y + 1);
                *reinterpret_cast<uint4 *>(dst) = *reinterpret_cast<const uint4 *>(hi);
                *reinterpret_cast<uint4 *>(dst + kSplitBytes) = *reinterpret_cast<const uint4 *>(lo);
            }
            fence_proxy_async();
            tc_fence_before();
            mbar_arrive(bar_act_ready(t));
            mbar_arrive(bar_act_ready_b(t));
        };

        build_input(0, 0);
        if (prm.tiles > 1) build_input(1, 0);
        // per tile: next event = (pair ek, layer el), acc_full phase ph   (scalars: a runtime tile index
        // into arrays would put them in local memory)
        uint32_t ek0 = 0u, ek1 = prm.tiles > 1 ? 0u : np, ph0 = 0u, ph1 = 0u;      // one tile: Y has no events
        int el0 = 0, el1 = 0;
        while (ek0 < np || ek1 < np) {
            int t = -1;
            if (ek0 < np && __any_sync(0xFFFFFFFFu, mbar_test(bar_acc_full(0), ph0))) t = 0;
            else if (ek1 < np && __any_sync(0xFFFFFFFFu, mbar_test(bar_acc_full(1), ph1))) t = 1;
            if (t < 0) { __nanosleep(kPollNs); continue; }
            const int layer = t ? el1 : el0;
            const uint32_t k = t ? ek1 : ek0;
            if (t) ph1 ^= 1u; else ph0 ^= 1u;
            tc_fence_after();
            // Accumulator column = (2*x + p)*8 + y: one 8-column group g = 2*x + p is board row x of position p.
            // Each warp drains 64 columns (8 groups) of its 32 lanes in two 16-lane halves with the
            // fragment-layout load (thread: lanes t/4 and t/4+8, two adjacent squares), so that
            // stmatrix.trans can write whole 16-byte units (8 channels of one square) of the next
            // layer's operand: 8x fewer shared-memory store instructions than 2-byte stores.
            const int64_t pos0 = ((int64_t)blockIdx.x + (int64_t)k * gridDim.x) * ppi + t * kP;
            const float inv = prm.inv_scale[layer];
            const int t4 = lane >> 2, tq = lane & 3;
            uint32_t cur[32];
            if (layer < kTowerConvs) {
                const float *fold = prm.fold + kScaleTower + layer * 2 * kC;
                const float *bias_p = prm.blob + prm.bias_off[layer];
                const uint32_t st_base = smem_u32(smem) + (uint32_t)(t * kTileBytes) + ((lane >> 4) ? (uint32_t)kSplitBytes : 0u) +
                                         (uint32_t)((4 * quad + ((lane >> 3) & 1)) * kChunkStride + sq_off(0, 1, (lane & 7) + 1));
                __half2 amax = __float2half2_rn(0.f);           // range guard: largest |hi half| this thread wrote
                float vp[16];                                   // value conv1x1 partial sums (conv6 epilogue only)
#pragma unroll
                for (int i = 0; i < 16; ++i) vp[i] = 0.f;
                // one code copy per "does this layer's output also feed the value head" (conv6): the test is out of the inner loop
                auto epilogue_halves = [&](auto val_tag) {
                constexpr bool kVal = decltype(val_tag)::value;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    tmem_ld_16x256b_x8_async(((uint32_t)(quad * 32 + 16 * h) << 16) + (uint32_t)(t * kN + 64 * sub), cur);
                    const int ch0 = quad * 32 + 16 * h + t4;                 // and ch0 + 8
                    const float bias0 = bias_p[ch0], bias1 = bias_p[ch0 + 8];
                    const float sc0 = fold[ch0] * kActScale, sc1 = fold[ch0 + 8] * kActScale;
                    const float sh0 = fold[kC + ch0] * kActScale, sh1 = fold[kC + ch0 + 8] * kActScale;
                    float wv0 = 0.f, wv1 = 0.f;
                    if (kVal) { wv0 = prm.blob[prm.val1x1_k + ch0]; wv1 = prm.blob[prm.val1x1_k + ch0 + 8]; }
                    tmem_ld_wait32(cur);
                    if (h == 1) {
                        // the accumulator is read out completely and the first half of the channels (TMEM lanes 0-15 of every
                        // quadrant = the next layer's even k-steps) is in shared memory: the next layer may start on those
                        fence_proxy_async();
                        tc_fence_before();
                        mbar_arrive(bar_act_ready(t));
                    }
#if CK_TS_FAKE_EPI == 1
                    if (prm.max_n > 0) continue;          // TIMING EXPERIMENT ONLY: how much of the launch is exposed epilogue work
#endif
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int g = 8 * sub + j;
                        const float a0 = fmaf(fmaxf(fmaf(__uint_as_float(cur[4 * j + 0]), inv, bias0), 0.f), sc0, sh0);
                        const float a1 = fmaf(fmaxf(fmaf(__uint_as_float(cur[4 * j + 1]), inv, bias0), 0.f), sc0, sh0);
                        const float a2 = fmaf(fmaxf(fmaf(__uint_as_float(cur[4 * j + 2]), inv, bias1), 0.f), sc1, sh1);
                        const float a3 = fmaf(fmaxf(fmaf(__uint_as_float(cur[4 * j + 3]), inv, bias1), 0.f), sc1, sh1);
                        if (kVal) {
                            vp[2 * j] = fmaf(wv1, a2, fmaf(wv0, a0, vp[2 * j]));
                            vp[2 * j + 1] = fmaf(wv1, a3, fmaf(wv0, a1, vp[2 * j + 1]));
                        }
                        const __half2 h0 = __floats2half2_rn(a0, a1), h1 = __floats2half2_rn(a2, a3);
                        const float2 f0 = __half22float2(h0), f1 = __half22float2(h1);
                        const __half2 l0 = __floats2half2_rn(a0 - f0.x, a1 - f0.y), l1 = __floats2half2_rn(a2 - f1.x, a3 - f1.y);
                        amax = __hmax2_nan(amax, __hmax2_nan(__habs2(h0), __habs2(h1)));
                        stmatrix_x4_trans(st_base + (uint32_t)(2 * h * kChunkStride + g * 160),
                                          *reinterpret_cast<const uint32_t *>(&h0), *reinterpret_cast<const uint32_t *>(&h1),
                                          *reinterpret_cast<const uint32_t *>(&l0), *reinterpret_cast<const uint32_t *>(&l1));
                    }
                }
                };
                if (layer == 6) epilogue_halves(std::true_type{}); else epilogue_halves(std::false_type{});
                if (layer == 6) {
                    // value head conv1x1 128 -> 1 (training_pipeline.py:102-105) on the conv6 output while it is in
                    // registers: reduce over the 8 lanes that hold different channels, park the per-quadrant
                    // partial sums in shared memory; the LAST of the 8 warps to get here adds the four
                    // quadrants in a fixed order and writes ReLU/BN'd results (deterministic, no barrier).
                    float *vred = reinterpret_cast<float *>(smem + kVredOff) + (t * 4 + quad) * kN + 64 * sub;
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        float v = vp[i];
                        v += __shfl_xor_sync(0xFFFFFFFFu, v, 4);
                        v += __shfl_xor_sync(0xFFFFFFFFu, v, 8);
                        v += __shfl_xor_sync(0xFFFFFFFFu, v, 16);
                        if (t4 == 0) vred[8 * (i >> 1) + 2 * tq + (i & 1)] = v;
                    }
                    __threadfence_block();
                    __syncwarp();
                    int *vcnt = reinterpret_cast<int *>(smem + kVredOff + 2 * 4 * kN * 4) + t;
                    int old = 0;
                    if (lane == 0) old = atomicAdd(vcnt, 1);
                    old = __shfl_sync(0xFFFFFFFFu, old, 0);
                    if (old == kEpiWarps - 1) {
                        if (lane == 0) *vcnt = 0;
                        __threadfence_block();
                        const float *vr = reinterpret_cast<const float *>(smem + kVredOff) + t * 4 * kN;
                        const float bv = prm.blob[prm.val1x1_b], scv = prm.fold[kScaleVal1x1], shv = prm.fold[kScaleVal1x1 + 1];
                        constexpr float q = 1.0f / kActScale;
                        for (int col = lane; col < kN; col += 32) {
                            const float sum = ((vr[col] + vr[kN + col]) + vr[2 * kN + col]) + vr[3 * kN + col];
                            const int g = col >> 3, x = g >> 1, p = g & 1, y = col & 7;
                            if (pos0 + p < n) prm.vconv[(pos0 + p) * 64 + x * 8 + y] = fmaf(fmaxf(fmaf(sum, q, bv), 0.f), scv, shv);
                        }
                    }
                }
                // an activation beyond fp16's range became +-inf above (and its lo half NaN): never silent
                {
                    const uint32_t ab = *reinterpret_cast<const uint32_t *>(&amax);      // exponent all ones: inf or NaN
                    if ((ab & 0x7C00u) == 0x7C00u || (ab & 0x7C000000u) == 0x7C000000u) atomicExch(prm.range_flag, 1);
                }
                fence_proxy_async();
                tc_fence_before();
                mbar_arrive(bar_act_ready_b(t));
                if (t) el1 = layer + 1; else el0 = layer + 1;
            } else {
                // fused policy conv1x1 128 -> 8 (+ bias, ReLU, BN; training_pipeline.py:89-96): accumulator rows 0..7,
                // i.e. the first 16-lane half of quadrant 0; flattened in (x, y, c) order for the Dense(512)
                if (quad == 0) {
                    tmem_ld_16x256b_x8_async((uint32_t)(t * kN + 64 * sub), cur);
                    const int o = t4;
                    const float bias0 = prm.blob[prm.bias_off[8] + o];
                    const float sc0 = prm.fold[kScalePol1x1 + o], sh0 = prm.fold[kScalePol1x1 + 8 + o];
                    tmem_ld_wait32(cur);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int g = 8 * sub + j, x = g >> 1, p = g & 1;
                        if (pos0 + p < n) {
                            float *dst = prm.pflat + (pos0 + p) * 512 + (x * 8 + 2 * tq) * 8 + o;
                            dst[0] = fmaf(fmaxf(fmaf(__uint_as_float(cur[4 * j + 0]), inv, bias0), 0.f), sc0, sh0);
                            dst[8] = fmaf(fmaxf(fmaf(__uint_as_float(cur[4 * j + 1]), inv, bias0), 0.f), sc0, sh0);
                        }
                    }
                }
                tc_fence_before();
                if (t) { el1 = 0; ek1 = k + 1; } else { el0 = 0; ek0 = k + 1; }
                if (k + 1 < np) build_input(t, k + 1);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == kMmaWarp0) tmem_dealloc<512>(0u);
}

// ---- weight packing (once per ck_net_set_weights) ---------------------------------------------
struct PackAux { int64_t koff[9]; int cin[9]; };     // conv0..6, policy conv3x3, policy conv1x1 (128 -> 8)

// per layer: power-of-two scale S with max|w| * S in [8192, 16384)
__global__ void wscale_kernel(const float *__restrict__ blob, PackAux aux, float *__restrict__ wscale, float *__restrict__ inv_scale) {
    const int layer = blockIdx.x;
    const int64_t n = layer < kTowerConvs ? (int64_t)9 * aux.cin[layer] * kC : (int64_t)kC * 8;
    const float *w = blob + aux.koff[layer];
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

// one thread per 16-byte unit: [k-step][unit: hi ci 0-7, hi ci 8-15, lo ci 0-7, lo ci 8-15][co]
__global__ void wpack_kernel(const float *__restrict__ blob, PackAux aux, const float *__restrict__ wscale, uint4 *__restrict__ out) {
    const int64_t total = (int64_t)kG * 512;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int co = (int)(i & 127), u = (int)((i >> 7) & 3), g = (int)(i >> 9);
        int layer, tap, ci0;
        if (g < kKSteps0) { layer = 0; tap = g; ci0 = 0; }
        else if (g < kKSteps0 + 7 * kKStepsL) { const int j = g - kKSteps0; layer = 1 + j / kKStepsL; const int k = j % kKStepsL; tap = k >> 3; ci0 = (k & 7) * 16; }
        else { layer = kTowerConvs; tap = 0; ci0 = (g - kKSteps0 - 7 * kKStepsL) * 16; }
        const int cin = aux.cin[layer];
        const int cout = layer < kTowerConvs ? kC : 8;            // the conv1x1 fills accumulator rows 0..7 only
        const float S = wscale[layer];
        __half h[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int ci = ci0 + (u & 1) * 8 + e;
            float w = 0.f;
            if (ci < cin && co < cout) w = blob[aux.koff[layer] + ((int64_t)tap * cin + ci) * cout + co] * S;
            const __half hi = __float2half_rn(w);
            h[e] = (u & 2) ? __float2half_rn(w - __half2float(hi)) : hi;
        }
        out[i] = *reinterpret_cast<const uint4 *>(h);
    }
}

// cross-terms-first stream: slot images [unit 0,1: first operand half (columns 0-7), unit 2,3: second half][co].
// Phase-A image of k-step k: (Whi_k | Wlo_k); phase-B image j: (Whi_2j | Whi_2j+1), zeros beyond the layer's last k-step.
__global__ void wpack_cf_kernel(const float *__restrict__ blob, PackAux aux, const float *__restrict__ wscale, uint4 *__restrict__ out) {
    const int64_t total = (int64_t)kGcf * 512;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int co = (int)(i & 127), u = (int)((i >> 7) & 3);
        int g = (int)(i >> 9);
        int layer = 0, ksteps = kKSteps0;
        while (g >= cf_slots(ksteps)) { g -= cf_slots(ksteps); ++layer; ksteps = layer < kTowerConvs ? kKStepsL : kKSteps8; }
        int kstep;
        bool lo;
        if (g < ksteps) {                                                       // phase A: hi | lo of one k-step
            kstep = g;
#if CK_TS_HALF_SPLIT && CK_TS_ISSUE2 == 4
            // the even k-steps (input channels 0-15, 32-47, ...) of every tap first, then the odd ones: the order in which the
            // previous layer's epilogue delivers the channels
            if (ksteps == kKStepsL) { const int half = g / 36, r = g % 36; kstep = (r >> 2) * 8 + 2 * (r & 3) + half; }
            else if (ksteps == kKSteps8 && layer == kTowerConvs) { kstep = 2 * (g & 3) + (g >> 2); }
#endif
            lo = (u & 2) != 0;
        }
        else { kstep = 2 * (g - ksteps) + ((u & 2) ? 1 : 0); lo = false; }      // phase B: hi of two k-steps
        int tap, ci0;
        if (layer == 0) { tap = kstep; ci0 = 0; }
        else if (layer < kTowerConvs) { tap = kstep >> 3; ci0 = (kstep & 7) * 16; }
        else { tap = 0; ci0 = kstep * 16; }
        const int cin = aux.cin[layer];
        const int cout = layer < kTowerConvs ? kC : 8;
        const float S = wscale[layer];
        __half h[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int ci = ci0 + (u & 1) * 8 + e;
            float w = 0.f;
            if (kstep < ksteps && ci < cin && co < cout) w = blob[aux.koff[layer] + ((int64_t)tap * cin + ci) * cout + co] * S;
            const __half hi = __float2half_rn(w);
            h[e] = lo ? __float2half_rn(w - __half2float(hi)) : hi;
        }
        out[i] = *reinterpret_cast<const uint4 *>(h);
    }
}

}  // namespace ts

int net_ts_prepare(ck_net *net) {
    const NetLayout L = net_layout();
    if (!net->d_wts) {
        // kCopies x [packed weights], then [wscale 16 f][inv 16 f][plane5 81 f], then the cross-terms-first stream
        CK_CUDA(cudaMalloc(&net->d_wts, ts::kCopies * ts::kWtsBytes + 1024 + ts::kWtsBytesCf));
    }
    float *aux = (float *)((uint8_t *)net->d_wts + ts::kCopies * ts::kWtsBytes);
    ts::PackAux h;
    for (int i = 0; i < 8; ++i) { h.koff[i] = L.conv[i].kernel; h.cin[i] = L.conv[i].cin; }
    h.koff[8] = L.pol1x1.kernel; h.cin[8] = kC;
    float p5[81];
    for (int i = 0; i <= 80; ++i) p5[i] = (float)((double)i / 80.0);
    CK_CUDA(cudaMemcpy(aux + 32, p5, sizeof(p5), cudaMemcpyHostToDevice));
    ts::wscale_kernel<<<ts::kLayers, 256>>>(net->d_blob, h, aux, aux + 16);
    ts::wpack_kernel<<<1024, 256>>>(net->d_blob, h, aux, (uint4 *)net->d_wts);
    ts::wpack_cf_kernel<<<1024, 256>>>(net->d_blob, h, aux, (uint4 *)((uint8_t *)net->d_wts + ts::kCopies * ts::kWtsBytes + 1024));
    CK_CUDA(cudaGetLastError());
    for (int c = 1; c < ts::kCopies; ++c)
        CK_CUDA(cudaMemcpyAsync((uint8_t *)net->d_wts + c * ts::kWtsBytes, net->d_wts, ts::kWtsBytes, cudaMemcpyDeviceToDevice, 0));
    return CK_OK;
}

template <int kEpiWarps, bool kCrossFirst>
static int launch_tower_ts(ck_net *net, const ts::TowerParams &prm, int64_t max_n, cudaStream_t stream) {
    static bool attr_done[64] = {false};          // per device (and per template instantiation)
    static_assert(ts::kSmem <= 232448, "tower tiles do not fit in shared memory");
    if (!attr_done[net->device & 63]) {
        CK_CUDA(cudaFuncSetAttribute(ts::tower_ts_kernel<kEpiWarps, kCrossFirst>, cudaFuncAttributeMaxDynamicSharedMemorySize, ts::kSmem));
        attr_done[net->device & 63] = true;
    }
    const int64_t pairs = (max_n + prm.tiles * ts::kP - 1) / (prm.tiles * ts::kP);
    const int grid = (int)std::min<int64_t>(pairs, num_sms(net->device));
    ts::tower_ts_kernel<kEpiWarps, kCrossFirst><<<grid, (ts::kLoaderWarps + 2 + kEpiWarps) * 32, ts::kSmem, stream>>>(prm);
    CK_CUDA(cudaGetLastError());
    return CK_OK;
}

int net_ts_tower(ck_net *net, const ck_leaf *d_leaves, int64_t max_n, const int32_t *n_dev, float *d_pflat, float *d_vconv,
                 cudaStream_t stream, int *launches) {
    if (!net->d_wts) return fail(CK_ERR_NO_NET, "tcgen05 tower: weights were never packed");
    const NetLayout L = net_layout();
    const float *aux = (const float *)((const uint8_t *)net->d_wts + ts::kCopies * ts::kWtsBytes);
    ts::TowerParams prm;
    prm.leaves = d_leaves; prm.n_dev = n_dev; prm.max_n = (int32_t)max_n;
    // small batches (arena, tournaments, a few hundred self-play games): one tile per CTA fills twice as many
    // SMs and halves the latency of the launch; from two positions per SM on, two tiles share the weight stream
    static const int force_tiles = [] { const char *v = getenv("CK_TS_TILES"); return v ? atoi(v) : 0; }();
    prm.tiles = force_tiles == 1 || force_tiles == 2 ? force_tiles : (max_n <= (int64_t)ts::kP * num_sms(net->device) ? 1 : 2);
    prm.wts = (const uint4 *)net->d_wts; prm.blob = net->d_blob; prm.fold = net->d_scale;
    prm.wts_cf = (const uint4 *)((const uint8_t *)net->d_wts + ts::kCopies * ts::kWtsBytes + 1024);
    prm.inv_scale = aux + 16;
    prm.plane5 = aux + 32;
    for (int i = 0; i < 8; ++i) prm.bias_off[i] = L.conv[i].bias;
    prm.bias_off[8] = L.pol1x1.bias;
    prm.val1x1_k = L.val1x1.kernel; prm.val1x1_b = L.val1x1.bias;
    prm.pflat = d_pflat; prm.vconv = d_vconv;
    prm.range_flag = net->d_range_flag;
    // CK_TS_ORDER=il selects the interleaved MMA order of round 1 (one chain of 216 MMAs per layer; A/B of the accuracy fix)
    static const bool interleaved = [] { const char *v = getenv("CK_TS_ORDER"); return v && v[0] == 'i'; }();
    const int rc = interleaved ? launch_tower_ts<8, false>(net, prm, max_n, stream) : launch_tower_ts<8, true>(net, prm, max_n, stream);
    if (rc != CK_OK) return rc;
    if (launches) *launches += 1;
    return CK_OK;
}

}  // namespace ck
