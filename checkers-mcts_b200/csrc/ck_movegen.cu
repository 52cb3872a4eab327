// ck_movegen.cu -- K1 legal-successor sweep, K4 random playouts, prior mask/renormalise,
// and the library's error plumbing.  sm_100a only.
#include <vector>
#include "ck_common.cuh"
#include "ck_rules.cuh"
#include "ck_device_fn.cuh"

namespace ck {

static thread_local std::string g_err;
void set_error(const std::string &msg) { g_err = msg; }
int fail(int code, const std::string &msg) { g_err = msg; return code; }

int num_sms(int device) {
    int n = 148;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device);
    return n;
}

// ---- K1 ------------------------------------------------------------------------------
// One thread per position: 16 B coalesced load, successors in reference order, 32 B mask.
// HBM-bound streaming kernel (16 + 4 + 32 + 16*b bytes per position).
__global__ void __launch_bounds__(256)
movegen_kernel(const uint4 *__restrict__ pos, int64_t n, int max_children, ck_pos *__restrict__ children,
               int32_t *__restrict__ counts, uint4 *__restrict__ masks, uint8_t *__restrict__ status,
               uint8_t *__restrict__ plane5) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint4 v = __ldg(pos + i);
        ck_pos p;
        p.p1 = v.x; p.p2 = v.y; p.k = v.z; p.meta = v.w;
        uint32_t mask[8];
        int cnt;
        if (children != nullptr)
            cnt = gen_moves(p, ArraySink{children + i * max_children, max_children}, mask);
        else
            cnt = gen_moves(p, NullSink{}, mask);
        int p5;
        const int st = outcome_of(p, cnt > 0, &p5);
        if (counts) counts[i] = cnt;
        if (masks) {
            masks[2 * i] = make_uint4(mask[0], mask[1], mask[2], mask[3]);
            masks[2 * i + 1] = make_uint4(mask[4], mask[5], mask[6], mask[7]);
        }
        if (status) status[i] = (uint8_t)st;
        if (plane5) plane5[i] = (uint8_t)p5;
    }
}

// ---- K1, packed output -------------------------------------------------------------------
// Same successors as movegen_kernel, written back to back (CSR): the children of position i are
// children[offsets[i] .. offsets[i+1]) in the reference's list order.  The strided [n][max_children]
// layout touches 768 B of address space per position to store ~70 B; here the kernels move the
// algorithmic bytes only (16 B in, 4 B offset, 32 B mask, 2 B status/plane5 and 16 B per child out).
// Two launches on the caller's stream:
//   count  thread per position: legal-action planes, status, plane5, and the tile's (256 positions)
//          successor count; the last tile to finish scans the tile counts into tile bases (offsets[n] = total)
//   emit   tile: positions again (L2 hits), block scan -> offsets, a (square, direction, owner) byte list of
//          the tile's moves in the reference's order; then ONE LANE PER SUCCESSOR builds it with
//          make_child_fast.  Every lane does the same work and consecutive lanes store consecutive 16-byte
//          successors (one lane per position ran its piece x direction loops around the full successor
//          construction with 11.6 of 32 lanes active and scattered its stores).
// (A single-pass variant that chained the tiles with a decoupled look-back was latency-bound: with
// 256-position tiles ~1200 tiles are in flight and each walked ~37 predecessor probes.)
constexpr int kCsrTile = 256;                        // positions per tile; owner indices are stored in one byte
static_assert(kCsrTile <= 256, "s_owner holds the position index of a move in a uint8_t");

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t *s_warp_tot, uint32_t *total) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t u = __shfl_up_sync(0xFFFFFFFFu, incl, o);
        if (lane >= o) incl += u;
    }
    if (lane == 31) s_warp_tot[w] = incl;
    __syncthreads();
    uint32_t warp_off = 0, tot = 0;
#pragma unroll
    for (int q = 0; q < kCsrTile / 32; ++q) {
        const uint32_t u = s_warp_tot[q];
        if (q < w) warp_off += u;
        tot += u;
    }
    *total = tot;
    return warp_off + incl - v;
}

__global__ void __launch_bounds__(kCsrTile)
movegen_count_kernel(const uint4 *__restrict__ pos, int64_t n, uint4 *__restrict__ masks, uint8_t *__restrict__ status,
                     uint8_t *__restrict__ plane5, uint32_t *__restrict__ tile_tot, unsigned int *__restrict__ done,
                     uint32_t *__restrict__ total_out) {
    __shared__ uint32_t s_warp_tot[kCsrTile / 32];
    const int64_t i = (int64_t)blockIdx.x * kCsrTile + threadIdx.x;
    int cnt = 0;
    if (i < n) {
        const uint4 v = __ldg(pos + i);
        ck_pos p;
        p.p1 = v.x; p.p2 = v.y; p.k = v.z; p.meta = v.w;
        uint32_t mask[8];
        cnt = gen_moves(p, NullSink{}, mask);
        int p5;
        const int st = outcome_of(p, cnt > 0, &p5);
        if (masks) {
            masks[2 * i] = make_uint4(mask[0], mask[1], mask[2], mask[3]);
            masks[2 * i + 1] = make_uint4(mask[4], mask[5], mask[6], mask[7]);
        }
        if (status) status[i] = (uint8_t)st;
        if (plane5) plane5[i] = (uint8_t)p5;
    }
    uint32_t total;
    block_exclusive_scan((uint32_t)cnt, s_warp_tot, &total);
    // the last tile to finish turns the tile counts into tile bases (exclusive scan) for the emit launch
    __shared__ bool s_last;
    if (threadIdx.x == 0) {
        tile_tot[blockIdx.x] = total;
        __threadfence();
        s_last = atomicAdd(done, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    uint32_t carry = 0;
    for (int64_t b = 0; b < (int64_t)gridDim.x; b += kCsrTile) {
        const int64_t j = b + threadIdx.x;
        const uint32_t v = j < (int64_t)gridDim.x ? *(volatile uint32_t *)(tile_tot + j) : 0u;
        __syncthreads();                                  // s_warp_tot is reused
        uint32_t tot;
        const uint32_t ex = block_exclusive_scan(v, s_warp_tot, &tot);
        if (j < (int64_t)gridDim.x) tile_tot[j] = carry + ex;
        carry += tot;
    }
    if (threadIdx.x == 0) {
        if (total_out) *total_out = carry;
        *done = 0u;                                       // ready for the next call
    }
}

__global__ void __launch_bounds__(kCsrTile)
movegen_emit_kernel(const uint4 *__restrict__ pos, int64_t n, ck_pos *__restrict__ children, uint32_t child_cap,
                    uint32_t *__restrict__ offsets, const uint32_t *__restrict__ tile_base) {
    __shared__ uint32_t s_warp_tot[kCsrTile / 32];
    __shared__ uint4 s_pos[kCsrTile], s_hop[kCsrTile];    // position; its hop sets (make_child_fast)
    __shared__ uint32_t s_off[kCsrTile];                  // first successor inside the tile | jump flag << 31
    __shared__ uint16_t s_move[kCsrTile * CK_MAX_CHILDREN]; // the tile's move list: source square | direction << 5 | owner thread << 8
    const int tid = threadIdx.x;
    const int64_t i = (int64_t)blockIdx.x * kCsrTile + tid;
    ck_pos p;
    p.p1 = p.p2 = p.k = p.meta = 0;
    uint32_t mask[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int cnt = 0;
    if (i < n) {
        const uint4 v = __ldg(pos + i);
        p.p1 = v.x; p.p2 = v.y; p.k = v.z; p.meta = v.w;
        cnt = gen_moves(p, NullSink{}, mask);
    }
    uint32_t tile_total;
    const uint32_t loc = block_exclusive_scan((uint32_t)cnt, s_warp_tot, &tile_total);
    const uint32_t base = tile_base[blockIdx.x];
    if (i < n && offsets) offsets[i] = base + loc;
    if (children == nullptr) return;
    {
        // per position: the (square, direction) list in the reference's order -- a few instructions per
        // move, so the divergence of this loop is cheap; the heavy part (building the successor) runs
        // below with one lane per successor
        const bool jump = (mask[4] | mask[5] | mask[6] | mask[7]) != 0;
        const Side sd = side_of(p);
        uint32_t J[4];
        hop_sets(sd, J);
        s_pos[tid] = make_uint4(p.p1, p.p2, p.k, p.meta);
        s_hop[tid] = make_uint4(J[0], J[1], J[2], J[3]);
        s_off[tid] = loc | (jump ? 0x80000000u : 0u);
        const uint32_t u0 = jump ? mask[4] : mask[0], u1 = jump ? mask[5] : mask[1], u2 = jump ? mask[6] : mask[2], u3 = jump ? mask[7] : mask[3];
        // The reference's order with the direction slots fixed per position up front instead of selected
        // per move: men in ascending square order, each with its two forward directions (right then left for plain moves,
        // ydir = -1 then +1 for hops); then kings likewise with UL,UR,BL,BR (moves) or UL,BL,UR,BR (hops).  One 16-bit
        // store per move: square | direction << 5 | owner << 8.  (ncu source view, profiles/r2k: this loop was 40 % of the
        // kernel's instructions with per-move direction selects and two byte stores.)
        const bool p0 = sd.player == 0;
        const int fwd = p0 ? 2 : 0;
        const uint32_t f_lo = p0 ? u2 : u0, f_hi = p0 ? u3 : u1;                 // forward-left / forward-right of this side
        const uint32_t ma = (jump ? f_lo : f_hi) & ~sd.kings, mb = (jump ? f_hi : f_lo) & ~sd.kings;
        const uint32_t da = (uint32_t)(jump ? fwd : fwd + 1) << 5, db = (uint32_t)(jump ? fwd + 1 : fwd) << 5;
        const uint32_t own16 = (uint32_t)tid << 8;
        uint32_t m = loc;
        for (uint32_t rem = ma | mb; rem; rem &= rem - 1) {
            const uint32_t s = (uint32_t)ffs32(rem);
            if ((ma >> s) & 1u) s_move[m++] = (uint16_t)(s | da | own16);
            if ((mb >> s) & 1u) s_move[m++] = (uint16_t)(s | db | own16);
        }
        const uint32_t k0 = u0 & sd.kings, k1 = (jump ? u2 : u1) & sd.kings, k2 = (jump ? u1 : u2) & sd.kings, k3 = u3 & sd.kings;
        const uint32_t d1 = (uint32_t)(jump ? 2 : 1) << 5, d2 = (uint32_t)(jump ? 1 : 2) << 5;
        for (uint32_t rem = k0 | k1 | k2 | k3; rem; rem &= rem - 1) {
            const uint32_t s = (uint32_t)ffs32(rem);
            if ((k0 >> s) & 1u) s_move[m++] = (uint16_t)(s | own16);
            if ((k1 >> s) & 1u) s_move[m++] = (uint16_t)(s | d1 | own16);
            if ((k2 >> s) & 1u) s_move[m++] = (uint16_t)(s | d2 | own16);
            if ((k3 >> s) & 1u) s_move[m++] = (uint16_t)(s | (3u << 5) | own16);
        }
    }
    __syncthreads();
    for (uint32_t c = tid; c < tile_total; c += kCsrTile) {
        const uint32_t mv = s_move[c];
        const int lo = (int)(mv >> 8);
        const uint4 pv = s_pos[lo], hv = s_hop[lo];
        ck_pos par;
        par.p1 = pv.x; par.p2 = pv.y; par.k = pv.z; par.meta = pv.w;
        const uint32_t J[4] = {hv.x, hv.y, hv.z, hv.w};
        const ck_pos ch = make_child_fast(par, side_of(par), J, (int)(mv & 31u), (int)((mv >> 5) & 7u), (s_off[lo] >> 31) != 0);
        if ((unsigned long long)base + c < child_cap) reinterpret_cast<uint4 *>(children)[base + c] = make_uint4(ch.p1, ch.p2, ch.k, ch.meta);
    }
}

// ---- K4 ------------------------------------------------------------------------------
// MCTS.default_policy without a net (MCTS.py:132-143): uniform random legal successors until
// determine_outcome reports the end of the game.  Integer-ALU bound; 16 B in, 5 B out.
// Thread per playout.  (A persistent variant -- thread t plays items t, t + T, ... one ply per loop iteration,
// so that no lane waits for its warp's longest playout -- measured no better, 3.75 against 3.55 ms for 2^20
// playouts: lanes of a warp then sit in different phases of their games and the generator's branches diverge
// more; capping the registers for 10 blocks per SM changes nothing either, the kernel is not occupancy-bound.)
__global__ void __launch_bounds__(128)
rollout_kernel(const uint4 *__restrict__ pos, int64_t n, uint64_t seed, int max_plies,
               uint8_t *__restrict__ outcome, int32_t *__restrict__ plies) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint4 v = __ldg(pos + i);
    ck_pos cur;
    cur.p1 = v.x; cur.p2 = v.y; cur.k = v.z; cur.meta = v.w;
    PhiloxChoice choose{Philox(mix64(seed ^ mix64((uint64_t)i))), 0x524F4C4Cu};
    int k = 0, st;
    while ((st = play_step(cur, k, max_plies, choose)) == kPlayMoved) ++k;
    if (outcome) outcome[i] = (uint8_t)st;
    if (plies) plies[i] = k;
}

// ---- Checkers.predict glue -----------------------------------------------------------
__global__ void __launch_bounds__(128)
mask_renorm_kernel(const float *__restrict__ policy, const uint32_t *__restrict__ masks, int64_t n,
                   float *__restrict__ prior) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (row >= n) return;
    uint32_t m[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) m[i] = masks[row * 8 + i];
    float masked[16];
    const float s = masked_policy_sum(policy + row * CK_POLICY_SIZE, m, lane, masked);
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const int a = (lane >> 3) * 128 + i * 8 + (lane & 7);
        prior[row * CK_POLICY_SIZE + a] = __fdiv_rn(masked[i], s);
    }
}

}  // namespace ck

using namespace ck;

extern "C" {

const char *ck_last_error(void) { return g_err.c_str(); }
int ck_abi_version(void) { return CK_ABI_VERSION; }
int ck_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int ck_movegen_device(const ck_pos *d_pos, int64_t n, int32_t max_children, ck_pos *d_children,
                      int32_t *d_counts, uint32_t *d_masks, uint8_t *d_status, uint8_t *d_plane5, void *stream) {
    if (n < 0 || (d_children && (max_children < 1 || max_children > CK_MAX_CHILDREN)))
        return fail(CK_ERR_ARG, "ck_movegen_device: bad n / max_children");
    if (n == 0) return CK_OK;
    int dev = 0;
    CK_CUDA(cudaGetDevice(&dev));
    const int threads = 256;
    int64_t blocks = (n + threads - 1) / threads;
    const int64_t cap = (int64_t)num_sms(dev) * 8 * 4;      // grid-stride above 4 waves of 8 CTAs/SM
    if (blocks > cap) blocks = cap;
    movegen_kernel<<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(
        (const uint4 *)d_pos, n, max_children, d_children, d_counts, (uint4 *)d_masks, d_status, d_plane5);
    CK_CUDA(cudaGetLastError());
    return CK_OK;
}

static void *g_csr_ws[64] = {nullptr};
static size_t g_csr_ws_bytes[64] = {0};

int ck_movegen_csr_device(const ck_pos *d_pos, int64_t n, ck_pos *d_children, int64_t child_cap, uint32_t *d_offsets,
                          uint32_t *d_masks, uint8_t *d_status, uint8_t *d_plane5, void *stream) {
    if (n < 0 || n > (1ll << 26) || child_cap < 0 || child_cap > 0xFFFFFFFFll || !d_pos)
        return fail(CK_ERR_ARG, "ck_movegen_csr_device: bad n / child_cap (n <= 2^26, child_cap < 2^32)");
    int dev = 0;
    CK_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return fail(CK_ERR_ARG, "ck_movegen_csr_device: device index out of range");
    if (n == 0) {
        if (d_offsets) CK_CUDA(cudaMemsetAsync(d_offsets, 0, sizeof(uint32_t), (cudaStream_t)stream));
        return CK_OK;
    }
    const int64_t tiles = (n + kCsrTile - 1) / kCsrTile;
    const size_t need = (size_t)(tiles + 2) * sizeof(uint32_t);               // [tile counts -> bases][total][arrival counter]
    if (g_csr_ws_bytes[dev] < need) {
        CK_CUDA(cudaDeviceSynchronize());
        cudaFree(g_csr_ws[dev]); g_csr_ws[dev] = nullptr; g_csr_ws_bytes[dev] = 0;
        const size_t room = need * 2;
        CK_CUDA(cudaMalloc(&g_csr_ws[dev], room));
        CK_CUDA(cudaMemset(g_csr_ws[dev], 0, room));                          // the counter resets itself after every call
        g_csr_ws_bytes[dev] = room;
    }
    // the counter sits at a fixed place (the end of the allocation) so that calls with different n share it
    uint32_t *tile_tot = (uint32_t *)g_csr_ws[dev];
    unsigned int *done = (unsigned int *)((uint8_t *)g_csr_ws[dev] + g_csr_ws_bytes[dev] - sizeof(unsigned int));
    cudaStream_t st = (cudaStream_t)stream;
    movegen_count_kernel<<<(unsigned)tiles, kCsrTile, 0, st>>>((const uint4 *)d_pos, n, (uint4 *)d_masks, d_status, d_plane5, tile_tot, done,
                                                              d_offsets ? d_offsets + n : tile_tot + tiles);
    movegen_emit_kernel<<<(unsigned)tiles, kCsrTile, 0, st>>>((const uint4 *)d_pos, n, d_children, (uint32_t)child_cap, d_offsets, tile_tot);
    CK_CUDA(cudaGetLastError());
    return CK_OK;
}

int ck_movegen(int device, const ck_pos *pos, int64_t n, int32_t max_children, ck_pos *children,
               int32_t *counts, uint32_t *masks, uint8_t *status, uint8_t *plane5) {
    if (n < 0 || !pos) return fail(CK_ERR_ARG, "ck_movegen: bad arguments");
    if (n == 0) return CK_OK;
    DeviceGuard g(device);
    if (!g.ok) return fail(CK_ERR_CUDA, "ck_movegen: cannot select CUDA device " + std::to_string(device));
    ck_pos *d_pos = nullptr, *d_ch = nullptr;
    int32_t *d_cnt = nullptr; uint32_t *d_mask = nullptr; uint8_t *d_st = nullptr, *d_p5 = nullptr;
    int rc = CK_OK;
    auto cleanup = [&]() {
        cudaFree(d_pos); cudaFree(d_ch); cudaFree(d_cnt); cudaFree(d_mask); cudaFree(d_st); cudaFree(d_p5);
    };
#define CK_TRY(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { rc = fail(CK_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); cleanup(); return rc; } } while (0)
    CK_TRY(cudaMalloc(&d_pos, n * sizeof(ck_pos)));
    CK_TRY(cudaMemcpy(d_pos, pos, n * sizeof(ck_pos), cudaMemcpyHostToDevice));
    if (children) CK_TRY(cudaMalloc(&d_ch, n * max_children * sizeof(ck_pos)));
    if (counts) CK_TRY(cudaMalloc(&d_cnt, n * sizeof(int32_t)));
    if (masks) CK_TRY(cudaMalloc(&d_mask, n * 8 * sizeof(uint32_t)));
    if (status) CK_TRY(cudaMalloc(&d_st, n));
    if (plane5) CK_TRY(cudaMalloc(&d_p5, n));
    if (d_ch) CK_TRY(cudaMemset(d_ch, 0, n * max_children * sizeof(ck_pos)));
    rc = ck_movegen_device(d_pos, n, max_children, d_ch, d_cnt, d_mask, d_st, d_p5, nullptr);
    if (rc != CK_OK) { cleanup(); return rc; }
    CK_TRY(cudaDeviceSynchronize());
    if (children) CK_TRY(cudaMemcpy(children, d_ch, n * max_children * sizeof(ck_pos), cudaMemcpyDeviceToHost));
    if (counts) CK_TRY(cudaMemcpy(counts, d_cnt, n * sizeof(int32_t), cudaMemcpyDeviceToHost));
    if (masks) CK_TRY(cudaMemcpy(masks, d_mask, n * 8 * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    if (status) CK_TRY(cudaMemcpy(status, d_st, n, cudaMemcpyDeviceToHost));
    if (plane5) CK_TRY(cudaMemcpy(plane5, d_p5, n, cudaMemcpyDeviceToHost));
    cleanup();
    return CK_OK;
}

int ck_movegen_csr(int device, const ck_pos *pos, int64_t n, ck_pos *children, int64_t child_cap, uint32_t *offsets,
                   uint32_t *masks, uint8_t *status, uint8_t *plane5) {
    if (n < 0 || !pos || !offsets || child_cap < 0) return fail(CK_ERR_ARG, "ck_movegen_csr: bad arguments");
    DeviceGuard g(device);
    if (!g.ok) return fail(CK_ERR_CUDA, "ck_movegen_csr: cannot select CUDA device " + std::to_string(device));
    if (n == 0) { offsets[0] = 0; return CK_OK; }
    ck_pos *d_pos = nullptr, *d_ch = nullptr;
    uint32_t *d_off = nullptr, *d_mask = nullptr; uint8_t *d_st = nullptr, *d_p5 = nullptr;
    int rc = CK_OK;
    auto cleanup = [&]() { cudaFree(d_pos); cudaFree(d_ch); cudaFree(d_off); cudaFree(d_mask); cudaFree(d_st); cudaFree(d_p5); };
    CK_TRY(cudaMalloc(&d_pos, n * sizeof(ck_pos)));
    CK_TRY(cudaMemcpy(d_pos, pos, n * sizeof(ck_pos), cudaMemcpyHostToDevice));
    CK_TRY(cudaMalloc(&d_off, (n + 1) * sizeof(uint32_t)));
    if (children && child_cap > 0) CK_TRY(cudaMalloc(&d_ch, child_cap * sizeof(ck_pos)));
    if (masks) CK_TRY(cudaMalloc(&d_mask, n * 8 * sizeof(uint32_t)));
    if (status) CK_TRY(cudaMalloc(&d_st, n));
    if (plane5) CK_TRY(cudaMalloc(&d_p5, n));
    rc = ck_movegen_csr_device(d_pos, n, d_ch, d_ch ? child_cap : 0, d_off, d_mask, d_st, d_p5, nullptr);
    if (rc != CK_OK) { cleanup(); return rc; }
    CK_TRY(cudaDeviceSynchronize());
    CK_TRY(cudaMemcpy(offsets, d_off, (n + 1) * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    if (d_ch) {
        const int64_t have = std::min<int64_t>(child_cap, (int64_t)offsets[n]);
        if (have > 0) CK_TRY(cudaMemcpy(children, d_ch, have * sizeof(ck_pos), cudaMemcpyDeviceToHost));
    }
    if (masks) CK_TRY(cudaMemcpy(masks, d_mask, n * 8 * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    if (status) CK_TRY(cudaMemcpy(status, d_st, n, cudaMemcpyDeviceToHost));
    if (plane5) CK_TRY(cudaMemcpy(plane5, d_p5, n, cudaMemcpyDeviceToHost));
    cleanup();
    return CK_OK;
}

int ck_rollout(int device, const ck_pos *pos, int64_t n, uint64_t seed, int32_t max_plies,
               uint8_t *outcome, int32_t *plies) {
    if (n < 0 || !pos) return fail(CK_ERR_ARG, "ck_rollout: bad arguments");
    if (n == 0) return CK_OK;
    DeviceGuard g(device);
    if (!g.ok) return fail(CK_ERR_CUDA, "ck_rollout: cannot select CUDA device " + std::to_string(device));
    ck_pos *d_pos = nullptr; uint8_t *d_out = nullptr; int32_t *d_pl = nullptr;
    int rc = CK_OK;
    auto cleanup = [&]() { cudaFree(d_pos); cudaFree(d_out); cudaFree(d_pl); };
    CK_TRY(cudaMalloc(&d_pos, n * sizeof(ck_pos)));
    CK_TRY(cudaMalloc(&d_out, n));
    CK_TRY(cudaMalloc(&d_pl, n * sizeof(int32_t)));
    CK_TRY(cudaMemcpy(d_pos, pos, n * sizeof(ck_pos), cudaMemcpyHostToDevice));
    rollout_kernel<<<(unsigned)((n + 127) / 128), 128>>>((const uint4 *)d_pos, n, seed, max_plies, d_out, d_pl);
    CK_TRY(cudaGetLastError());
    CK_TRY(cudaDeviceSynchronize());
    if (outcome) CK_TRY(cudaMemcpy(outcome, d_out, n, cudaMemcpyDeviceToHost));
    if (plies) CK_TRY(cudaMemcpy(plies, d_pl, n * sizeof(int32_t), cudaMemcpyDeviceToHost));
    cleanup();
    return CK_OK;
}

int ck_rollout_device(const ck_pos *d_pos, int64_t n, uint64_t seed, int32_t max_plies,
                      uint8_t *d_outcome, int32_t *d_plies, void *stream) {
    if (n < 0 || !d_pos) return fail(CK_ERR_ARG, "ck_rollout_device: bad arguments");
    if (n == 0) return CK_OK;
    rollout_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>((const uint4 *)d_pos, n, seed, max_plies,
                                                                           d_outcome, d_plies);
    CK_CUDA(cudaGetLastError());
    return CK_OK;
}

int ck_mask_renorm(int device, const float *policy, const uint32_t *masks, int64_t n, float *prior) {
    if (n < 0 || !policy || !masks || !prior) return fail(CK_ERR_ARG, "ck_mask_renorm: bad arguments");
    if (n == 0) return CK_OK;
    DeviceGuard g(device);
    if (!g.ok) return fail(CK_ERR_CUDA, "ck_mask_renorm: cannot select CUDA device " + std::to_string(device));
    float *d_pol = nullptr, *d_pri = nullptr; uint32_t *d_m = nullptr;
    int rc = CK_OK;
    auto cleanup = [&]() { cudaFree(d_pol); cudaFree(d_pri); cudaFree(d_m); };
    CK_TRY(cudaMalloc(&d_pol, n * CK_POLICY_SIZE * sizeof(float)));
    CK_TRY(cudaMalloc(&d_pri, n * CK_POLICY_SIZE * sizeof(float)));
    CK_TRY(cudaMalloc(&d_m, n * 8 * sizeof(uint32_t)));
    CK_TRY(cudaMemcpy(d_pol, policy, n * CK_POLICY_SIZE * sizeof(float), cudaMemcpyHostToDevice));
    CK_TRY(cudaMemcpy(d_m, masks, n * 8 * sizeof(uint32_t), cudaMemcpyHostToDevice));
    mask_renorm_kernel<<<(unsigned)((n * 32 + 127) / 128), 128>>>(d_pol, d_m, n, d_pri);
    CK_TRY(cudaGetLastError());
    CK_TRY(cudaDeviceSynchronize());
    CK_TRY(cudaMemcpy(prior, d_pri, n * CK_POLICY_SIZE * sizeof(float), cudaMemcpyDeviceToHost));
    cleanup();
    return CK_OK;
#undef CK_TRY
}

}  // extern "C"
