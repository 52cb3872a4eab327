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
// layout touches 768 B of address space per position to store ~70 B, which is what kept K1 at 15 % of
// the HBM roofline; here the kernel moves the algorithmic bytes only (16 B in, 4 B offset, 32 B mask,
// 2 B status/plane5 and 16 B per child out).  One pass: tiles of 256 positions take a ticket, scan
// their counts (warp shuffles + shared memory) and chain the tile totals with a decoupled look-back
// over 64-bit {flag, value} words, so offsets are deterministic and no second sweep is needed.
constexpr int kCsrTile = 256;
constexpr unsigned long long kCsrAgg = 1ull << 62, kCsrIncl = 2ull << 62, kCsrVal = (1ull << 62) - 1;

__global__ void __launch_bounds__(kCsrTile)
movegen_csr_kernel(const uint4 *__restrict__ pos, int64_t n, ck_pos *__restrict__ children, uint32_t child_cap,
                   uint32_t *__restrict__ offsets, uint4 *__restrict__ masks, uint8_t *__restrict__ status,
                   uint8_t *__restrict__ plane5, unsigned int *__restrict__ ticket, unsigned long long *__restrict__ tile_state) {
    __shared__ unsigned int s_tile;
    __shared__ uint32_t s_warp_tot[kCsrTile / 32];
    __shared__ unsigned long long s_base;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    if (tid == 0) s_tile = atomicAdd(ticket, 1u);          // tiles are chained in ticket order: every predecessor is already running
    __syncthreads();
    const unsigned int tile = s_tile;
    const int64_t i = (int64_t)tile * kCsrTile + tid;
    ck_pos p;
    p.p1 = p.p2 = p.k = p.meta = 0;
    uint32_t mask[8];
    int cnt = 0;
    if (i < n) {
        const uint4 v = __ldg(pos + i);
        p.p1 = v.x; p.p2 = v.y; p.k = v.z; p.meta = v.w;
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
    // exclusive scan of the counts inside the tile
    uint32_t incl = (uint32_t)cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, o);
        if (lane >= o) incl += v;
    }
    if (lane == 31) s_warp_tot[w] = incl;
    __syncthreads();
    uint32_t warp_off = 0, tile_total = 0;
#pragma unroll
    for (int q = 0; q < kCsrTile / 32; ++q) {
        const uint32_t v = s_warp_tot[q];
        if (q < w) warp_off += v;
        tile_total += v;
    }
    const uint32_t loc = warp_off + incl - (uint32_t)cnt;
    // chain the tile totals (decoupled look-back, one warp, 32 predecessors per probe)
    if (w == 0) {
        unsigned long long base = 0;
        if (lane == 0) atomicExch(tile_state + tile, (tile == 0 ? kCsrIncl : kCsrAgg) | tile_total);
        if (tile > 0) {
            int64_t look = (int64_t)tile - 1;
            for (;;) {
                const int64_t idx = look - lane;
                unsigned long long v = kCsrIncl;                       // before tile 0: inclusive prefix 0
                if (idx >= 0) v = *reinterpret_cast<volatile unsigned long long *>(tile_state + idx);
                if (__any_sync(0xFFFFFFFFu, (v >> 62) == 0)) continue; // a predecessor has not published yet
                const uint32_t incl_mask = __ballot_sync(0xFFFFFFFFu, (v >> 62) == 2);
                const int first = incl_mask ? __ffs((int)incl_mask) - 1 : 31;
                unsigned long long part = lane <= first ? (v & kCsrVal) : 0ull;
#pragma unroll
                for (int o = 16; o; o >>= 1) part += __shfl_xor_sync(0xFFFFFFFFu, part, o);
                base += part;
                if (incl_mask) break;
                look -= 32;
            }
            if (lane == 0) atomicExch(tile_state + tile, kCsrIncl | (base + tile_total));
        }
        if (lane == 0) {
            s_base = base;
            if ((int64_t)(tile + 1) * kCsrTile >= n && offsets) offsets[n] = (uint32_t)(base + tile_total);
        }
    }
    __syncthreads();
    const unsigned long long off = s_base + loc;
    if (i < n) {
        if (offsets) offsets[i] = (uint32_t)off;
        if (children != nullptr && cnt > 0) {
            const int room = off >= child_cap ? 0 : (int)min((unsigned long long)cnt, (unsigned long long)child_cap - off);
            if (room > 0) gen_moves(p, ArraySink{children + off, room}, mask);
        }
    }
}

// ---- K4 ------------------------------------------------------------------------------
// MCTS.default_policy without a net (MCTS.py:132-143): uniform random legal successors until
// determine_outcome reports the end of the game.  Integer-ALU bound; 16 B in, 5 B out.
__global__ void __launch_bounds__(128)
rollout_kernel(const uint4 *__restrict__ pos, int64_t n, uint64_t seed, int max_plies,
               uint8_t *__restrict__ outcome, int32_t *__restrict__ plies) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint4 v = __ldg(pos + i);
    ck_pos cur;
    cur.p1 = v.x; cur.p2 = v.y; cur.k = v.z; cur.meta = v.w;
    const Philox rng(mix64(seed ^ mix64((uint64_t)i)));
    int k = 0, st;
    uint32_t r[4];
    for (;;) {
        uint32_t mask[8];
        const int cnt = gen_moves(cur, NullSink{}, mask);
        st = outcome_of(cur, cnt > 0, nullptr);
        if (st != CK_ONGOING) break;
        if (max_plies > 0 && k >= max_plies) break;
        if ((k & 3) == 0) rng((uint32_t)(k >> 2), 0u, 0u, 0x524F4C4Cu, r);
        const int pick = (int)(((uint64_t)r[k & 3] * (uint64_t)cnt) >> 32);
        ck_pos nxt = cur;
        gen_moves(cur, PickSink{&nxt, pick}, mask);
        cur = nxt;
        ++k;
    }
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
    const size_t need = 16 + (size_t)tiles * sizeof(unsigned long long);      // [ticket, pad][tile states]
    if (g_csr_ws_bytes[dev] < need) {
        CK_CUDA(cudaDeviceSynchronize());
        cudaFree(g_csr_ws[dev]); g_csr_ws[dev] = nullptr; g_csr_ws_bytes[dev] = 0;
        CK_CUDA(cudaMalloc(&g_csr_ws[dev], need));
        g_csr_ws_bytes[dev] = need;
    }
    CK_CUDA(cudaMemsetAsync(g_csr_ws[dev], 0, need, (cudaStream_t)stream));
    movegen_csr_kernel<<<(unsigned)tiles, kCsrTile, 0, (cudaStream_t)stream>>>(
        (const uint4 *)d_pos, n, d_children, (uint32_t)child_cap, d_offsets, (uint4 *)d_masks, d_status, d_plane5,
        (unsigned int *)g_csr_ws[dev], (unsigned long long *)((uint8_t *)g_csr_ws[dev] + 16));
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
