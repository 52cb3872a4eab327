// ck_tc_ptx.cuh -- inline-PTX wrappers shared by the tcgen05 tower kernels (sm_100a).
#pragma once
#include <cuda_fp16.h>
#include <stdint.h>

namespace ck {
namespace ptx {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc], kind::f16, fp32 accumulate
__device__ __forceinline__ void tc_mma_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem desc]: the A operand (M = 128 rows = TMEM lanes, K = 16 fp16 =
// 8 consecutive 32-bit columns, element 2c in the low half of column c) is read from tensor memory
__device__ __forceinline__ void tc_mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
        "}\n" ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// Warp-converged variants: every lane executes the statement, elect.sync picks the one lane that
// issues.  No C++-level divergence, so warp-uniform operands stay in uniform registers.
__device__ __forceinline__ void tc_mma_ts_elect(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p, e;\n"
        "elect.sync _|e, 0xffffffff;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
        "}\n" ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// One k-step of the split-fp16 product for one tile, issued by the elected lane in a single statement:
//   D (+)= Ahi*Bhi;  D += Ahi*Blo;  D += Alo*Bhi;  commit -> bar
// a_tmem: A hi at columns [a, a+8), A lo at [a+8, a+16); b_lo descriptor = b_hi descriptor + lo_off
// (lo_off in descriptor address units of 16 bytes).
__device__ __forceinline__ void tc_kstep_ts_elect(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_hi, uint32_t lo_off, uint32_t idesc,
                                                  uint32_t accumulate, uint32_t bar) {
    asm volatile(
        "{\n"
        ".reg .pred p, e, t;\n"
        ".reg .b64 blo, off;\n"
        ".reg .b32 alo;\n"
        "elect.sync _|e, 0xffffffff;\n"
        "setp.ne.b32 p, %5, 0;\n"
        "setp.eq.b32 t, 0, 0;\n"
        "cvt.u64.u32 off, %3;\n"
        "add.u64 blo, %2, off;\n"
        "add.u32 alo, %1, 8;\n"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %4, p;\n"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], blo, %4, t;\n"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [alo], %2, %4, t;\n"
        "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%6];\n"
        "}\n" ::"r"(d_tmem), "r"(a_tmem), "l"(b_hi), "r"(lo_off), "r"(idesc), "r"(accumulate), "r"(bar) : "memory");
}
// Two MMAs on one weight slot, then the commit that frees it:  D (+)= A0*B0;  D += A1*B1;  commit -> bar
// (a0 / a1: TMEM column addresses of the two 8-column operand halves of the slot, b0 / b1: shared-memory descriptors)
__device__ __forceinline__ void tc_pair_ts_elect(uint32_t d_tmem, uint32_t a0, uint64_t b0, uint32_t a1, uint64_t b1, uint32_t idesc,
                                                 uint32_t accumulate, uint32_t bar) {
    asm volatile(
        "{\n"
        ".reg .pred p, e, t;\n"
        "elect.sync _|e, 0xffffffff;\n"
        "setp.ne.b32 p, %6, 0;\n"
        "setp.eq.b32 t, 0, 0;\n"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %5, p;\n"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [%3], %4, %5, t;\n"
        "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%7];\n"
        "}\n" ::"r"(d_tmem), "r"(a0), "l"(b0), "r"(a1), "l"(b1), "r"(idesc), "r"(accumulate), "r"(bar) : "memory");
}
// Two weight slots in one go: four MMAs, each slot's commit right after its pair
__device__ __forceinline__ void tc_quad_ts_elect(uint32_t d_tmem, uint32_t a0, uint64_t b0, uint32_t a1, uint64_t b1,
                                                 uint32_t a2, uint64_t b2, uint32_t a3, uint64_t b3, uint32_t idesc,
                                                 uint32_t accumulate, uint32_t bar_a, uint32_t bar_b) {
    asm volatile(
        "{\n"
        ".reg .pred p, e, t;\n"
        "elect.sync _|e, 0xffffffff;\n"
        "setp.ne.b32 p, %10, 0;\n"
        "setp.eq.b32 t, 0, 0;\n"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %9, p;\n"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [%3], %4, %9, t;\n"
        "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%11];\n"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [%5], %6, %9, t;\n"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [%7], %8, %9, t;\n"
        "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%12];\n"
        "}\n" ::"r"(d_tmem), "r"(a0), "l"(b0), "r"(a1), "l"(b1), "r"(a2), "l"(b2), "r"(a3), "l"(b3), "r"(idesc), "r"(accumulate),
        "r"(bar_a), "r"(bar_b) : "memory");
}
// Four weight slots in one go (a[2i], a[2i+1] / b[2i], b[2i+1]: the two MMAs of slot i; bar[i]: its commit)
__device__ __forceinline__ void tc_oct_ts_elect(uint32_t d_tmem, const uint32_t a[8], const uint64_t b[8], uint32_t idesc,
                                                uint32_t accumulate, const uint32_t bar[4]) {
    asm volatile(
        "{\n"
        ".reg .pred p, e, t;\n"
        "elect.sync _|e, 0xffffffff;\n"
        "setp.ne.b32 p, %18, 0;\n"
        "setp.eq.b32 t, 0, 0;\n"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %9, %17, p;\n"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [%2], %10, %17, t;\n"
        "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%19];\n"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [%3], %11, %17, t;\n"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [%4], %12, %17, t;\n"
        "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%20];\n"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [%5], %13, %17, t;\n"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [%6], %14, %17, t;\n"
        "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%21];\n"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [%7], %15, %17, t;\n"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [%8], %16, %17, t;\n"
        "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%22];\n"
        "}\n" ::"r"(d_tmem), "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]),
        "l"(b[0]), "l"(b[1]), "l"(b[2]), "l"(b[3]), "l"(b[4]), "l"(b[5]), "l"(b[6]), "l"(b[7]), "r"(idesc), "r"(accumulate),
        "r"(bar[0]), "r"(bar[1]), "r"(bar[2]), "r"(bar[3]) : "memory");
}
__device__ __forceinline__ void tc_commit_elect(uint32_t bar) {
    asm volatile(
        "{\n"
        ".reg .pred e;\n"
        "elect.sync _|e, 0xffffffff;\n"
        "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n"
        "}\n" ::"r"(bar) : "memory");
}
// non-blocking phase test
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
// UMMA shared-memory descriptor, K-major, SWIZZLE_NONE (cute::UMMA::SmemDescriptor):
// [0,14) start>>4, [16,30) LBO>>4 (stride between the two 8-element K chunks of one MMA),
// [32,46) SBO>>4 (stride between 8-row groups along M/N), [46,48) version = 1, layout_type 0.
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((addr >> 4) & 0x3FFFu) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
// kind::f16 instruction descriptor (cute::UMMA::InstrDescriptor): c_format F32 (bit 4), A and B
// F16 (0), both K-major, N>>3 at [17,23), M>>4 at [24,29)
__host__ __device__ constexpr uint32_t make_idesc_f16(int m, int n) {
    return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float v[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
        "tcgen05.wait::ld.sync.aligned;\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
// asynchronous variant: the registers are only valid after tmem_ld_wait16 on the same array
__device__ __forceinline__ void tmem_ld16_async(uint32_t taddr, uint32_t r[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
}
// waits for every tcgen05.ld issued so far; the "+r" operands tie the loaded registers to the wait
// so that the compiler cannot move their uses above it
__device__ __forceinline__ void tmem_ld_wait16(uint32_t r[16]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;\n"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                   "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
                 :: "memory");
}
// 16 lanes x 64 columns in the mma-fragment layout: for column group j (8 columns) thread t holds
//   r[4j+0], r[4j+1] = (lane t/4,     columns 8j + 2(t%4), +1)
//   r[4j+2], r[4j+3] = (lane t/4 + 8, columns 8j + 2(t%4), +1)
// (cute::SM100_TMEM_LOAD_16dp256b8x).  Asynchronous: valid after tmem_ld_wait32 on the same array.
__device__ __forceinline__ void tmem_ld_16x256b_x8_async(uint32_t taddr, uint32_t r[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.16x256b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait32(uint32_t r[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;\n"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                   "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]),
                   "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]),
                   "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
                 :: "memory");
}
// Four 8x8 b16 matrices, transposed on the way out: register i holds matrix i in the mma-fragment layout
// (thread t: row t/4, columns 2(t%4) and +1); stored row j of matrix i (16 bytes = fragment column j,
// fragment rows 0..7) goes to the address supplied by thread 8i + j.
__device__ __forceinline__ void stmatrix_x4_trans(uint32_t addr, uint32_t r0, uint32_t r1, uint32_t r2, uint32_t r3) {
    asm volatile("stmatrix.sync.aligned.m8n8.x4.trans.shared.b16 [%0], {%1, %2, %3, %4};\n"
                 :: "r"(addr), "r"(r0), "r"(r1), "r"(r2), "r"(r3) : "memory");
}
// 32 lanes x 16 columns: register i of thread t -> TMEM lane (lane base + t), column (col base + i)
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t r[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n"
        :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
           "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

template <int kCols>
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "n"(kCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int kCols>
__device__ __forceinline__ void tmem_dealloc(uint32_t base) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "n"(kCols) : "memory");
}

}  // namespace ptx
}  // namespace ck
