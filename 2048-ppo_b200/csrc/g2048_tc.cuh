// g2048_tc.cuh -- minimal tcgen05 / TMEM toolkit for sm_100a (inline PTX, no CUTLASS).
//
// Operand tiles are bf16, K-major, 128-byte swizzled ("canonical" UMMA layout
// Swizzle<3,4,3> o ((8,n),2):((8,SBO),1) in 16-byte units):
//   a tile of R rows x 64 K-elements is R rows of 128 bytes; 8 consecutive rows form a
//   1024-byte atom; inside an atom the 16-byte unit u of row r sits at unit (u ^ (r & 7)).
//   Wider K = several such [R x 64] blocks one after the other.
// One tcgen05.mma (kind::f16, bf16 x bf16 -> fp32) consumes K = 16 elements = 32 bytes of every
// row: k-step j of a block starts 32*j bytes into the block.
// The accumulator D[M=128][N] lives in TMEM: lane = row, column = n (32-bit each).
#pragma once
#include <cstdint>
#include <cuda_bf16.h>

namespace g2048 {
namespace tc {

constexpr int BLOCK_K = 64;                 // bf16 elements per 128-byte swizzle row
constexpr int UMMA_K = 16;                  // K per tcgen05.mma for 16-bit inputs

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return uint32_t(__cvta_generic_to_shared(p)); }

// byte offset of element (row r, k) inside a tile of `rows` rows (k may exceed 64: next block)
__host__ __device__ __forceinline__ uint32_t sw128_offset(int rows, int r, int k) {
    const int blk = k >> 6, kk = k & 63;
    return uint32_t(blk) * uint32_t(rows) * 128u + uint32_t(r >> 3) * 1024u + uint32_t(r & 7) * 128u +
           uint32_t(((kk >> 3) ^ (r & 7)) << 4) + uint32_t(kk & 7) * 2u;
}

// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start>>4 [0,14), LBO>>4 [16,30),
// SBO>>4 [32,46), version=1 [46,48), base_offset [49,52), layout_type [61,64) (2 = SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr) {
    uint64_t d = 0;
    d |= uint64_t((saddr >> 4) & 0x3FFFu);
    d |= uint64_t(1) << 16;                  // LBO: unused for swizzled K-major, canonical value 1
    d |= uint64_t(1024 >> 4) << 32;          // SBO: 8 rows x 128 B between row groups
    d |= uint64_t(1) << 46;                  // descriptor version (Blackwell)
    d |= uint64_t(2) << 61;                  // SWIZZLE_128B
    return d;
}

// instruction descriptor (cute::UMMA::InstrDescriptor) for kind::f16: C=F32 [4,6)=1, A=BF16 [7,10)=1,
// B=BF16 [10,13)=1, A,B K-major (bits 15,16 = 0), N>>3 at [17,23), M>>4 at [24,29)
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | (uint32_t(N >> 3) << 17) | (uint32_t(M >> 4) << 24);
}

// ---- 32-byte-swizzled tiles (used by the update GEMMs, g2048_linear.cu) --------------------------
// A block is R rows of 32 bytes = 16 bf16 = exactly one k-step; 8 rows form a 256-byte atom and the
// 16-byte unit u of row r sits at unit u ^ ((r >> 2) & 1) (Swizzle<1,4,3>: address bit 4 ^= bit 7).
// The same bytes are a K-major operand (rows = M/N index, 16 K-elements per row; SBO = 256) and an
// MN-major operand (rows = K index, 16 MN-elements per row; SBO = 256 between 8-row K groups, LBO =
// distance between consecutive blocks = the next 16 MN-elements).
__host__ __device__ __forceinline__ uint32_t sw32_offset(int r, int k) {
    return uint32_t(r) * 32u + uint32_t((((k >> 3) ^ (r >> 2)) & 1) << 4) + uint32_t(k & 7) * 2u;
}
__device__ __forceinline__ uint64_t make_desc_sw32(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= uint64_t((saddr >> 4) & 0x3FFFu);
    d |= uint64_t((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= uint64_t((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= uint64_t(1) << 46;                  // descriptor version (Blackwell)
    d |= uint64_t(6) << 61;                  // SWIZZLE_32B
    return d;
}
// kind::f16 instruction descriptor with explicit operand majors (bit 15: A is MN-major, bit 16: B)
__host__ __device__ constexpr uint32_t make_idesc_bf16_major(int M, int N, bool a_mn, bool b_mn) {
    return make_idesc_bf16(M, N) | (uint32_t(a_mn) << 15) | (uint32_t(b_mn) << 16);
}

__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {   // whole warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(smem_dst)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {     // whole warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// make generic-proxy st.shared visible to the async proxy (tcgen05.mma operand reads)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem]^T, issued by ONE thread
__device__ __forceinline__ void mma_bf16_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, bool accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(uint32_t(accumulate))
        : "memory");
}
// arrive on an mbarrier once all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_addr(bar)) : "memory");
}

// 16 consecutive fp32 columns of this thread's TMEM lane (lane = 32*(warp%4) + laneid)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
// 8 consecutive fp32 columns
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* v) {
    uint32_t r[8];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
        : "r"(taddr)
        : "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
// two 8-column loads in flight, one wait
__device__ __forceinline__ void tmem_ld8x2(uint32_t ta, float* a, uint32_t tb, float* b) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%16];\n\t"
        "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%8,%9,%10,%11,%12,%13,%14,%15}, [%17];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(ta), "r"(tb)
        : "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        a[i] = __uint_as_float(r[i]);
        b[i] = __uint_as_float(r[8 + i]);
    }
}
// Split-phase 16-column load for register-resident rows: issue any number of loads, then
// tmem_ld_wait_all(), then pass every destination register through tmem_ld_pin() -- an empty
// volatile asm that keeps the compiler from scheduling a use of the register above the wait.
__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld8_issue(uint32_t taddr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tmem_ld_wait_all() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float tmem_ld_pin(uint32_t r) {
    asm volatile("" : "+r"(r)::"memory");
    return __uint_as_float(r);
}

// 16 columns, load and wait fused in one asm statement so that no use can be scheduled in between
__device__ __forceinline__ void tmem_ld16p(uint32_t taddr, float* v) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float* v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr),
                 "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
                 "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])),
                 "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7]))
                 : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
        "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
        "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
        "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
        "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15]))
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---- fp16 operands (kind::f16 with A = B = F16: format fields 0) -------------------------------------------------
// Two fp16 terms hold 22 mantissa bits of an fp32 value (x = hi + lo, |x| < 65504), so the three products
// lo*hi + hi*lo + hi*hi are fp32-grade (~2e-7 of scale) where two bf16 terms give 16 bits (~1e-5).
__host__ __device__ constexpr uint32_t make_idesc_f16(int M, int N) {
    return (1u << 4) | (uint32_t(N >> 3) << 17) | (uint32_t(M >> 4) << 24);
}
// two fp32 values -> packed fp16 hi terms and packed fp16 lo terms (a in the low half).  SASS: F2FP, 2 x FHFMA
// (fp32 += fp16 * fp16, sm_100+), F2FP.  The conversions saturate (F2FP.SATFINITE, same cost): a value beyond the fp16 range
// becomes +-65504 (+ a saturated lo term: hi + lo reaches +-131008) instead of an infinity that turns the whole MMA row into NaN --
// out of range the result is then merely inaccurate, like the fp32 reference's own ill-conditioned regime, not poisoned.
__device__ __forceinline__ void split2_f16(float a, float b, uint32_t& hi, uint32_t& lo) {
    float ra, rb;
    asm("{\n\t.reg .f16 l, u, m;\n\t"
        "cvt.rn.satfinite.f16x2.f32 %0, %4, %3;\n\t"
        "mov.b32 {l, u}, %0;\n\t"
        "mov.b16 m, 0xBC00;\n\t"                 // -1.0
        "fma.rn.f32.f16 %1, l, m, %3;\n\t"
        "fma.rn.f32.f16 %2, u, m, %4;\n\t}"
        : "=&r"(hi), "=f"(ra), "=f"(rb)
        : "f"(a), "f"(b));
    asm("cvt.rn.satfinite.f16x2.f32 %0, %2, %1;" : "=r"(lo) : "f"(ra), "f"(rb));
}
__device__ __forceinline__ void tmem_ld4_issue(uint32_t taddr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tmem_st4(uint32_t taddr, const float* v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(taddr), "r"(__float_as_uint(v[0])),
                 "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3]))
                 : "memory");
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
// shared -> global bulk copy (bulk async-group completion): issue, commit, wait until the source may be overwritten
// (`read`) or until the writes are complete
__device__ __forceinline__ void bulk_s2g(void* gdst, const void* ssrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_addr(ssrc)), "r"(bytes) : "memory");
}
// the same with an L2 eviction policy (createpolicy): streaming output that nobody in this kernel reads again should not push the
// kernel's own L2-resident scratch out
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void bulk_s2g_hint(void* gdst, const void* ssrc, uint32_t bytes, uint64_t policy) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(gdst), "r"(smem_addr(ssrc)), "r"(bytes),
                 "l"(policy)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// raises the pending transaction count WITHOUT arriving (mbar_expect_tx above is arrive + expect_tx)
__device__ __forceinline__ void mbar_add_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "TW_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra TD_%=;\n\t"
        "bra TW_%=;\n\t"
        "TD_%=:\n\t}" ::"r"(smem_addr(bar)),
        "r"(parity)
        : "memory");
}
// one non-blocking probe of a phase; true once the phase with this parity has completed
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_addr(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_addr(dst)),
                 "l"(src), "r"(bytes), "r"(smem_addr(bar))
                 : "memory");
}

}  // namespace tc
}  // namespace g2048
