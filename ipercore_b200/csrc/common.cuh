// Shared device helpers for the iper_b200 kernels (sm_100a only): mbarrier, TMA, tcgen05/TMEM PTX wrappers.
// Hand-written inline PTX; no CUTLASS/CuTe dependency.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_fp8.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <cstdio>

#define IPER_DEVINL __device__ __forceinline__

namespace iper {

// ---------------------------------------------------------------------------------------------------------
// error plumbing shared by the C-ABI entry points (api.cu owns the storage)
// ---------------------------------------------------------------------------------------------------------
void set_last_error(const char* fmt, ...);
#define IPER_CHECK_CUDA(expr)                                                                       \
    do {                                                                                            \
        cudaError_t _e = (expr);                                                                    \
        if (_e != cudaSuccess) {                                                                    \
            iper::set_last_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__,  \
                                 __LINE__);                                                         \
            return 2;                                                                               \
        }                                                                                           \
    } while (0)
#define IPER_REQUIRE(cond, ...)                  \
    do {                                         \
        if (!(cond)) {                           \
            iper::set_last_error(__VA_ARGS__);   \
            return 1;                            \
        }                                        \
    } while (0)

// ---------------------------------------------------------------------------------------------------------
// small utilities
// ---------------------------------------------------------------------------------------------------------
IPER_DEVINL uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

IPER_DEVINL uint32_t elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t.reg .pred P;\n\t.reg .b32 R;\n\t"
        "elect.sync R|P, 0xffffffff;\n\t"
        "selp.b32 %0, 1, 0, P;\n\t}"
        : "=r"(pred));
    return pred;
}

// ---------------------------------------------------------------------------------------------------------
// mbarrier
// ---------------------------------------------------------------------------------------------------------
IPER_DEVINL void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
IPER_DEVINL void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
IPER_DEVINL void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

IPER_DEVINL void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
IPER_DEVINL void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
IPER_DEVINL uint32_t mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred P;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, P;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok;
}
// Bounded wait: a pipeline bug must surface as a trapped kernel (an error the host sees), never as a hang.
IPER_DEVINL void mbar_wait(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 4000000000LL) {  // ~2 s at 2 GHz
            printf("iper_b200: mbarrier wait timed out (block %d thread %d)\n", (int)blockIdx.x, (int)threadIdx.x);
            __trap();
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// TMA (cp.async.bulk.tensor) loads, tile mode, completion on an mbarrier
// ---------------------------------------------------------------------------------------------------------
IPER_DEVINL void tma_prefetch_desc(const void* desc) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(desc)) : "memory");
}
IPER_DEVINL void tma_load_2d(void* smem, const void* desc, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem)), "l"(reinterpret_cast<uint64_t>(desc)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
IPER_DEVINL void tma_load_4d(void* smem, const void* desc, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(smem)), "l"(reinterpret_cast<uint64_t>(desc)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2),
        "r"(c3)
        : "memory");
}
IPER_DEVINL void tma_load_5d(void* smem, const void* desc, uint64_t* bar, int c0, int c1, int c2, int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(smem_u32(smem)), "l"(reinterpret_cast<uint64_t>(desc)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2),
        "r"(c3), "r"(c4)
        : "memory");
}

// ---------------------------------------------------------------------------------------------------------
// tcgen05 / TMEM
// ---------------------------------------------------------------------------------------------------------
IPER_DEVINL void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {  // one full warp, ncols power of two >= 32
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
                 "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
IPER_DEVINL void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // same warp that allocated
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
IPER_DEVINL void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
IPER_DEVINL void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc], kind::f16 (fp16/bf16 operands, fp32 accumulate), issued by ONE thread
IPER_DEVINL void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// same, kind::f8f6f4 (e4m3 operands, K = 32 per instruction, fp32 accumulate)
IPER_DEVINL void umma_f8(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// ---- CTA-pair (cta_group::2) forms: the leader CTA of a 2-CTA cluster issues MMAs that read A/B from BOTH CTAs' smem ----
IPER_DEVINL uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
IPER_DEVINL void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
IPER_DEVINL void mbar_arrive_remote(uint64_t* bar, uint32_t cta) {   // arrive on the same-offset barrier of CTA `cta`
    asm volatile(
        "{\n\t.reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
        "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}"
        ::"r"(smem_u32(bar)), "r"(cta)
        : "memory");
}
constexpr uint32_t PEER_BIT_MASK = 0xFEFFFFFFu;   // clears the CTA-rank bit: the barrier of CTA 0 of the pair
IPER_DEVINL void tma_load_2d_2sm(void* smem, const void* desc, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem)), "l"(reinterpret_cast<uint64_t>(desc)), "r"(smem_u32(bar) & PEER_BIT_MASK), "r"(c0), "r"(c1)
        : "memory");
}
IPER_DEVINL void tma_load_4d_2sm(void* smem, const void* desc, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(smem)), "l"(reinterpret_cast<uint64_t>(desc)), "r"(smem_u32(bar) & PEER_BIT_MASK), "r"(c0), "r"(c1),
        "r"(c2), "r"(c3)
        : "memory");
}
IPER_DEVINL void tma_load_5d_2sm(void* smem, const void* desc, uint64_t* bar, int c0, int c1, int c2, int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(smem_u32(smem)), "l"(reinterpret_cast<uint64_t>(desc)), "r"(smem_u32(bar) & PEER_BIT_MASK), "r"(c0), "r"(c1),
        "r"(c2), "r"(c3), "r"(c4)
        : "memory");
}
IPER_DEVINL void tmem_alloc_2sm(uint32_t* dst_smem, uint32_t ncols) {   // same warp id in both CTAs, same dst offset
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
IPER_DEVINL void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem of both CTAs] (+)= A[256 rows: 128 from each CTA] * B[N rows: N/2 from each CTA]
IPER_DEVINL void umma_f16_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
IPER_DEVINL void umma_f8_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive on the same-offset mbarrier of every CTA in `mask` once the MMAs issued so far have completed
IPER_DEVINL void umma_commit_2sm(uint64_t* bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(mask)
                 : "memory");
}

// arrive on an mbarrier once all previously issued MMAs of this thread have completed
IPER_DEVINL void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
// 32 lanes x 32 consecutive fp32 columns -> 32 registers per thread (thread = TMEM lane = accumulator row)
IPER_DEVINL void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
IPER_DEVINL void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// registers -> 32 lanes x 32 consecutive fp32 columns (the inverse of tmem_ld32)
IPER_DEVINL void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
          "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]),
          "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]),
          "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
        : "memory");
}
IPER_DEVINL void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor for a K-major operand tile stored as rows of 128 bytes with the 128-byte
// swizzle (what TMA writes with CU_TENSOR_MAP_SWIZZLE_128B): 8-row groups every 1024 bytes (SBO), LBO unused.
// bit layout (tcgen05 matrix descriptor): [0,14) addr>>4, [16,30) LBO>>4, [32,46) SBO>>4, [46,48) version=1,
// [49,52) base offset, [61,64) layout type (2 = SWIZZLE_128B).
IPER_DEVINL uint64_t umma_desc_sw128(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;             // LBO (ignored for swizzled K-major; 1 by convention)
    d |= (uint64_t)(1024 >> 4) << 32;   // SBO
    d |= (uint64_t)1 << 46;             // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;             // SWIZZLE_128B
    return d;
}
// K-major operand tile stored as rows of 64 bytes with the 64-byte swizzle (8-bit operands, 64 elements per row):
// 8-row groups every 512 bytes, layout type 4 = SWIZZLE_64B.
IPER_DEVINL uint64_t umma_desc_sw64(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(512 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)4 << 61;
    return d;
}
// Instruction descriptor, kind::f8f6f4 with e4m3 (format 0) A/B, fp32 D.
IPER_DEVINL constexpr uint32_t umma_idesc_e4m3(int M, int N) {
    return (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// Instruction descriptor, kind::f16: fp16 A/B (format 0), fp32 D (format 1), both K-major, dense.
// [4,6) D fmt, [7,10) A fmt, [10,13) B fmt, [15] A major, [16] B major, [17,23) N>>3, [24,29) M>>4.
IPER_DEVINL constexpr uint32_t umma_idesc_f16(int M, int N) {
    return (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ---------------------------------------------------------------------------------------------------------
// activation "planes" formats (value of the *_planes arguments of the C ABI)
//   1  FMT_H   : one fp16 plane                                   x ~ hi                      (11 bits)
//   2  FMT_HL  : fp16 hi + fp16 lo                                x ~ hi + lo                 (~22 bits)
//   3  FMT_H8  : fp16 hi + two e4m3 planes in the space of the lo plane:
//                a8 = e4m3(x * 2^3), l8 = e4m3((x - hi) * 2^14)   x ~ hi + l8 / 2^14         (~15 bits)
//                (the 8-bit planes feed the fp8 cross-term MMAs: hi*w_lo and lo*w_hi need only ~4 bits each)
// A plane holds `plane_stride` elements; the 8-bit planes start at byte offset 2*plane_stride (a8) and
// 3*plane_stride (l8) from the tensor base.
// ---------------------------------------------------------------------------------------------------------
constexpr int FMT_H = 1, FMT_HL = 2, FMT_H8 = 3;
constexpr float ACT_S8 = 8.0f;          // 2^3
constexpr float ACT_SL8 = 16384.0f;     // 2^14  ( = ACT_S8 * 2^11 )

IPER_DEVINL uint8_t to_e4m3(float v) { return (uint8_t)__nv_cvt_float_to_fp8(v, __NV_SATFINITE, __NV_E4M3); }
IPER_DEVINL float from_e4m3(uint8_t b) {
    const __half_raw h = __nv_cvt_fp8_to_halfraw((__nv_fp8_storage_t)b, __NV_E4M3);
    return __half2float(*reinterpret_cast<const __half*>(&h));
}
IPER_DEVINL const uint8_t* plane_a8(const __half* x, long long plane_stride) {
    return reinterpret_cast<const uint8_t*>(x + plane_stride);
}
IPER_DEVINL const uint8_t* plane_l8(const __half* x, long long plane_stride) {
    return reinterpret_cast<const uint8_t*>(x + plane_stride) + plane_stride;
}
IPER_DEVINL float load_plane_val(const __half* x, int fmt, long long plane_stride, size_t off) {
    float v = __half2float(x[off]);
    if (fmt == FMT_HL) v += __half2float(x[plane_stride + off]);
    else if (fmt == FMT_H8) v += from_e4m3(plane_l8(x, plane_stride)[off]) * (1.0f / ACT_SL8);
    return v;
}
IPER_DEVINL void store_plane_val(__half* o, int fmt, long long plane_stride, size_t off, float v) {
    const __half hi = __float2half_rn(v);
    const float lo = v - __half2float(hi);
    o[off] = hi;
    if (fmt == FMT_HL) o[plane_stride + off] = __float2half_rn(lo);
    else if (fmt == FMT_H8) {
        uint8_t* a8 = reinterpret_cast<uint8_t*>(o + plane_stride);
        a8[off] = to_e4m3(v * ACT_S8);
        a8[plane_stride + off] = to_e4m3(lo * ACT_SL8);
    }
}
// 8 consecutive channels (16-byte aligned) of one pixel
IPER_DEVINL void load_planes8(const __half* x, int fmt, long long plane_stride, size_t off, float (&v)[8]) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(x + off));
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[j]));
        v[2 * j] = f.x; v[2 * j + 1] = f.y;
    }
    if (fmt == FMT_HL) {
        const uint4 u2 = __ldg(reinterpret_cast<const uint4*>(x + plane_stride + off));
        const uint32_t w2[4] = {u2.x, u2.y, u2.z, u2.w};
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w2[j]));
            v[2 * j] += f.x; v[2 * j + 1] += f.y;
        }
    } else if (fmt == FMT_H8) {
        const uint2 l = __ldg(reinterpret_cast<const uint2*>(plane_l8(x, plane_stride) + off));
        const uint32_t w2[2] = {l.x, l.y};
#pragma unroll
        for (int j = 0; j < 8; j++) v[j] += from_e4m3((uint8_t)(w2[j >> 2] >> (8 * (j & 3)))) * (1.0f / ACT_SL8);
    }
}
IPER_DEVINL void store_planes8(__half* o, int fmt, long long plane_stride, size_t off, const float (&v)[8]) {
    uint32_t hi[4], lo[4];
    float lof[8];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const __half h0 = __float2half_rn(v[2 * j]), h1 = __float2half_rn(v[2 * j + 1]);
        lof[2 * j] = v[2 * j] - __half2float(h0); lof[2 * j + 1] = v[2 * j + 1] - __half2float(h1);
        hi[j] = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
        lo[j] = (uint32_t)__half_as_ushort(__float2half_rn(lof[2 * j])) |
                ((uint32_t)__half_as_ushort(__float2half_rn(lof[2 * j + 1])) << 16);
    }
    *reinterpret_cast<uint4*>(o + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    if (fmt == FMT_HL) {
        *reinterpret_cast<uint4*>(o + plane_stride + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    } else if (fmt == FMT_H8) {
        uint32_t a[2] = {0, 0}, l[2] = {0, 0};
#pragma unroll
        for (int j = 0; j < 8; j++) {
            a[j >> 2] |= (uint32_t)to_e4m3(v[j] * ACT_S8) << (8 * (j & 3));
            l[j >> 2] |= (uint32_t)to_e4m3(lof[j] * ACT_SL8) << (8 * (j & 3));
        }
        uint8_t* a8 = reinterpret_cast<uint8_t*>(o + plane_stride);
        *reinterpret_cast<uint2*>(a8 + off) = make_uint2(a[0], a[1]);
        *reinterpret_cast<uint2*>(a8 + plane_stride + off) = make_uint2(l[0], l[1]);
    }
}

// Warp transpose-reduce: every lane holds v[0..31] (32 channels of ITS pixel); afterwards lane c holds, in v[0], the sum
// over the 32 lanes (pixels) of channel c.  31 shuffles instead of 32 x 5.
IPER_DEVINL float warp_transpose_sum32(float (&v)[32], int lane) {
#pragma unroll
    for (int half = 16; half >= 1; half >>= 1) {
        const bool up = (lane & half) != 0;
#pragma unroll
        for (int j = 0; j < half; j++) {
            const float keep = up ? v[j + half] : v[j];
            const float send = up ? v[j] : v[j + half];
            v[j] = keep + __shfl_xor_sync(0xffffffffu, send, half);
        }
    }
    return v[0];
}

// split an fp32 value into fp16 hi + fp16 lo (hi + lo carries ~22 significand bits)
IPER_DEVINL void split_half(float v, __half& hi, __half& lo) {
    hi = __float2half_rn(v);
    lo = __float2half_rn(v - __half2float(hi));
}
IPER_DEVINL uint32_t pack_half2(__half a, __half b) {
    return (uint32_t)__half_as_ushort(a) | ((uint32_t)__half_as_ushort(b) << 16);
}

}  // namespace iper
