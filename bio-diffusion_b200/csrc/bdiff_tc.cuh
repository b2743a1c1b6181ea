// bdiff_tc.cuh — tcgen05 / TMEM / UMMA-descriptor helpers for the tensor-core edge pass (sm_100a).
//
// Conventions used by every UMMA operand in this library (bf16, K-major, 128-byte swizzle):
//   an operand "K-block" is [rows][64 bf16] = rows x 128 B, 1024-byte aligned; element (r, k) lives at
//       r*128 + (((k >> 3) ^ (r & 7)) << 4) + (k & 7)*2            (the TMA SWIZZLE_128B pattern)
//   the shared-memory descriptor points at the block (plus 32 B per K=16 step, plus r0*128 for a row offset
//   that is a multiple of 8), SBO = 1024 B (8 rows), LBO unused, version 1, layout SWIZZLE_128B.
//   Accumulators: M=128 rows -> TMEM lanes 0..127, N columns -> consecutive 32-bit TMEM columns.
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>

#include "bdiff_common.cuh"

namespace bdiff {

__device__ __forceinline__ uint32_t sw128_offset(int r, int k) {   // byte offset inside a K-block
  return (uint32_t)(r * 128 + ((((k >> 3) ^ (r & 7)) & 7) << 4) + (k & 7) * 2);
}

// 64-bit UMMA shared-memory descriptor for a K-major SWIZZLE_128B operand at byte address `saddr`.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);          // start address  [0,14)
  d |= (uint64_t)0 << 16;                          // leading byte offset [16,30): unused (one atom along K)
  d |= (uint64_t)((1024 >> 4) & 0x3FFF) << 32;     // stride byte offset  [32,46): 8 rows * 128 B
  d |= (uint64_t)1 << 46;                          // descriptor version  [46,48) = 1 on sm_100
  d |= (uint64_t)2 << 61;                          // layout type [61,64): SWIZZLE_128B
  return d;
}

// 32-bit instruction descriptor, kind::f16: D=f32, A=B=bf16, both K-major, M=128, N given.
__device__ __forceinline__ uint32_t umma_idesc_bf16(int n, bool negate_a) {
  uint32_t d = 0;
  d |= 1u << 4;                    // c_format = F32
  d |= 1u << 7;                    // a_format = BF16
  d |= 1u << 10;                   // b_format = BF16
  d |= (negate_a ? 1u : 0u) << 13; // a_negate
  d |= (uint32_t)(n >> 3) << 17;   // n_dim
  d |= (uint32_t)(128 >> 4) << 24; // m_dim
  return d;
}

__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          bool accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate ? 1u : 0u)
      : "memory");
}

// All previously issued MMAs of this thread arrive (once) on `bar` when they complete.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {   // one full warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {     // same warp that allocated
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}


// ------------------------------------------------------------------------------------------ CTA pairs (cta_group::2)
// A thread-block cluster of two CTAs can issue one MMA over both SMs: M = 256 (128 rows per CTA, each CTA's own A tile and
// TMEM), while the N x K B operand is SPLIT — CTA 0's shared memory holds rows [0, N/2), CTA 1's rows [N/2, N) at the same
// offset — so every SM reads and receives only half of each weight plane.  Issued by one thread of the leader CTA (rank 0).
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `p` (a shared::cta pointer of this CTA) in the CTA of rank `cta`
__device__ __forceinline__ uint32_t mapa_u32(const void* p, uint32_t cta) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(p)), "r"(cta));
  return r;
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// (a plain cp.async.bulk whose mbarrier lives in the OTHER CTA of the pair never completes — tried; only the tensor-map TMA
// forms have the cta_group::2 variant that signals the leader's barrier — hence the relay lanes of the layer megakernel)
// same without the release: for pure event forwarding (nothing this thread wrote has to become visible with the arrival)
__device__ __forceinline__ void mbar_arrive_remote_relaxed(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* smem_dst, uint32_t ncols) {   // one full warp in EACH CTA of the pair
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// instruction descriptor for the pair MMA: M = 256
__device__ __forceinline__ uint32_t umma_idesc_bf16_m256(int n) {
  uint32_t d = 0;
  d |= 1u << 4;
  d |= 1u << 7;
  d |= 1u << 10;
  d |= (uint32_t)(n >> 3) << 17;
  d |= (uint32_t)(256 >> 4) << 24;
  return d;
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, bool accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate ? 1u : 0u)
      : "memory");
}
// all previously issued pair MMAs of this thread arrive on the mbarrier at this offset in BOTH CTAs (mask 0b11)
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"((unsigned short)3)
               : "memory");
}

// TMEM -> registers: this thread's lane (32*(warp%4) + laneid), `N` consecutive 32-bit columns from taddr.
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* v) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,"
      "%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
// Asynchronous variants: issue several loads / stores back to back, then ONE wait (tcgen05.wait covers all
// previously issued operations of the thread), instead of paying the TMEM round trip once per instruction.
__device__ __forceinline__ void tmem_ld8_nw(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld32_nw(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,"
      "%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st8_nw(uint32_t taddr, const float* v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr),
               "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
               "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])),
               "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7]))
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// N x 8 columns in one go
template <int N8>
__device__ __forceinline__ void tmem_ld8xN(uint32_t taddr, float* v) {
  uint32_t r[N8 * 8];
#pragma unroll
  for (int q = 0; q < N8; ++q) tmem_ld8_nw(taddr + q * 8, r + q * 8);
  tmem_ld_wait();
#pragma unroll
  for (int i = 0; i < N8 * 8; ++i) v[i] = __uint_as_float(r[i]);
}
template <int N8>
__device__ __forceinline__ void tmem_st8xN(uint32_t taddr, const float* v) {   // caller waits (tmem_st_wait) later
#pragma unroll
  for (int q = 0; q < N8; ++q) tmem_st8_nw(taddr + q * 8, v + q * 8);
}
// two 32-column chunks with one wait
__device__ __forceinline__ void tmem_ld64(uint32_t taddr, float* v) {
  uint32_t r[64];
  tmem_ld32_nw(taddr, r);
  tmem_ld32_nw(taddr + 32, r + 32);
  tmem_ld_wait();
#pragma unroll
  for (int i = 0; i < 64; ++i) v[i] = __uint_as_float(r[i]);
}

// registers -> TMEM (per-thread scratch columns)
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float* v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr),
               "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
               "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])),
               "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7]))
               : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// fast activations for the tensor path (one MUFU each; operands are rounded to bf16 anyway)
__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float sigmoid_fast(float x) { return fmaf(0.5f, tanh_fast(0.5f * x), 0.5f); }
__device__ __forceinline__ float silu_fast(float x) { return x * sigmoid_fast(x); }

// packed fp32x2 versions (FFMA2 / FMUL2 / FADD2 on sm_100a: two lanes per issued instruction)
__device__ __forceinline__ float2 sigmoid_fast2(float2 x) {
  const float2 hx = __fmul2_rn(x, make_float2(0.5f, 0.5f));
  const float2 t = make_float2(tanh_fast(hx.x), tanh_fast(hx.y));
  return __ffma2_rn(t, make_float2(0.5f, 0.5f), make_float2(0.5f, 0.5f));
}
__device__ __forceinline__ float2 silu_fast2(float2 x) { return __fmul2_rn(x, sigmoid_fast2(x)); }

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t u) {
  __nv_bfloat162 h = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(h);
}


// 64-bit UMMA shared-memory descriptor for a K-major operand WITHOUT swizzle ("interleaved" canonical layout,
// cute UMMA::LayoutType::SWIZZLE_NONE): core matrices of 8 rows x 16 bytes are contiguous (128 B); consecutive
// 8-row groups are `sbo` bytes apart, the two 16-byte K chunks of one K=16 step are `lbo` bytes apart.
__device__ __forceinline__ uint64_t umma_desc_k16(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}

// ------------------------------------------------------------------------------------- split-bf16 operands
// Every GEMM operand of the tensor path is the pair (hi, lo) of bf16 numbers with hi = RN_bf16(value)
// and lo = RN_bf16(value - hi): hi + lo carries >= 16 significant bits (relative error
// <= 2^-18) and is exactly representable in fp32.  A product is evaluated as A_hi.W_hi + A_lo.W_hi + A_hi.W_lo
// (three tcgen05 MMAs accumulating into the same fp32 TMEM columns; the dropped A_lo.W_lo term is 2^-18 relative).
__device__ __forceinline__ void split_bf16x2(float a, float b, uint32_t& hi, uint32_t& lo) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  hi = *reinterpret_cast<uint32_t*>(&h);
  const float ra = a - __uint_as_float(hi << 16), rb = b - __uint_as_float(hi & 0xffff0000u);
  __nv_bfloat162 l = __floats2bfloat162_rn(ra, rb);
  lo = *reinterpret_cast<uint32_t*>(&l);
}
__device__ __forceinline__ float2 join_bf16x2(uint32_t hi, uint32_t lo) {
  return make_float2(__uint_as_float(hi << 16) + __uint_as_float(lo << 16),
                     __uint_as_float(hi & 0xffff0000u) + __uint_as_float(lo & 0xffff0000u));
}
__device__ __forceinline__ void split_bf16(float a, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(a);
  lo = __float2bfloat16_rn(a - __bfloat162float(hi));
}

// fp32-class activations for the tensor path: ex2.approx / rcp.approx are accurate to ~2 ulp (the former
// tanh.approx form was good to 2^-11 only).  sigmoid(x) = 1 / (1 + 2^(-x log2 e)).
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float sigmoid_acc(float x) { return rcp_approx(1.0f + ex2_approx(-1.4426950408889634f * x)); }
__device__ __forceinline__ float silu_acc(float x) { return x * sigmoid_acc(x); }
__device__ __forceinline__ float2 sigmoid_acc2(float2 x) {
  const float2 t = __fmul2_rn(x, make_float2(-1.4426950408889634f, -1.4426950408889634f));
  const float2 d = __fadd2_rn(make_float2(ex2_approx(t.x), ex2_approx(t.y)), make_float2(1.0f, 1.0f));
  return make_float2(rcp_approx(d.x), rcp_approx(d.y));
}
__device__ __forceinline__ float2 silu_acc2(float2 x) { return __fmul2_rn(x, sigmoid_acc2(x)); }

// A-operand tile: 128 rows x 64 bf16 per K-block (16 KiB), K-blocks consecutive.
constexpr int X_BLOCK = 128 * 128;
__device__ __forceinline__ void x_store8(unsigned char* X, int r, int kk, const float* v) {   // kk % 8 == 0
  *reinterpret_cast<uint4*>(X + (kk >> 6) * X_BLOCK + sw128_offset(r, kk & 63)) =
      make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
}
__device__ __forceinline__ void x_store1(unsigned char* X, int r, int kk, float v) {
  *reinterpret_cast<__nv_bfloat16*>(X + (kk >> 6) * X_BLOCK + sw128_offset(r, kk & 63)) = __float2bfloat16_rn(v);
}
__device__ __forceinline__ void x_load8(const unsigned char* X, int r, int kk, float* v) {
  const uint4 u = *reinterpret_cast<const uint4*>(X + (kk >> 6) * X_BLOCK + sw128_offset(r, kk & 63));
  float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), e = unpack_bf16x2(u.w);
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = e.x; v[7] = e.y;
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

}  // namespace bdiff
