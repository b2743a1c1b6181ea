// bdiff_edge_tc.cu — tensor-core (tcgen05 / TMEM / TMA-bulk) version of the fused edge pass.
//
// Same math as k_edge_message (gcpnet.py:676-737: 4 residual GCP2s on the per-edge message, attention gate,
// segmented row-sum) but the dense scalar GEMMs run on the 5th-gen tensor cores:
//   * persistent CTAs (one per SM), tile = 128 consecutive edges = the 128 TMEM lanes (thread t <-> edge t);
//   * A operand = the running scalar message m.s as bf16 in shared memory (K-major, 128B-swizzled), written by
//     the epilogue warps; B operand = pre-swizzled bf16 weight K-blocks streamed from L2 by TMA bulk copies
//     through a 2-stage mbarrier ring; fp32 accumulators in TMEM;
//   * the vector gate  sigmoid(Wg_k silu(S_k) + b)  is folded into the surrounding GEMMs as extra N=32 column
//     groups using  Wg_k new_k = Wg_k m_k - Wg_k m_{k-1}  (second term issued with the A-negate bit), so the
//     only A operand ever needed is m.s;
//   * the equivariant vector channel (32x3 per edge) lives in TMEM scratch columns of the owning lane and is
//     updated with thread-local FMAs (weights broadcast from shared memory);
//   * warp roles: warps 0-7 epilogue/compute (a thread PAIR per edge), warp 8 TMA producer (+TMEM allocator),
//     warp 9 MMA issuer.
#include "bdiff_kernels.h"
#include "bdiff_tc.cuh"

namespace bdiff {

#ifndef BDIFF_STAMP
#define BDIFF_STAMP(slot) do { if (w.dbg && (slot) < 64) w.dbg[(size_t)blockIdx.x * 64 + (slot)] = clock64(); } while (0)
#endif

constexpr int TC_THREADS = 192;
constexpr int TMT = 128;                 // edges per tile
constexpr int RING_STAGE = 320 * 128;    // bytes of the largest weight chunk (320 rows x 64 bf16)
// TMEM column map (512 columns allocated)
constexpr int TM_S = 0, TM_U0 = 256, TM_U1 = 288, TM_MV = 320, TM_VD0 = 416;

__host__ __device__ inline int tc_nc0(int Ed, int Xd) {          // weight chunks of message GCP 0
  const int k0raw = Ed + (64 + Xd) / 4 + 9;
  return ((k0raw + 15) / 16 + 3) / 4;
}
size_t tc_blob_bytes(int Ed, int Xd) {
  return (size_t)tc_nc0(Ed, Xd) * 256 * 128 + 3 * (4 * (size_t)RING_STAGE + 256 * 128) + 4 * 32 * 128;
}

// ---------------------------------------------------------------------------------------------- weight blob
// Builds the per-layer bf16 blob: a sequence of K-blocks [rows][64] in the 128B-swizzled shared-memory image,
// in the exact order the producer streams them:  G0 chunks | for k=1..3: 4 x [W_k | Wg_{k-1} | Wg_k] , W_k tail |
// 4 x Wg_3.
__global__ void k_tc_pack_layer(LayerW lw, Dims d, unsigned char* __restrict__ blob) {
  const int nc0 = tc_nc0(d.Ed, d.Xd);
  const long long total_rows = (long long)nc0 * 256 + 3 * (4 * 320 + 256) + 4 * 32;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total_rows * 64) return;
  long long rowg = idx / 64;
  const int kc = (int)(idx - rowg * 64);
  size_t base = 0;
  float v = 0.f;
  int r = 0;
  if (rowg < (long long)nc0 * 256) {
    const int j = (int)(rowg / 256);
    r = (int)(rowg - (long long)j * 256);
    base = (size_t)j * 256 * 128;
    const int kk = j * 64 + kc;
    v = kk < d.K0 ? lw.W0e[(size_t)kk * 256 + r] : 0.f;
  } else {
    rowg -= (long long)nc0 * 256;
    base = (size_t)nc0 * 256 * 128;
    const long long per_k = 4 * 320 + 256;
    if (rowg < 3 * per_k) {
      const int k = (int)(rowg / per_k);           // message GCP k+1
      long long rr = rowg - k * per_k;
      base += (size_t)k * per_k * 128;
      if (rr < 4 * 320) {
        const int j = (int)(rr / 320);
        r = (int)(rr - j * 320);
        base += (size_t)j * 320 * 128;
        const int kk = j * 64 + kc;
        if (r < 256) v = lw.Wk[k][(size_t)kk * 256 + r];
        else if (r < 288) v = (k == 0 ? lw.Wg0 : lw.Wgk[k - 1])[(size_t)kk * 32 + (r - 256)];
        else v = lw.Wgk[k][(size_t)kk * 32 + (r - 288)];
      } else {
        r = (int)(rr - 4 * 320);
        base += (size_t)4 * 320 * 128;
        const int kk = 256 + kc;
        v = kk < kKM ? lw.Wk[k][(size_t)kk * 256 + r] : 0.f;
      }
    } else {
      long long rr = rowg - 3 * per_k;
      base += (size_t)3 * per_k * 128;
      const int j = (int)(rr / 32);
      r = (int)(rr - j * 32);
      base += (size_t)j * 32 * 128;
      v = lw.Wgk[2][(size_t)(j * 64 + kc) * 32 + r];
    }
  }
  *reinterpret_cast<__nv_bfloat16*>(blob + base + sw128_offset(r, kc)) = __float2bfloat16_rn(v);
}

// ------------------------------------------------------------------------------------------------ self test
// C[128][320] = A[128][128] . W[320][128]^T with the exact machinery of the edge kernel: swizzled bf16 A written
// by threads, W image fetched by a TMA bulk copy, three MMAs per K step (N=256 | N=32 | N=32 with A negated),
// TMEM scratch round trip in column 320.  Used by tests/test_gpu_tc.py before the fused kernel is trusted.
__global__ void __launch_bounds__(TC_THREADS, 1) k_umma_selftest(const float* __restrict__ A,
                                                                 const unsigned char* __restrict__ wimg,
                                                                 float* __restrict__ C) {
  extern __shared__ unsigned char smem_raw[];
  // align to 1024 B by OFFSETTING the __shared__ array (keeps the shared address space -> LDS/STS, not generic LD/ST)
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char* X = smem;                          // 2 K-blocks of A
  unsigned char* Wb = smem + 2 * X_BLOCK;           // 2 K-blocks of W (320 rows each)
  uint64_t* bars = reinterpret_cast<uint64_t*>(Wb + 2 * RING_STAGE);
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 4);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    mbar_init(&bars[0], 1);      // weights landed
    mbar_init(&bars[1], 1);      // MMAs done
    mbar_fence_init();
  }
  if (warp == 4) tmem_alloc(tmem_ptr, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr;
  if (tid < TMT) {
    for (int k8 = 0; k8 < 16; ++k8) {
      uint32_t pk[4];
#pragma unroll
      for (int q = 0; q < 4; ++q)
        pk[q] = pack_bf16x2(A[(size_t)tid * 128 + k8 * 8 + 2 * q], A[(size_t)tid * 128 + k8 * 8 + 2 * q + 1]);
      const int kk = k8 * 8;
      *reinterpret_cast<uint4*>(X + (kk / 64) * X_BLOCK + sw128_offset(tid, kk % 64)) =
          make_uint4(pk[0], pk[1], pk[2], pk[3]);
    }
    fence_proxy_async();
  }
  if (tid == TMT) {   // warp 4 lane 0: fetch both weight K-blocks
    mbar_expect_tx(&bars[0], 2 * RING_STAGE);
    bulk_g2s(Wb, wimg, RING_STAGE, &bars[0]);
    bulk_g2s(Wb + RING_STAGE, wimg + RING_STAGE, RING_STAGE, &bars[0]);
  }
  __syncthreads();
  if (tid == TMT + 32) {   // warp 5 lane 0: MMA issuer
    mbar_wait(&bars[0], 0);
    tc_fence_after();
    const uint32_t i256 = umma_idesc_bf16(256, false), i32 = umma_idesc_bf16(32, false),
                   i32n = umma_idesc_bf16(32, true);
    for (int j = 0; j < 2; ++j)
      for (int s = 0; s < 4; ++s) {
        const uint64_t ad = umma_desc_sw128(smem_u32(X + j * X_BLOCK) + s * 32);
        const uint32_t wb = smem_u32(Wb + j * RING_STAGE) + s * 32;
        const bool acc = (j | s) > 0;
        umma_bf16(tmem + TM_S, ad, umma_desc_sw128(wb), i256, acc);
        umma_bf16(tmem + TM_U0, ad, umma_desc_sw128(wb + 256 * 128), i32, acc);
        umma_bf16(tmem + TM_U1, ad, umma_desc_sw128(wb + 288 * 128), i32n, acc);
      }
    umma_commit(&bars[1]);
  }
  if (tid < TMT) {
    mbar_wait(&bars[1], 0);
    tc_fence_after();
    const uint32_t tl = tmem + ((uint32_t)(warp * 32) << 16);
    float sc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) sc[i] = (float)(tid * 8 + i);
    tmem_st8(tl + TM_MV, sc);
    for (int c0 = 0; c0 < 320; c0 += 32) {
      float v[32];
      tmem_ld32(tl + c0, v);
#pragma unroll
      for (int i = 0; i < 32; ++i) C[(size_t)tid * 328 + c0 + i] = v[i];
    }
    float back[8];
    tmem_ld8(tl + TM_MV, back);
#pragma unroll
    for (int i = 0; i < 8; ++i) C[(size_t)tid * 328 + 320 + i] = back[i];
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem, 512);
  (void)lane;
}

// [320][128] fp32 -> two swizzled bf16 K-blocks of 320 rows (the self test's weight image)
__global__ void k_selftest_pack(const float* __restrict__ W, unsigned char* __restrict__ img) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= 320 * 128) return;
  const int r = idx / 128, k = idx - r * 128;
  *reinterpret_cast<__nv_bfloat16*>(img + (k / 64) * RING_STAGE + sw128_offset(r, k % 64)) =
      __float2bfloat16_rn(W[idx]);
}

// --------------------------------------------------------------------------------------------- fused kernel
// Thread roles: warps 0-7 epilogue/compute — edge r of the tile is owned by the thread PAIR (r, r+128): "half" 0
// works on accumulator columns [0,128) and vector channels [0,16), half 1 on columns [128,256) and channels
// [16,32) (both warps of a pair address the same TMEM lanes: lane quarter = warp % 4); warp 8 = TMA producer
// (+ TMEM allocator), warp 9 = MMA issuer.
constexpr int TC_EPI = 256;
constexpr int TC_THREADS2 = TC_EPI + 64;

struct alignas(16) SmallW {   // fp32 copies of the thread-local (vector channel) weights, broadcast-read
  float Wd0x[16 * 20];     // [Xd][hid0]
  float Wf0x[16 * 3];      // [Xd][3]
  float Wu0[20 * 32];      // [hid0][32]
  float Wdk[3][32 * 8];    // [32][8]
  float Wfk[3][32 * 3];    // [32][3]
  float Wuk[3][8 * 32];    // [8][32]
  float bg[4][32];
  float bk[3][256];
  float wa[256];
  float ba[4];
};

constexpr int ST_LD = 37;
struct TcSmemTail {
  float sT[2][TMT][ST_LD]; // per-half transpose buffer of the final reduction; reused as the pair-exchange buffer and as
                           // the staging area of the coalesced xi / P_j gathers
  SmallW sw;
  float sAttn[2][TMT];
  int sRow[TMT], sCol[TMT], sB[TMT], sNa[TMT];
  uint64_t full[2], empty[2], a_ready, d_full, wbar;
  uint32_t tmem_ptr;
};

constexpr size_t TC_SMEM_BYTES = 5 * (size_t)X_BLOCK + 2 * (size_t)RING_STAGE + sizeof(TcSmemTail) + 1024;

// Gate of the previous GCP from TMEM (U), vector-message update in TMEM scratch for this thread's 16 channels,
// and this thread's partial vector_down / vector_down_frames sums of the NEXT GCP.
// HP = hidden dim of the previous GCP; vdp = its vector_down output (full, [HP][3]).
template <int HP, bool FIRST, bool LAST>
__device__ __forceinline__ void gate_update(uint32_t tl, int half, int ucol, const float* __restrict__ vdp,
                                            const float* __restrict__ Wu, const float* __restrict__ bgp,
                                            const float* __restrict__ Wdn, const float* __restrict__ Wfn,
                                            float* __restrict__ part) {   // part[33]: partial VD_next(24)+VDF_next(9)
  float2 p2[4][3];                 // VD_next accumulators, pairs of hidden rows (h = 2hp, 2hp+1) per component
  if (!LAST) {
#pragma unroll
    for (int i = 0; i < 33; ++i) part[i] = 0.f;
#pragma unroll
    for (int hp = 0; hp < 4; ++hp) { p2[hp][0] = make_float2(0.f, 0.f); p2[hp][1] = p2[hp][0]; p2[hp][2] = p2[hp][0]; }
  }
  for (int oc = half * 2; oc < half * 2 + 2; ++oc) {
    float u[8], mv[24];
    {
      uint32_t ru[8], rm[24];
      tmem_ld8_nw(tl + ucol + oc * 8, ru);
      if (!FIRST) {
        tmem_ld8_nw(tl + TM_MV + oc * 24, rm);
        tmem_ld8_nw(tl + TM_MV + oc * 24 + 8, rm + 8);
        tmem_ld8_nw(tl + TM_MV + oc * 24 + 16, rm + 16);
      }
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 8; ++i) u[i] = __uint_as_float(ru[i]);
      if (!FIRST) {
#pragma unroll
        for (int i = 0; i < 24; ++i) mv[i] = __uint_as_float(rm[i]);
      }
    }
#pragma unroll
    for (int jp = 0; jp < 4; ++jp) {          // two output channels (j = 2jp, 2jp+1) per packed instruction
      const int o = oc * 8 + 2 * jp;
      const float2 g = sigmoid_fast2(__fadd2_rn(make_float2(u[2 * jp], u[2 * jp + 1]),
                                                *reinterpret_cast<const float2*>(bgp + o)));
      float2 s0 = make_float2(0.f, 0.f), s1 = s0, s2 = s0;
#pragma unroll
      for (int h = 0; h < HP; ++h) {
        const float2 wu = *reinterpret_cast<const float2*>(Wu + h * 32 + o);
        s0 = __ffma2_rn(wu, make_float2(vdp[h * 3 + 0], vdp[h * 3 + 0]), s0);
        s1 = __ffma2_rn(wu, make_float2(vdp[h * 3 + 1], vdp[h * 3 + 1]), s1);
        s2 = __ffma2_rn(wu, make_float2(vdp[h * 3 + 2], vdp[h * 3 + 2]), s2);
      }
      const int ja = 2 * jp * 3, jb = (2 * jp + 1) * 3;
      float2 r0, r1, r2;
      if (FIRST) { r0 = __fmul2_rn(s0, g); r1 = __fmul2_rn(s1, g); r2 = __fmul2_rn(s2, g); }
      else {
        r0 = __ffma2_rn(s0, g, make_float2(mv[ja + 0], mv[jb + 0]));
        r1 = __ffma2_rn(s1, g, make_float2(mv[ja + 1], mv[jb + 1]));
        r2 = __ffma2_rn(s2, g, make_float2(mv[ja + 2], mv[jb + 2]));
      }
      mv[ja + 0] = r0.x; mv[jb + 0] = r0.y; mv[ja + 1] = r1.x; mv[jb + 1] = r1.y; mv[ja + 2] = r2.x; mv[jb + 2] = r2.y;
    }
    tmem_st8xN<3>(tl + TM_MV + oc * 24, mv);      // completion awaited once, at the end of the function
    if (!LAST) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int c = oc * 8 + j;
        const float4 wd0 = *reinterpret_cast<const float4*>(Wdn + c * 8), wd1 = *reinterpret_cast<const float4*>(Wdn + c * 8 + 4);
        const float2 wdp[4] = {make_float2(wd0.x, wd0.y), make_float2(wd0.z, wd0.w), make_float2(wd1.x, wd1.y),
                               make_float2(wd1.z, wd1.w)};
#pragma unroll
        for (int x = 0; x < 3; ++x) {
          const float2 mb = make_float2(mv[j * 3 + x], mv[j * 3 + x]);
#pragma unroll
          for (int hp = 0; hp < 4; ++hp) p2[hp][x] = __ffma2_rn(wdp[hp], mb, p2[hp][x]);
        }
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
          const float wf = Wfn[c * 3 + ch];
          part[24 + ch * 3 + 0] = fmaf(wf, mv[j * 3 + 0], part[24 + ch * 3 + 0]);
          part[24 + ch * 3 + 1] = fmaf(wf, mv[j * 3 + 1], part[24 + ch * 3 + 1]);
          part[24 + ch * 3 + 2] = fmaf(wf, mv[j * 3 + 2], part[24 + ch * 3 + 2]);
        }
      }
    }
  }
  if (!LAST) {
#pragma unroll
    for (int hp = 0; hp < 4; ++hp)
#pragma unroll
      for (int x = 0; x < 3; ++x) { part[(2 * hp) * 3 + x] = p2[hp][x].x; part[(2 * hp + 1) * 3 + x] = p2[hp][x].y; }
  }
  tmem_st_wait();
}

template <int ED, int XD>
__global__ void __launch_bounds__(TC_THREADS2, 1)
    k_edge_message_tc(Plan p, LayerW lw, const unsigned char* __restrict__ blob, Work w, int ntiles) {
  constexpr int HID0 = (64 + XD) / 4;
  constexpr int H2 = HID0 / 2;             // vector_down rows of GCP 0 computed per half
  constexpr int K0RAW = ED + HID0 + 9;
  constexpr int K0S = (K0RAW + 15) / 16;
  constexpr int NC0 = (K0S + 3) / 4;
  static_assert(HID0 % 2 == 0 && H2 * 3 <= 32 && HID0 + 9 <= 32 && ED % 16 == 0, "layout");

  extern __shared__ unsigned char smem_raw[];
  // align to 1024 B by OFFSETTING the __shared__ array (keeps the shared address space -> LDS/STS, not generic LD/ST)
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char* X = smem;
  unsigned char* ring = smem + 5 * X_BLOCK;
  TcSmemTail& T = *reinterpret_cast<TcSmemTail*>(ring + 2 * RING_STAGE);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    mbar_init(&T.full[0], 1); mbar_init(&T.full[1], 1);
    mbar_init(&T.empty[0], 1); mbar_init(&T.empty[1], 1);
    mbar_init(&T.a_ready, TC_EPI);
    mbar_init(&T.d_full, 1);
    mbar_fence_init();
  }
  if (warp == 8) tmem_alloc(&T.tmem_ptr, 512);
  // small (vector-channel) weights -> shared memory: one TMA bulk copy per array, all in flight at once
  // (every packed array starts 256-byte aligned and is padded, so sizes are rounded up to 16 bytes)
  if (tid == 0) {
    mbar_init(&T.wbar, 1);
    mbar_fence_init();
    SmallW& s = T.sw;
    auto sz = [](int n) { return (uint32_t)((n * 4 + 15) & ~15); };
    mbar_expect_tx(&T.wbar, sz(XD * HID0) + sz(XD * 3) + sz(HID0 * 32) +
                                3 * (sz(256) + sz(256) + sz(256) + sz(96) + sz(32)) + sz(32) + sz(256) + sz(1));
    auto cp = [&](float* dst, const float* src, int n) { bulk_g2s(dst, src, sz(n), &T.wbar); };
    cp(s.Wd0x, lw.Wd0x, XD * HID0); cp(s.Wf0x, lw.Wf0x, XD * 3); cp(s.Wu0, lw.Wu0, HID0 * 32);
    for (int k = 0; k < 3; ++k) {
      cp(s.Wdk[k], lw.Wdk[k], 256); cp(s.Wuk[k], lw.Wuk[k], 256); cp(s.bk[k], lw.bk[k], 256);
      cp(s.Wfk[k], lw.Wfk[k], 96); cp(s.bg[k + 1], lw.bgk[k], 32);
    }
    cp(s.bg[0], lw.bg0, 32); cp(s.wa, lw.wa, 256); cp(s.ba, lw.ba, 1);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = T.tmem_ptr;
  if (warp < 8) mbar_wait(&T.wbar, 0);      // small weights have landed (only the compute warps read them)

  if (warp == 8) {
    // ===================================================================== TMA producer (one lane)
    if (lane == 0) {
      uint32_t ci = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        size_t off = 0;
        auto push = [&](uint32_t bytes) {
          const uint32_t s = ci & 1;
          mbar_wait_backoff(&T.empty[s], ((ci >> 1) & 1) ^ 1);
          mbar_expect_tx(&T.full[s], bytes);
          bulk_g2s(ring + s * RING_STAGE, blob + off, bytes, &T.full[s]);
          off += bytes;
          ++ci;
        };
        for (int j = 0; j < NC0; ++j) push(256 * 128);
        for (int k = 0; k < 3; ++k) {
          for (int j = 0; j < 4; ++j) push(320 * 128);
          push(256 * 128);
        }
        for (int j = 0; j < 4; ++j) push(32 * 128);
      }
    }
  } else if (warp == 9) {
    // ======================================================================= MMA issuer (one lane)
    if (lane == 0) {
      const uint32_t i256 = umma_idesc_bf16(256, false), i32 = umma_idesc_bf16(32, false),
                     i32n = umma_idesc_bf16(32, true);
      const uint32_t xaddr = smem_u32(X), raddr = smem_u32(ring);
      uint32_t ci = 0, pa = 0;
      auto wait_a = [&]() { mbar_wait_backoff(&T.a_ready, pa); pa ^= 1; tc_fence_after(); };
      auto wait_w = [&]() -> uint32_t {
        const uint32_t s = ci & 1;
        mbar_wait_backoff(&T.full[s], (ci >> 1) & 1);
        tc_fence_after();
        return raddr + s * RING_STAGE;
      };
      auto done_w = [&]() { umma_commit(&T.empty[ci & 1]); ++ci; };
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        // ---- G0: S = [e | vn0 | q0] . W0e^T
        wait_a();
        for (int j = 0; j < NC0; ++j) {
          const uint32_t wb = wait_w();
          const int ns = min(4, K0S - 4 * j);
          for (int s = 0; s < ns; ++s)
            umma_bf16(tmem + TM_S, umma_desc_sw128(xaddr + j * X_BLOCK + s * 32), umma_desc_sw128(wb + s * 32), i256,
                      (j | s) > 0);
          done_w();
        }
        umma_commit(&T.d_full);
        for (int k = 1; k <= 3; ++k) {
          // ---- G(k)a: S = m_{k-1} . W_k[:, :256]^T ;  U[(k-1)&1] += Wg_{k-1} m_{k-1} ;  U[k&1] = -Wg_k m_{k-1}
          wait_a();
          const uint32_t up = tmem + (((k - 1) & 1) ? TM_U1 : TM_U0), un = tmem + ((k & 1) ? TM_U1 : TM_U0);
          for (int j = 0; j < 4; ++j) {
            const uint32_t wb = wait_w();
            for (int s = 0; s < 4; ++s) {
              const uint64_t ad = umma_desc_sw128(xaddr + j * X_BLOCK + s * 32);
              const bool acc = (j | s) > 0;
              umma_bf16(tmem + TM_S, ad, umma_desc_sw128(wb + s * 32), i256, acc);
              umma_bf16(up, ad, umma_desc_sw128(wb + 256 * 128 + s * 32), i32, k == 1 ? acc : true);
              umma_bf16(un, ad, umma_desc_sw128(wb + 288 * 128 + s * 32), i32n, acc);
            }
            done_w();
          }
          umma_commit(&T.d_full);
          // ---- G(k)b: S += [vn_k | q_k] . W_k[:, 256:288]^T
          wait_a();
          {
            const uint32_t wb = wait_w();
            for (int s = 0; s < 2; ++s)
              umma_bf16(tmem + TM_S, umma_desc_sw128(xaddr + 4 * X_BLOCK + s * 32), umma_desc_sw128(wb + s * 32), i256,
                        true);
            done_w();
          }
          umma_commit(&T.d_full);
        }
        // ---- G4: U1 += Wg_3 m_3
        wait_a();
        for (int j = 0; j < 4; ++j) {
          const uint32_t wb = wait_w();
          for (int s = 0; s < 4; ++s)
            umma_bf16(tmem + TM_U1, umma_desc_sw128(xaddr + j * X_BLOCK + s * 32), umma_desc_sw128(wb + s * 32), i32,
                      true);
          done_w();
        }
        umma_commit(&T.d_full);
      }
    }
  } else {
    // ====================================================== epilogue / compute warps (thread pair <-> edge r)
    const int half = tid >> 7, r = tid & 127;
    const uint32_t tl = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    const SmallW& sw = T.sw;
    float* exch_mine = &T.sT[half][r][0];
    const float* exch_other = &T.sT[half ^ 1][r][0];
    uint32_t pd = 0;
    int es = 0;
    auto wait_d = [&]() { if (tid == 0) BDIFF_STAMP(es++); mbar_wait(&T.d_full, pd); pd ^= 1; tc_fence_after(); if (tid == 0) BDIFF_STAMP(es++); };
    auto publish = [&]() { fence_proxy_async(); tc_fence_before(); mbar_arrive(&T.a_ready); if (tid == 0) BDIFF_STAMP(es++); };
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      if (tid == 0) BDIFF_STAMP(es++);
      const long long g = (long long)tile * TMT + r;
      // ---- T0: A operand of GCP 0 = [e | vn0 | q0]; VD0 goes to TMEM scratch for the vector_up of GCP 0.
      // Global reads are issued so that (a) one warp instruction touches a few 128-byte lines (the L1 handles one
      // line tag per cycle, so "lane = edge row" gathers cost 32 cycles each) and (b) independent loads are in
      // flight together: the per-edge record, e, xi and the frames first; then, once (row, col) are known, the
      // endpoint vector parts.  e and xi are contiguous per tile and are staged by all threads in row-major order;
      // P_j rows are gathered 8..18 lanes per row through shared memory.
      float* sF = &T.sT[0][0][0];
      constexpr int PV = HID0 * 3 + 9, PF4 = (PV + 3) / 4, PLD = PF4 * 4 + 1;
      constexpr int XF4 = XD * 3 / 4, XLD = XD * 3 + 1;
      static_assert(TMT * PLD <= 2 * TMT * ST_LD && TMT * XLD <= 2 * TMT * ST_LD, "staging area");
      const int4 rc = __ldg(p.edge_rc + g);
      float f[9];
      {
        constexpr int F4_ROW = ED / 4;
        const float4* eb = reinterpret_cast<const float4*>(w.e + (size_t)tile * TMT * ED);
        const float4* xb = reinterpret_cast<const float4*>(w.xi + (size_t)tile * TMT * XD * 3);
        float4 ev[TMT * F4_ROW / TC_EPI], xv[TMT * XF4 / TC_EPI];
#pragma unroll
        for (int i = 0; i < TMT * F4_ROW / TC_EPI; ++i) ev[i] = eb[i * TC_EPI + tid];
#pragma unroll
        for (int i = 0; i < TMT * XF4 / TC_EPI; ++i) xv[i] = xb[i * TC_EPI + tid];
#pragma unroll
        for (int q = 0; q < 9; ++q) f[q] = w.frames[(size_t)g * 9 + q];
        if (half == 0) { T.sRow[r] = rc.x; T.sCol[r] = rc.y; T.sB[r] = rc.z; T.sNa[r] = rc.w; }
#pragma unroll
        for (int i = 0; i < TMT * F4_ROW / TC_EPI; ++i) {
          const int idx = i * TC_EPI + tid;
          const int rr = idx / F4_ROW, kk = (idx % F4_ROW) * 4;
          *reinterpret_cast<uint2*>(X + (kk >> 6) * X_BLOCK + sw128_offset(rr, kk & 63)) =
              make_uint2(pack_bf16x2(ev[i].x, ev[i].y), pack_bf16x2(ev[i].z, ev[i].w));
        }
#pragma unroll
        for (int i = 0; i < TMT * XF4 / TC_EPI; ++i) {
          const int idx = i * TC_EPI + tid;
          float* dst = sF + (idx / XF4) * XLD + (idx % XF4) * 4;
          dst[0] = xv[i].x; dst[1] = xv[i].y; dst[2] = xv[i].z; dst[3] = xv[i].w;
        }
      }
      const int row = rc.x;
      const float* pi = w.PI + (size_t)(row < 0 ? 0 : row) * kPStride;
      named_bar_sync(3, TC_EPI);           // xi staged; sRow / sCol visible
      {
        // vector parts of P_j (vector_down + vector_down_frames contributions of the target node): gather
        float4 pv[(TMT * PF4 + TC_EPI - 1) / TC_EPI];
#pragma unroll
        for (int i = 0; i < (TMT * PF4 + TC_EPI - 1) / TC_EPI; ++i) {
          const int idx = i * TC_EPI + tid;
          pv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (idx < TMT * PF4) {
            const int rr = idx / PF4, cj = T.sCol[rr];
            if (cj >= 0) pv[i] = *reinterpret_cast<const float4*>(w.PJ + (size_t)cj * kPStride + kH + (idx - rr * PF4) * 4);
          }
        }
        // ... and of P_i (rows shared by consecutive edges -> broadcast loads), in flight at the same time
        float vdh[32], vdf0[9];
#pragma unroll
        for (int i = 0; i < 32; ++i) vdh[i] = 0.f;
#pragma unroll
        for (int i = 0; i < 9; ++i) vdf0[i] = 0.f;
        if (row >= 0) {
#pragma unroll
          for (int i = 0; i < H2 * 3; ++i) vdh[i] = pi[kH + half * H2 * 3 + i];
          if (half == 0) {
#pragma unroll
            for (int i = 0; i < 9; ++i) vdf0[i] = pi[kH + HID0 * 3 + i];
          }
        }
        float xi[XD * 3];
#pragma unroll
        for (int c = 0; c < XD * 3; ++c) xi[c] = sF[r * XLD + c];
        named_bar_sync(3, TC_EPI);         // everybody has read xi
#pragma unroll
        for (int i = 0; i < (TMT * PF4 + TC_EPI - 1) / TC_EPI; ++i) {
          const int idx = i * TC_EPI + tid;
          if (idx < TMT * PF4) {
            const int rr = idx / PF4;
            float* dst = sF + rr * PLD + (idx - rr * PF4) * 4;
            dst[0] = pv[i].x; dst[1] = pv[i].y; dst[2] = pv[i].z; dst[3] = pv[i].w;
          }
        }
        named_bar_sync(3, TC_EPI);
        const float* pjv = sF + r * PLD;
        // vector_down rows [half*H2, half*H2 + H2) of GCP 0 (split form: endpoint parts gathered)
#pragma unroll
        for (int i = 0; i < H2 * 3; ++i) vdh[i] += pjv[half * H2 * 3 + i];
        if (half == 0) {
#pragma unroll
          for (int i = 0; i < 9; ++i) vdf0[i] += pjv[HID0 * 3 + i];
        }
#pragma unroll
        for (int c = 0; c < XD; ++c) {
#pragma unroll
          for (int h = 0; h < H2; ++h) {
            const float wd = sw.Wd0x[c * HID0 + half * H2 + h];
            vdh[h * 3 + 0] = fmaf(wd, xi[c * 3 + 0], vdh[h * 3 + 0]);
            vdh[h * 3 + 1] = fmaf(wd, xi[c * 3 + 1], vdh[h * 3 + 1]);
            vdh[h * 3 + 2] = fmaf(wd, xi[c * 3 + 2], vdh[h * 3 + 2]);
          }
        }
#pragma unroll
        for (int h = 0; h < H2; ++h)
          x_store1(X, r, ED + half * H2 + h, safe_norm3(vdh[h * 3], vdh[h * 3 + 1], vdh[h * 3 + 2]));
        tmem_st8xN<4>(tl + TM_VD0 + half * 32, vdh);
        tmem_st_wait();
        if (half == 0) {
          // vector_down_frames of GCP 0 and its scalarisation q0 (9 values), plus the zero padding
#pragma unroll
          for (int c = 0; c < XD; ++c)
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) {
              const float wf = sw.Wf0x[c * 3 + ch];
              vdf0[ch * 3 + 0] = fmaf(wf, xi[c * 3 + 0], vdf0[ch * 3 + 0]);
              vdf0[ch * 3 + 1] = fmaf(wf, xi[c * 3 + 1], vdf0[ch * 3 + 1]);
              vdf0[ch * 3 + 2] = fmaf(wf, xi[c * 3 + 2], vdf0[ch * 3 + 2]);
            }
#pragma unroll
          for (int ch = 0; ch < 3; ++ch)
#pragma unroll
            for (int ax = 0; ax < 3; ++ax)
              x_store1(X, r, ED + HID0 + ch * 3 + ax,
                       f[ax * 3] * vdf0[ch * 3] + f[ax * 3 + 1] * vdf0[ch * 3 + 1] + f[ax * 3 + 2] * vdf0[ch * 3 + 2]);
#pragma unroll
          for (int i = HID0 + 9; i < 32; ++i) x_store1(X, r, ED + i, 0.f);
        }
      }
      publish();

      // ---- E0: m_0 = silu(S0 + P_i[row] + P_j[col]), this half's 128 columns in 4 rounds of 32.  P_i rows are
      //      shared by consecutive edges (broadcast loads); the scalar part of P_j is stored column-major in blocks
      //      of 32 nodes and the edges of a warp have consecutive target nodes, so a column load touches 1-3 lines.  The loads of round
      //      c+1 are in flight while round c is computed.
      {
        float4 pa[8];
        float pj[32];
        const int colj = rc.y;
        const int cjn = colj < 0 ? 0 : colj;
        const float* pjt = w.PJT + ((size_t)(cjn >> 5) * 256 + half * 128) * 32 + (cjn & 31);   // blocked layout
        auto prefetch = [&](int c) {
          const float* pc = pjt + c * 1024;      // 32 columns x 32 nodes per round; immediate offsets below
#pragma unroll
          for (int i = 0; i < 32; ++i) pj[i] = pc[i * 32];
#pragma unroll
          for (int q = 0; q < 8; ++q) pa[q] = *reinterpret_cast<const float4*>(pi + half * 128 + c * 32 + q * 4);
        };
        prefetch(0);
        wait_d();
        for (int c = 0; c < 4; ++c) {
          const int c0 = half * 128 + c * 32;
          float v[32];
          tmem_ld32(tl + TM_S + c0, v);
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            v[q * 4 + 0] += pa[q].x + pj[q * 4 + 0]; v[q * 4 + 1] += pa[q].y + pj[q * 4 + 1];
            v[q * 4 + 2] += pa[q].z + pj[q * 4 + 2]; v[q * 4 + 3] += pa[q].w + pj[q * 4 + 3];
          }
          if (c < 3) prefetch(c + 1);
#pragma unroll
          for (int q = 0; q < 16; ++q) {
            const float2 p0 = silu_fast2(make_float2(v[2 * q], v[2 * q + 1]));
            v[2 * q] = p0.x; v[2 * q + 1] = p0.y;
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) x_store8(X, r, c0 + q * 8, v + q * 8);
        }
      }
      publish();

      float vd[24], vdf[9];
      float2 adot2 = make_float2(0.f, 0.f);
      for (int k = 1; k <= 3; ++k) {
        // ---- E(k)a: gate_{k-1}, m.v update (this half's 16 channels), vector_down of GCP k -> A block 4
        wait_d();
        float part[33];
        if (k == 1) {
          float vd0[HID0 * 3];
          {
            float t0[64];
            float* t1 = t0 + 32;
            tmem_ld64(tl + TM_VD0, t0);
#pragma unroll
            for (int i = 0; i < H2 * 3; ++i) { vd0[i] = t0[i]; vd0[H2 * 3 + i] = t1[i]; }
          }
          gate_update<HID0, true, false>(tl, half, TM_U0, vd0, sw.Wu0, sw.bg[0], sw.Wdk[0], sw.Wfk[0], part);
        } else {
          gate_update<8, false, false>(tl, half, ((k - 1) & 1) ? TM_U1 : TM_U0, vd, sw.Wuk[k - 2], sw.bg[k - 1],
                                       sw.Wdk[k - 1], sw.Wfk[k - 1], part);
        }
#pragma unroll
        for (int i = 0; i < 33; ++i) exch_mine[i] = part[i];
        named_bar_sync(3, TC_EPI);
#pragma unroll
        for (int i = 0; i < 24; ++i) vd[i] = part[i] + exch_other[i];
#pragma unroll
        for (int i = 0; i < 9; ++i) vdf[i] = part[24 + i] + exch_other[24 + i];
        {
          // [vn(8) | q(9) | 0...] -> A K-block 4; half 0 writes columns 0..15, half 1 columns 16..31
          float a[16];
          if (half == 0) {
#pragma unroll
            for (int h = 0; h < 8; ++h) a[h] = safe_norm3(vd[h * 3], vd[h * 3 + 1], vd[h * 3 + 2]);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int ch = i / 3, ax = i - ch * 3;
              a[8 + i] = f[ax * 3] * vdf[ch * 3] + f[ax * 3 + 1] * vdf[ch * 3 + 1] + f[ax * 3 + 2] * vdf[ch * 3 + 2];
            }
          } else {
            a[0] = f[6] * vdf[6] + f[7] * vdf[7] + f[8] * vdf[8];      // q[8]: ch 2, axis 2
#pragma unroll
            for (int i = 1; i < 16; ++i) a[i] = 0.f;
          }
          x_store8(X, r, 256 + half * 16, a);
          x_store8(X, r, 256 + half * 16 + 8, a + 8);
        }
        named_bar_sync(3, TC_EPI);     // exchange buffer may be overwritten by the next phase
        publish();
        // ---- E(k)b: m_k = m_{k-1} + silu(S_k + b_k), this half's 128 columns
        wait_d();
        for (int c0 = half * 128; c0 < half * 128 + 128; c0 += 64) {
          float v[64];
          tmem_ld64(tl + TM_S + c0, v);
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            float m[8];
            x_load8(X, r, c0 + q * 8, m);
            const float4 b0 = *reinterpret_cast<const float4*>(&sw.bk[k - 1][c0 + q * 8]);
            const float4 b1 = *reinterpret_cast<const float4*>(&sw.bk[k - 1][c0 + q * 8 + 4]);
            const float2 bb[4] = {make_float2(b0.x, b0.y), make_float2(b0.z, b0.w), make_float2(b1.x, b1.y),
                                  make_float2(b1.z, b1.w)};
            float2 mm[4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
              mm[i] = __fadd2_rn(make_float2(m[2 * i], m[2 * i + 1]),
                                 silu_fast2(__fadd2_rn(make_float2(v[q * 8 + 2 * i], v[q * 8 + 2 * i + 1]), bb[i])));
            if (k == 3) {
              const float4 w0 = *reinterpret_cast<const float4*>(&sw.wa[c0 + q * 8]);
              const float4 w1 = *reinterpret_cast<const float4*>(&sw.wa[c0 + q * 8 + 4]);
              adot2 = __ffma2_rn(mm[0], make_float2(w0.x, w0.y), adot2);
              adot2 = __ffma2_rn(mm[1], make_float2(w0.z, w0.w), adot2);
              adot2 = __ffma2_rn(mm[2], make_float2(w1.x, w1.y), adot2);
              adot2 = __ffma2_rn(mm[3], make_float2(w1.z, w1.w), adot2);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) { m[2 * i] = mm[i].x; m[2 * i + 1] = mm[i].y; }
            x_store8(X, r, c0 + q * 8, m);
          }
        }
        if (k == 3) T.sAttn[half][r] = adot2.x + adot2.y;
        publish();
      }
      // ---- E4: gate_3 and the last m.v update
      wait_d();
      {
        float dummy[33];
        gate_update<8, false, true>(tl, half, TM_U1, vd, sw.Wuk[2], sw.bg[3], nullptr, nullptr, dummy);
      }
      tc_fence_before();
      named_bar_sync(3, TC_EPI);     // m.v of both halves in TMEM, sAttn/sRow/... visible
      tc_fence_after();
      const float attn = sigmoid_fast(T.sAttn[0][r] + T.sAttn[1][r] + sw.ba[0]);
      // ---- segmented sum over the source node, 64 message columns per round (rounds 0..3 = m.s * attention from the
      // bf16 A tile, 4 = m.v[0:64], 5 = m.v[64:96]).  All 256 threads stage their 32 values of the round as float2
      // pairs into a [128 rows][32 pair slots] buffer (slot index xor-swizzled with the row -> conflict-free 64-bit
      // stores and loads); then warp w scans its 16 rows with lane = column pair: acc = keep*acc + x (keep = 0 at a
      // segment start), and every segment end adds its partial sum to the aggregate row with one 8-byte reduction.
      // (Row segments are cut at the 16-row windows, so all pieces go through RED.ADD; the aggregate rows are zero
      // on entry — the node pass resets them.)
      const int wr0 = warp * 16;
      uint32_t m_start, m_end;
      {
        const int li = lane & 15;
        const int rw = T.sRow[wr0 + li];
        const int rp = li > 0 ? T.sRow[wr0 + li - 1] : -2;
        const int rn = li < 15 ? T.sRow[wr0 + li + 1] : -2;
        m_start = __ballot_sync(0xffffffffu, rw != rp) & 0xffffu;
        m_end = __ballot_sync(0xffffffffu, rw != rn && rw >= 0) & 0xffffu;
      }
      float2* sR2 = reinterpret_cast<float2*>(&T.sT[0][0][0]);
      static_assert(TMT * 32 * 2 <= 2 * TMT * ST_LD, "reduction buffer");
      for (int t = 0; t < 6; ++t) {
        {
          float v[32];
          if (t < 4) {
#pragma unroll
            for (int q = 0; q < 4; ++q) x_load8(X, r, t * 64 + half * 32 + q * 8, v + q * 8);
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] *= attn;
          } else if (t == 4) {
            tmem_ld32(tl + TM_MV + half * 32, v);
          } else {
            tmem_ld8xN<2>(tl + TM_MV + 64 + half * 16, v);
          }
          if (t < 5) {
#pragma unroll
            for (int k = 0; k < 16; ++k) sR2[r * 32 + half * 16 + (k ^ (r & 15))] = make_float2(v[2 * k], v[2 * k + 1]);
          } else {
#pragma unroll
            for (int k = 0; k < 8; ++k) sR2[r * 32 + ((half * 8 + k) ^ (r & 15))] = make_float2(v[2 * k], v[2 * k + 1]);
          }
        }
        named_bar_sync(3, TC_EPI);
        {
          const bool active = t < 5 || lane < 16;
          const int hb = (lane >> 4) * 16, kq = lane & 15;
          float2 acc = make_float2(0.f, 0.f);
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float2 x = sR2[(wr0 + i) * 32 + hb + (kq ^ i)];
            const float keep = ((m_start >> i) & 1u) ? 0.f : 1.f;
            acc = __ffma2_rn(make_float2(keep, keep), acc, x);
            if (((m_end >> i) & 1u) && active)
              atomicAdd(reinterpret_cast<float2*>(w.agg + (size_t)T.sRow[wr0 + i] * kMsg + t * 64 + 2 * lane), acc);
          }
        }
        named_bar_sync(3, TC_EPI);
      }
      named_bar_sync(3, TC_EPI);       // sRow / exchange buffers free for the next tile
      if (tid == 0) BDIFF_STAMP(es++);
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem, 512);
}

// ============================================================================================ launchers
cudaError_t tc_configure() {
  cudaError_t e = cudaFuncSetAttribute(k_edge_message_tc<64, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)TC_SMEM_BYTES);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(k_edge_message_tc<16, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM_BYTES);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(k_umma_selftest, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              2 * X_BLOCK + 2 * RING_STAGE + 1024);
}

bool tc_supported(int Ed, int Xd) { return (Ed == 64 && Xd == 16) || (Ed == 16 && Xd == 8); }

void launch_tc_pack(cudaStream_t st, const LayerW& lw, const Dims& d, unsigned char* blob) {
  const long long rows = (long long)tc_nc0(d.Ed, d.Xd) * 256 + 3 * (4 * 320 + 256) + 4 * 32;
  const long long total = rows * 64;
  k_tc_pack_layer<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(lw, d, blob);
}

void launch_edge_message_tc(cudaStream_t st, const Plan& p, const Dims& d, const LayerW& lw, const unsigned char* blob,
                            const Work& w, int num_sms) {
  const int ntiles = (int)((p.E + TMT - 1) / TMT);
  if (ntiles == 0) return;
  const int grid = ntiles < num_sms ? ntiles : num_sms;
  if (d.Ed == 64) k_edge_message_tc<64, 16><<<grid, TC_THREADS2, TC_SMEM_BYTES, st>>>(p, lw, blob, w, ntiles);
  else k_edge_message_tc<16, 8><<<grid, TC_THREADS2, TC_SMEM_BYTES, st>>>(p, lw, blob, w, ntiles);
}

void launch_umma_selftest(cudaStream_t st, const float* A, const float* W, unsigned char* img_scratch, float* C) {
  k_selftest_pack<<<(320 * 128 + 255) / 256, 256, 0, st>>>(W, img_scratch);
  k_umma_selftest<<<1, TC_THREADS, 2 * X_BLOCK + 2 * RING_STAGE + 1024, st>>>(A, img_scratch, C);
}

}  // namespace bdiff
