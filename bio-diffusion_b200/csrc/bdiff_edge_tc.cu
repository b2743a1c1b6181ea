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
#include "bdiff_edge_tc.cuh"

namespace bdiff {


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
#include "edge_tile_producer.inc"
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
#include "edge_tile_mma.inc"
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
#include "edge_tile_epilogue.inc"
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
