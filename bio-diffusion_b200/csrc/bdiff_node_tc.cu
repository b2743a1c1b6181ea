// bdiff_node_tc.cu — tensor-core (tcgen05 / TMEM / TMA-bulk) version of the per-layer node pass.
//
// Same math as k_node_update (gcpnet.py:893-930, :834-857): feed-forward GCP2 on [aggregate | node] with
// residual + mask, position-update GCP2 with x += v, then either the next layer's endpoint projections or the
// final scalar projection.  Tile = 128 nodes = the 128 TMEM lanes, a thread PAIR per node (half 0: accumulator
// columns [0,128) / vector channels [0,16); half 1 the rest).  The A operand (bf16, K-major, 128B swizzle, 5
// K-blocks) is rewritten in place between the chained GEMMs; weights stream from L2 as pre-swizzled bf16
// K-blocks through a 2-stage TMA-bulk ring; the feed-forward vector gate  sigmoid(Wg Z2 + b)  is folded into the
// neighbouring GEMMs via  Wg Z2 = Wg h_new - Wg h_old  (A-negate), like in the edge kernel (a direct N=32 MMA on
// Z2 was tried: same accuracy, one more phase).
// TMEM columns: S 0..255 | U 256..287 | chi (96) 288..383 | VD_ff (48) 384..431 | pair exchange 2x40 432..511.
#include "bdiff_kernels.h"
#include "bdiff_tc.cuh"
#include "bdiff_node_tc.cuh"

namespace bdiff {



size_t tc_node_blob_bytes() { return (size_t)(4 * 256 + 4 * 288 + 256 + 4 * 256 + 4 * 288 + 256 + 8 * 256) * 128; }

// Per-layer bf16 blob in streaming order:
//   G1a 4x[256]: W1[:, 0:256]   | G1b 4x[288]: W1[:, 256:512] + Wg_ff | G1c [256]: W1[:, 512:544]
//   G2  4x[256]: W2             | G3a 4x[288]: Wp[:, 0:256] + Wg_ff   | G3b [256]: Wp[:, 256:288]
//   not last: G4 4x[256]: next.Wsi, G5 4x[256]: next.Wsj            last: Gp 5x[32]: projection scalar_out
__global__ void k_tc_pack_node(LayerW lw, LayerW wn, EmbedW ew, Dims d, int last, unsigned char* __restrict__ blob) {
  const long long total_rows = 4 * 256 + 4 * 288 + 256 + 4 * 256 + 4 * 288 + 256 + (last ? 5 * 32 : 8 * 256);
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total_rows * 64) return;
  long long rowg = idx / 64;
  const int kc = (int)(idx - rowg * 64);
  size_t base = 0;
  float v = 0.f;
  int r = 0;
  auto seg = [&](long long nrows) -> bool {      // is rowg inside the next segment of nrows rows?
    if (rowg < nrows) return true;
    rowg -= nrows;
    base += (size_t)nrows * 128;
    return false;
  };
  if (seg(4 * 256)) {
    const int j = (int)(rowg / 256); r = (int)(rowg % 256); base += (size_t)j * 256 * 128;
    v = lw.W1[(size_t)(j * 64 + kc) * 256 + r];
  } else if (seg(4 * 288)) {
    const int j = (int)(rowg / 288); r = (int)(rowg % 288); base += (size_t)j * 288 * 128;
    const int kk = j * 64 + kc;
    v = r < 256 ? lw.W1[(size_t)(256 + kk) * 256 + r] : lw.Wgf[(size_t)kk * 32 + (r - 256)];
  } else if (seg(256)) {
    r = (int)rowg;
    const int kk = 512 + kc;
    v = kk < kKFF ? lw.W1[(size_t)kk * 256 + r] : 0.f;
  } else if (seg(4 * 256)) {
    const int j = (int)(rowg / 256); r = (int)(rowg % 256); base += (size_t)j * 256 * 128;
    v = lw.W2[(size_t)(j * 64 + kc) * 256 + r];
  } else if (seg(4 * 288)) {
    const int j = (int)(rowg / 288); r = (int)(rowg % 288); base += (size_t)j * 288 * 128;
    const int kk = j * 64 + kc;
    v = r < 256 ? lw.Wp[(size_t)kk * 256 + r] : lw.Wgf[(size_t)kk * 32 + (r - 256)];
  } else if (seg(256)) {
    r = (int)rowg;
    const int kk = 256 + kc;
    v = kk < kKM ? lw.Wp[(size_t)kk * 256 + r] : 0.f;
  } else if (!last) {
    const int which = (int)(rowg / (4 * 256));
    long long rr = rowg - (long long)which * 4 * 256;
    base += (size_t)which * 4 * 256 * 128;
    const int j = (int)(rr / 256); r = (int)(rr % 256); base += (size_t)j * 256 * 128;
    v = (which == 0 ? wn.Wsi : wn.Wsj)[(size_t)(j * 64 + kc) * 256 + r];
  } else {
    const int j = (int)(rowg / 32); r = (int)(rowg % 32); base += (size_t)j * 32 * 128;
    const int kk = j * 64 + kc;
    v = (kk < 300 && r < d.Hin) ? ew.pWs[(size_t)kk * d.Hin + r] : 0.f;
  }
  *reinterpret_cast<__nv_bfloat16*>(blob + base + sw128_offset(r, kc)) = __float2bfloat16_rn(v);
}

struct NodeTcTail : TcBars {
  SmallWN sw;
  float sTw[8][32][33];    // per-warp 32x32 transposition scratch (coalesced global stores of accumulator tiles)
  float sMask[NTM];
  float sDot[2][NTM];
};

constexpr size_t NT_SMEM_BYTES = 5 * (size_t)X_BLOCK + NSTAGES * (size_t)NRING + sizeof(NodeTcTail) + 1024;

// Coalesced staging of a [128 x 128] fp32 block (row stride ld floats) into A columns [kk0, kk0+128) as bf16:
// each warp-iteration reads one contiguous 512-byte half-row and scatters 8-byte packs into the swizzled tile.
__device__ __forceinline__ void stage_rows(unsigned char* X, const float* __restrict__ base, int ld, int kk0, int wih,
                                           int lane) {
  const int kk = kk0 + lane * 4;
#pragma unroll
  for (int b = 0; b < 2; ++b) {           // 16 independent 512-byte row loads in flight per warp
    float4 v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = *reinterpret_cast<const float4*>(base + (size_t)(wih + 4 * (b * 16 + i)) * ld + lane * 4);
#pragma unroll
    for (int i = 0; i < 16; ++i)
      *reinterpret_cast<uint2*>(X + (kk >> 6) * X_BLOCK + sw128_offset(wih + 4 * (b * 16 + i), kk & 63)) =
          make_uint2(pack_bf16x2(v[i].x, v[i].y), pack_bf16x2(v[i].z, v[i].w));
  }
}

__global__ void __launch_bounds__(NT_THREADS, 1)
    k_node_update_tc(Plan p, Dims d, LayerW lw, LayerW wn, EmbedW ew, const unsigned char* __restrict__ blob, Work w,
                     int last, int ntiles) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char* X = smem;
  unsigned char* ring = smem + 5 * X_BLOCK;
  NodeTcTail& T = *reinterpret_cast<NodeTcTail*>(ring + NSTAGES * NRING);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int hid0 = d.hid0;

  if (tid == 0) {
    for (int i = 0; i < NSTAGES; ++i) { mbar_init(&T.full[i], 1); mbar_init(&T.empty[i], 1); }
    mbar_init(&T.a_ready, NT_EPI);
    mbar_init(&T.d_full, 1);
    mbar_fence_init();
  }
  if (warp == 8) tmem_alloc(&T.tmem_ptr, 512);
  // small weights -> shared memory by TMA bulk copies (all in flight at once; sizes rounded up to 16 bytes, every
  // packed array is 256-byte aligned and padded)
  if (tid == 0) {
    mbar_init(&T.wbar, 1);
    mbar_fence_init();
    SmallWN& s = T.sw;
    auto sz = [](int n) { return (uint32_t)((n * 4 + 15) & ~15); };
    uint32_t total = sz(1024) + sz(192) + sz(512) + sz(32) + 2 * sz(256) + sz(256) + sz(96) + sz(8) + 2 * sz(256) + sz(1);
    total += last ? sz(1024) + sz(96) + sz(d.Hin) : sz(256) + 2 * sz(32 * hid0) + 2 * sz(96);
    mbar_expect_tx(&T.wbar, total);
    auto cp = [&](float* dst, const float* src, int n) { bulk_g2s(dst, src, sz(n), &T.wbar); };
    cp(s.Wdf, lw.Wdf, 64 * 16); cp(s.Wff, lw.Wff, 64 * 3); cp(s.Wuf, lw.Wuf, 16 * 32); cp(s.bgf, lw.bgf, 32);
    cp(s.b1, lw.b1, 256); cp(s.b2, lw.b2, 256);
    cp(s.Wdp, lw.Wdp, 32 * 8); cp(s.Wfp, lw.Wfp, 32 * 3); cp(s.Wup, lw.Wup, 8); cp(s.bp, lw.bp, 256);
    cp(s.wgp, lw.Wgp, 256); cp(s.bgp, lw.bgp, 1);
    if (!last) {
      cp(s.b0, wn.b0, 256);
      cp(s.Wd0i, wn.Wd0i, 32 * hid0); cp(s.Wd0j, wn.Wd0j, 32 * hid0); cp(s.Wf0i, wn.Wf0i, 96); cp(s.Wf0j, wn.Wf0j, 96);
    } else {
      cp(s.pWd, ew.pWd, 32 * 32); cp(s.pWf, ew.pWf, 96); cp(s.pbs, ew.pbs, d.Hin);
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = T.tmem_ptr;
  if (warp < 8) mbar_wait(&T.wbar, 0);

  if (warp == 8) {
    // ===================================================================== TMA producer (one lane)
    if (lane == 0) {
      uint32_t ci = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        size_t off = 0;
        auto push = [&](uint32_t bytes) {
          const uint32_t s = ci % NSTAGES;
          mbar_wait_backoff(&T.empty[s], ((ci / NSTAGES) & 1) ^ 1);
          mbar_expect_tx(&T.full[s], bytes);
          bulk_g2s(ring + s * NRING, blob + off, bytes, &T.full[s]);
          off += bytes;
          ++ci;
        };
        for (int j = 0; j < 4; ++j) push(256 * 128);      // G1a
        for (int j = 0; j < 4; ++j) push(288 * 128);      // G1b (+ gate rows)
        push(256 * 128);                                  // G1c
        for (int j = 0; j < 4; ++j) push(256 * 128);      // G2
        for (int j = 0; j < 4; ++j) push(288 * 128);      // G3a (+ gate rows)
        push(256 * 128);                                  // G3b
        if (!last) { for (int j = 0; j < 8; ++j) push(256 * 128); }
        else { for (int j = 0; j < 5; ++j) push(32 * 128); }
      }
    }
  } else if (warp == 9) {
    // ======================================================================= MMA issuer (one lane)
    if (lane == 0) {
      const uint32_t i256 = umma_idesc_bf16(256, false), i32 = umma_idesc_bf16(32, false),
                     i32n = umma_idesc_bf16(32, true);
      const uint32_t xaddr = smem_u32(X), raddr = smem_u32(ring);
      uint32_t ci = 0, pa = 0;
      int ms = 28;
      auto wait_a = [&]() { mbar_wait_backoff(&T.a_ready, pa); pa ^= 1; tc_fence_after(); BDIFF_STAMP(ms++); };
      auto wait_w = [&]() -> uint32_t {
        const uint32_t s = ci % NSTAGES;
        mbar_wait_backoff(&T.full[s], (ci / NSTAGES) & 1);
        tc_fence_after();
        return raddr + s * NRING;
      };
      auto done_w = [&]() { umma_commit(&T.empty[ci % NSTAGES]); ++ci; };
      auto stamp_commit = [&]() { BDIFF_STAMP(ms++); };
      auto gemm256 = [&](bool fresh) {     // 4 K-blocks of X against 4 chunks of 256 rows -> S
        for (int j = 0; j < 4; ++j) {
          const uint32_t wb = wait_w();
          for (int s = 0; s < 4; ++s)
            umma_bf16(tmem + NM_S, umma_desc_sw128(xaddr + j * X_BLOCK + s * 32), umma_desc_sw128(wb + s * 32), i256,
                      fresh ? (j | s) > 0 : true);
          done_w();
        }
      };
      auto gemm288 = [&](bool fresh_s, bool negate_u) {   // ... plus the 32 gate columns -> U (= +/- Wg . A)
        for (int j = 0; j < 4; ++j) {
          const uint32_t wb = wait_w();
          for (int s = 0; s < 4; ++s) {
            const uint64_t ad = umma_desc_sw128(xaddr + j * X_BLOCK + s * 32);
            umma_bf16(tmem + NM_S, ad, umma_desc_sw128(wb + s * 32), i256, fresh_s ? (j | s) > 0 : true);
            umma_bf16(tmem + NM_U, ad, umma_desc_sw128(wb + 256 * 128 + s * 32), negate_u ? i32n : i32,
                      negate_u ? (j | s) > 0 : true);
          }
          done_w();
        }
      };
      auto gemm_extra = [&]() {            // K-block 4 (32 columns) against one chunk of 256 rows, accumulate
        const uint32_t wb = wait_w();
        for (int s = 0; s < 2; ++s)
          umma_bf16(tmem + NM_S, umma_desc_sw128(xaddr + 4 * X_BLOCK + s * 32), umma_desc_sw128(wb + s * 32), i256, true);
        done_w();
      };
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        wait_a(); gemm256(true); umma_commit(&T.d_full); stamp_commit();                            // G1a: agg_s . W1a
        wait_a(); gemm288(false, true); gemm_extra(); umma_commit(&T.d_full); stamp_commit();       // G1b/c: + h . W1b, U = -Wg h, + [vn|q] . W1c
        wait_a(); gemm256(true); umma_commit(&T.d_full); stamp_commit();                            // G2: Y . W2
        wait_a(); gemm288(true, false); umma_commit(&T.d_full); stamp_commit();                     // G3a: h_new . Wp, U += Wg h_new
        wait_a(); gemm_extra(); umma_commit(&T.d_full); stamp_commit();                             // G3b
        if (!last) {
          wait_a(); gemm256(true); umma_commit(&T.d_full); stamp_commit();                          // G4: h_new . Wsi(next)
          wait_a(); gemm256(true); umma_commit(&T.d_full); stamp_commit();                          // G5: h_new . Wsj(next)
        } else {
          wait_a();                                                                 // Gp: [h_new | vn | q] . Wproj -> U
          for (int j = 0; j < 5; ++j) {
            const uint32_t wb = wait_w();
            const int ns = j < 4 ? 4 : 3;
            for (int s = 0; s < ns; ++s)
              umma_bf16(tmem + NM_U, umma_desc_sw128(xaddr + j * X_BLOCK + s * 32), umma_desc_sw128(wb + s * 32), i32,
                        (j | s) > 0);
            done_w();
          }
          umma_commit(&T.d_full); stamp_commit();
        }
      }
    }
  } else {
    // ====================================================== epilogue / compute warps (thread pair <-> node r)
    const int half = tid >> 7, r = tid & 127;
    const uint32_t tl = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    const SmallWN& sw = T.sw;
    uint32_t pd = 0;
    int es = 0, fs = 52;
    auto wait_d = [&]() { if (tid == 0) BDIFF_STAMP(es++); mbar_wait(&T.d_full, pd); pd ^= 1; tc_fence_after(); if (tid == 0) BDIFF_STAMP(es++); };
    auto publish = [&]() { fence_proxy_async(); tc_fence_before(); mbar_arrive(&T.a_ready); if (tid == 0) BDIFF_STAMP(es++); };
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      if (tid == 0) BDIFF_STAMP(es++);
      const int node = tile * NTM + r;
      const bool valid = node < p.N;
      const float m = (valid && p.mask[node]) ? 1.f : 0.f;
      const int wih = warp & 3;                      // warp index inside the half == TMEM lane quarter
      float (*tw)[33] = T.sTw[warp];
      const int row0 = wih * 32;                     // first tile row of this warp
      if (half == 0) T.sMask[r] = m;
      float* ag = w.agg + (size_t)node * kMsg;
      float* crow = w.chi + (size_t)node * 96;
      float f[9];
      {
        const float4 f0 = *reinterpret_cast<const float4*>(w.fbar + (size_t)node * 12);
        const float4 f1 = *reinterpret_cast<const float4*>(w.fbar + (size_t)node * 12 + 4);
        f[0] = f0.x; f[1] = f0.y; f[2] = f0.z; f[3] = f0.w; f[4] = f1.x; f[5] = f1.y; f[6] = f1.z; f[7] = f1.w;
        f[8] = w.fbar[(size_t)node * 12 + 8];
      }
      // ---- T0: agg_s -> A blocks 0..3 (coalesced); vector_down (this half's 8 rows) / vector_down_frames of the FF GCP
      if (tid == 0) BDIFF_STAMP(fs++);
      stage_rows(X, w.agg + (size_t)tile * NTM * kMsg + half * 128, kMsg, half * 128, wih, lane);
      if (tid == 0) BDIFF_STAMP(fs++);
      {
        float vdh[24], vdf[9];
#pragma unroll
        for (int i = 0; i < 24; ++i) vdh[i] = 0.f;
#pragma unroll
        for (int i = 0; i < 9; ++i) vdf[i] = 0.f;
#pragma unroll 4
        for (int cc = 0; cc < 16; ++cc) {       // 4 channels (12 floats) at a time: [agg_v (32 ch) | chi (32 ch)]
          const float* src = cc < 8 ? ag + kH + cc * 12 : crow + (cc - 8) * 12;
          const float4 a = *reinterpret_cast<const float4*>(src), b = *reinterpret_cast<const float4*>(src + 4),
                       c = *reinterpret_cast<const float4*>(src + 8);
          const float vin[12] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int ch = cc * 4 + j;
            const float4 w0 = *reinterpret_cast<const float4*>(&sw.Wdf[ch * 16 + half * 8]);
            const float4 w1 = *reinterpret_cast<const float4*>(&sw.Wdf[ch * 16 + half * 8 + 4]);
            const float wd[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
            for (int h = 0; h < 8; ++h) {
              vdh[h * 3 + 0] = fmaf(wd[h], vin[j * 3 + 0], vdh[h * 3 + 0]);
              vdh[h * 3 + 1] = fmaf(wd[h], vin[j * 3 + 1], vdh[h * 3 + 1]);
              vdh[h * 3 + 2] = fmaf(wd[h], vin[j * 3 + 2], vdh[h * 3 + 2]);
            }
            if (half == 0) {
#pragma unroll
              for (int q = 0; q < 3; ++q) {
                const float wf = sw.Wff[ch * 3 + q];
                vdf[q * 3 + 0] = fmaf(wf, vin[j * 3 + 0], vdf[q * 3 + 0]);
                vdf[q * 3 + 1] = fmaf(wf, vin[j * 3 + 1], vdf[q * 3 + 1]);
                vdf[q * 3 + 2] = fmaf(wf, vin[j * 3 + 2], vdf[q * 3 + 2]);
              }
            }
          }
        }
        if (tid == 0) BDIFF_STAMP(fs++);
#pragma unroll
        for (int h = 0; h < 8; ++h)
          x_store1(X, r, 256 + half * 8 + h, safe_norm3(vdh[h * 3], vdh[h * 3 + 1], vdh[h * 3 + 2]));
        tmem_st8xN<3>(tl + NM_VDF + half * 24, vdh);
        if (half == 0) {
#pragma unroll
          for (int ch = 0; ch < 3; ++ch)
#pragma unroll
            for (int ax = 0; ax < 3; ++ax)
              x_store1(X, r, 256 + 16 + ch * 3 + ax,
                       f[ax * 3] * vdf[ch * 3] + f[ax * 3 + 1] * vdf[ch * 3 + 1] + f[ax * 3 + 2] * vdf[ch * 3 + 2]);
#pragma unroll
          for (int i = 25; i < 32; ++i) x_store1(X, r, 256 + i, 0.f);
        }
        // own 16 chi channels -> TMEM scratch (needed for the residual in E3a)
        {
          float cown[48];
#pragma unroll
          for (int q = 0; q < 12; ++q) {
            const float4 a = *reinterpret_cast<const float4*>(crow + half * 48 + q * 4);
            cown[q * 4] = a.x; cown[q * 4 + 1] = a.y; cown[q * 4 + 2] = a.z; cown[q * 4 + 3] = a.w;
          }
          tmem_st8xN<6>(tl + NM_CHI + half * 48, cown);
        }
        if (tid == 0) BDIFF_STAMP(fs++);
        tmem_st_wait();
        if (tid == 0) BDIFF_STAMP(fs++);
      }
      publish();
      // ---- T0b: h -> A blocks 0..3 (after G1a has consumed agg_s)
      wait_d();
      stage_rows(X, w.h + (size_t)tile * NTM * kH + half * 128, kH, half * 128, wih, lane);
      publish();
      // ---- E1: Y = silu(S + b1)
      wait_d();
      for (int c0 = half * 128; c0 < half * 128 + 128; c0 += 32) {
        float v[32];
        tmem_ld32(tl + NM_S + c0, v);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 bb = *reinterpret_cast<const float4*>(&sw.b1[c0 + q * 4]);
          v[q * 4 + 0] = silu_fast(v[q * 4 + 0] + bb.x);
          v[q * 4 + 1] = silu_fast(v[q * 4 + 1] + bb.y);
          v[q * 4 + 2] = silu_fast(v[q * 4 + 2] + bb.z);
          v[q * 4 + 3] = silu_fast(v[q * 4 + 3] + bb.w);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) x_store8(X, r, c0 + q * 8, v + q * 8);
      }
      publish();
      // ---- E2: h_new = (h + S + b2) * mask -> global h (fp32) through a per-warp 32x32 transpose (full 128-byte
      //      lines) and, from the same loop, bf16 into A blocks 0..3
      wait_d();
      for (int c0 = half * 128; c0 < half * 128 + 128; c0 += 32) {
        float v[32];
        tmem_ld32(tl + NM_S + c0, v);
#pragma unroll
        for (int i = 0; i < 32; ++i) tw[lane][i] = v[i];
        __syncwarp();
        float* hb = w.h + ((size_t)tile * NTM + row0) * kH + c0 + lane;
        const float b2v = sw.b2[c0 + lane];
        float ho[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) ho[i] = hb[(size_t)i * kH];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float hn = (ho[i] + tw[i][lane] + b2v) * T.sMask[row0 + i];
          hb[(size_t)i * kH] = hn;
          x_store1(X, row0 + i, c0 + lane, hn);
        }
        __syncwarp();
      }
      publish();
      // ---- E3a: FF vector gate, chi_new for this half's 16 channels, vector_down of the position GCP
      wait_d();
      float vdp[24], vdfp[9];
      {
        float vdff[48], u[16], co[48], part[40];
        {
          uint32_t r0[48], r1[16], r2[48];
#pragma unroll
          for (int q = 0; q < 6; ++q) tmem_ld8_nw(tl + NM_VDF + q * 8, r0 + q * 8);
          tmem_ld8_nw(tl + NM_U + half * 16, r1);
          tmem_ld8_nw(tl + NM_U + half * 16 + 8, r1 + 8);
#pragma unroll
          for (int q = 0; q < 6; ++q) tmem_ld8_nw(tl + NM_CHI + half * 48 + q * 8, r2 + q * 8);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 48; ++i) { vdff[i] = __uint_as_float(r0[i]); co[i] = __uint_as_float(r2[i]); }
#pragma unroll
          for (int i = 0; i < 16; ++i) u[i] = __uint_as_float(r1[i]);
        }
#pragma unroll
        for (int i = 0; i < 40; ++i) part[i] = 0.f;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int o = half * 16 + j;
          const float g = sigmoid_fast(u[j] + sw.bgf[o]);
          float s0 = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll
          for (int h = 0; h < 16; ++h) {
            const float wu = sw.Wuf[h * 32 + o];
            s0 = fmaf(wu, vdff[h * 3 + 0], s0);
            s1 = fmaf(wu, vdff[h * 3 + 1], s1);
            s2 = fmaf(wu, vdff[h * 3 + 2], s2);
          }
          const float c0v = (co[j * 3 + 0] + s0 * g) * m, c1v = (co[j * 3 + 1] + s1 * g) * m,
                      c2v = (co[j * 3 + 2] + s2 * g) * m;
          co[j * 3 + 0] = c0v; co[j * 3 + 1] = c1v; co[j * 3 + 2] = c2v;
          const float4 w0 = *reinterpret_cast<const float4*>(&sw.Wdp[o * 8]);
          const float4 w1 = *reinterpret_cast<const float4*>(&sw.Wdp[o * 8 + 4]);
          const float wd[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
          for (int h = 0; h < 8; ++h) {
            part[h * 3 + 0] = fmaf(wd[h], c0v, part[h * 3 + 0]);
            part[h * 3 + 1] = fmaf(wd[h], c1v, part[h * 3 + 1]);
            part[h * 3 + 2] = fmaf(wd[h], c2v, part[h * 3 + 2]);
          }
#pragma unroll
          for (int q = 0; q < 3; ++q) {
            const float wf = sw.Wfp[o * 3 + q];
            part[24 + q * 3 + 0] = fmaf(wf, c0v, part[24 + q * 3 + 0]);
            part[24 + q * 3 + 1] = fmaf(wf, c1v, part[24 + q * 3 + 1]);
            part[24 + q * 3 + 2] = fmaf(wf, c2v, part[24 + q * 3 + 2]);
          }
        }
        tmem_st8xN<6>(tl + NM_CHI + half * 48, co);
        tmem_st8xN<5>(tl + NM_EX + half * 40, part);
#pragma unroll
        for (int q = 0; q < 12; ++q)
          *reinterpret_cast<float4*>(crow + half * 48 + q * 4) = make_float4(co[q * 4], co[q * 4 + 1], co[q * 4 + 2], co[q * 4 + 3]);
        tmem_st_wait();
        tc_fence_before();
        named_bar_sync(3, NT_EPI);
        tc_fence_after();
        float other[40];
        tmem_ld8xN<5>(tl + NM_EX + (half ^ 1) * 40, other);
#pragma unroll
        for (int i = 0; i < 24; ++i) vdp[i] = part[i] + other[i];
#pragma unroll
        for (int i = 0; i < 9; ++i) vdfp[i] = part[24 + i] + other[24 + i];
        float a[16];
        if (half == 0) {
#pragma unroll
          for (int h = 0; h < 8; ++h) a[h] = safe_norm3(vdp[h * 3], vdp[h * 3 + 1], vdp[h * 3 + 2]);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int ch = i / 3, ax = i - ch * 3;
            a[8 + i] = f[ax * 3] * vdfp[ch * 3] + f[ax * 3 + 1] * vdfp[ch * 3 + 1] + f[ax * 3 + 2] * vdfp[ch * 3 + 2];
          }
        } else {
          a[0] = f[6] * vdfp[6] + f[7] * vdfp[7] + f[8] * vdfp[8];
#pragma unroll
          for (int i = 1; i < 16; ++i) a[i] = 0.f;
        }
        x_store8(X, r, 256 + half * 16, a);
        x_store8(X, r, 256 + half * 16 + 8, a + 8);
      }
      publish();
      // ---- E3b: position GCP: gate = sigmoid(wg . silu(S + bp) + bg), x += (Wu . VD) * gate
      wait_d();
      {
        float pdot = 0.f;
        for (int c0 = half * 128; c0 < half * 128 + 128; c0 += 32) {
          float v[32];
          tmem_ld32(tl + NM_S + c0, v);
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 bb = *reinterpret_cast<const float4*>(&sw.bp[c0 + q * 4]);
            const float4 wg = *reinterpret_cast<const float4*>(&sw.wgp[c0 + q * 4]);
            pdot = fmaf(silu_fast(v[q * 4 + 0] + bb.x), wg.x, pdot);
            pdot = fmaf(silu_fast(v[q * 4 + 1] + bb.y), wg.y, pdot);
            pdot = fmaf(silu_fast(v[q * 4 + 2] + bb.z), wg.z, pdot);
            pdot = fmaf(silu_fast(v[q * 4 + 3] + bb.w), wg.w, pdot);
          }
        }
        T.sDot[half][r] = pdot;
        named_bar_sync(3, NT_EPI);
        if (half == 0) {
          const float gp = sigmoid_fast(T.sDot[0][r] + T.sDot[1][r] + sw.bgp[0]);
#pragma unroll
          for (int x = 0; x < 3; ++x) {
            float s = 0.f;
#pragma unroll
            for (int h = 0; h < 8; ++h) s = fmaf(sw.Wup[h], vdp[h * 3 + x], s);
            const float xn = (w.x[(size_t)node * 3 + x] + s * gp) * m;
            w.x[(size_t)node * 3 + x] = xn;
            if (xn != xn) atomicExch(w.nan_flag, 1);
          }
        }
        named_bar_sync(3, NT_EPI);     // sDot free again
      }
      if (last) {
        // projection GCP2 (256,32)->(Hin,0): [vn(32) | q(9) | 0] -> A block 4 columns 0..47
        float chi[96];
        tmem_ld8xN<12>(tl + NM_CHI, chi);
        for (int h = half * 16; h < half * 16 + 16; ++h) {
          float s0 = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll
          for (int c = 0; c < 32; ++c) {
            const float wd = sw.pWd[c * 32 + h];
            s0 = fmaf(wd, chi[c * 3 + 0], s0);
            s1 = fmaf(wd, chi[c * 3 + 1], s1);
            s2 = fmaf(wd, chi[c * 3 + 2], s2);
          }
          x_store1(X, r, 256 + h, safe_norm3(s0, s1, s2));
        }
        if (half == 0) {
          for (int ch = 0; ch < 3; ++ch) {
            float s0 = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll
            for (int c = 0; c < 32; ++c) {
              const float wf = sw.pWf[c * 3 + ch];
              s0 = fmaf(wf, chi[c * 3 + 0], s0);
              s1 = fmaf(wf, chi[c * 3 + 1], s1);
              s2 = fmaf(wf, chi[c * 3 + 2], s2);
            }
#pragma unroll
            for (int ax = 0; ax < 3; ++ax)
              x_store1(X, r, 256 + 32 + ch * 3 + ax, f[ax * 3] * s0 + f[ax * 3 + 1] * s1 + f[ax * 3 + 2] * s2);
          }
#pragma unroll
          for (int i = 41; i < 48; ++i) x_store1(X, r, 256 + i, 0.f);
        }
      }
      publish();
      if (!last) {
        // ---- E4: PI scalar part = S + b0 (this half's columns); vector parts of PI (half 0) / PJ (half 1)
        wait_d();
        for (int c0 = half * 128; c0 < half * 128 + 128; c0 += 32) {
          float v[32];
          tmem_ld32(tl + NM_S + c0, v);
#pragma unroll
          for (int i = 0; i < 32; ++i) tw[lane][i] = v[i];
          __syncwarp();
          float* pb = w.PI + ((size_t)tile * NTM + row0) * kPStride + c0 + lane;
          const float b0v = sw.b0[c0 + lane];
#pragma unroll 8
          for (int i = 0; i < 32; ++i) pb[(size_t)i * kPStride] = tw[i][lane] + b0v;
          __syncwarp();
        }
        publish();
        {
          float chi[96];
          tmem_ld8xN<12>(tl + NM_CHI, chi);
          float* vrow = (half == 0 ? w.PI : w.PJ) + (size_t)node * kPStride + kH;
          const float* Wd = half == 0 ? sw.Wd0i : sw.Wd0j;
          const float* Wf = half == 0 ? sw.Wf0i : sw.Wf0j;
          for (int h = 0; h < hid0; ++h) {
            float s0 = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll
            for (int c = 0; c < 32; ++c) {
              const float wd = Wd[c * hid0 + h];
              s0 = fmaf(wd, chi[c * 3 + 0], s0);
              s1 = fmaf(wd, chi[c * 3 + 1], s1);
              s2 = fmaf(wd, chi[c * 3 + 2], s2);
            }
            vrow[h * 3 + 0] = s0; vrow[h * 3 + 1] = s1; vrow[h * 3 + 2] = s2;
          }
          for (int ch = 0; ch < 3; ++ch) {
            float s0 = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll
            for (int c = 0; c < 32; ++c) {
              const float wf = Wf[c * 3 + ch];
              s0 = fmaf(wf, chi[c * 3 + 0], s0);
              s1 = fmaf(wf, chi[c * 3 + 1], s1);
              s2 = fmaf(wf, chi[c * 3 + 2], s2);
            }
            vrow[hid0 * 3 + ch * 3 + 0] = s0; vrow[hid0 * 3 + ch * 3 + 1] = s1; vrow[hid0 * 3 + ch * 3 + 2] = s2;
          }
        }
        // ---- E5: PJ scalar part = S, stored column-major in blocks of 32 nodes (lane = node -> 128 contiguous bytes per column)
        wait_d();
        for (int c0 = half * 128; c0 < half * 128 + 128; c0 += 32) {
          float v[32];
          tmem_ld32(tl + NM_S + c0, v);
          float* pb = w.PJT + ((size_t)(tile * 4 + (warp & 3)) * 256 + c0) * 32 + lane;   // block of 32 nodes = this warp's rows
#pragma unroll
          for (int i = 0; i < 32; ++i) pb[i * 32] = v[i];
        }
      } else {
        // ---- Ep: projected scalars = U + bias
        wait_d();
        if (half == 0) {
          float v[32];
          tmem_ld32(tl + NM_U, v);
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (i < d.Hin) w.hproj[(size_t)node * 32 + i] = v[i] + sw.pbs[i];
        }
      }
      // reset the tile's aggregate rows for the next layer's edge pass (its atomics need zeros); the vector part was
      // read by both halves in T0, so order the reset after everybody's reads
      named_bar_sync(3, NT_EPI);
      for (int rr = warp; rr < NTM; rr += 8) {
        float4* z = reinterpret_cast<float4*>(w.agg + ((size_t)tile * NTM + rr) * kMsg);
        for (int c4 = lane; c4 < kMsg / 4; c4 += 32) z[c4] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      tc_fence_before();
      named_bar_sync(3, NT_EPI);     // both halves are done with S / U / scratch before the next tile's MMAs
      tc_fence_after();
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem, 512);
}

__global__ void k_node_update_r4(Plan p, Dims d, LayerW lw, LayerW wn, EmbedW ew, const unsigned char* __restrict__ blob,
                                 Work w, int last, int ntiles);

cudaError_t tc_node_configure() {
  cudaError_t e = cudaFuncSetAttribute(k_node_update_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)NT_SMEM_BYTES);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(k_node_update_r4, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)R4_SMEM_BYTES);
}

void launch_tc_pack_node(cudaStream_t st, const LayerW& lw, const LayerW& wn, const EmbedW& ew, const Dims& d, int last,
                         unsigned char* blob) {
  const long long rows = 4 * 256 + 4 * 288 + 256 + 4 * 256 + 4 * 288 + 256 + (last ? 5 * 32 : 8 * 256);
  const long long total = rows * 64;
  k_tc_pack_node<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(lw, wn, ew, d, last, blob);
}

void launch_node_update_tc(cudaStream_t st, const Plan& p, const Dims& d, const LayerW& lw, const LayerW& wn,
                           const EmbedW& ew, const unsigned char* blob, const Work& w, int last, int num_sms) {
  // small batches (one wave of 32-node tiles): the row-replicated kernel; BDIFF_NODE_R4=0/1 overrides
  const char* env = getenv("BDIFF_NODE_R4");      // read per launch so that tests can exercise both kernels
  const int force = env ? atoi(env) : -1;
  const int nt32 = (p.N + 31) / 32;
  if (force == 1 || (force < 0 && nt32 <= num_sms)) {
    const int grid = nt32 < num_sms ? nt32 : num_sms;
    k_node_update_r4<<<grid, NT_THREADS, R4_SMEM_BYTES, st>>>(p, d, lw, wn, ew, blob, w, last, nt32);
    return;
  }
  const int ntiles = (p.N + NTM - 1) / NTM;
  const int grid = ntiles < num_sms ? ntiles : num_sms;
  k_node_update_tc<<<grid, NT_THREADS, NT_SMEM_BYTES, st>>>(p, d, lw, wn, ew, blob, w, last, ntiles);
}

// ================================================================================================================
// Row-replicated variant for small batches (one wave of 32-node tiles, i.e. ceil(N/32) <= #SMs).
// The chained GEMMs of this pass have M = #nodes, which at QM9/GEOM sampling batch sizes is ~2.4k rows: 19 tiles of
// 128 leave 129 SMs idle and the pass is bound by the per-thread epilogue latency of those 19 CTAs.  Here a tile is
// 32 nodes and every node is written FOUR times into the 128-row A operand (rows l, l+32, l+64, l+96; +4096 bytes
// each in the swizzled K-block), so the M=128 MMA leaves node l's accumulator row in all four TMEM lane quarters.
// The 8 epilogue warps then share the same 32 nodes: warp s owns accumulator columns [32s, 32s+32) (it can reach
// node l in its own lane quarter) and 1/8 of the vector-channel work; 4x more CTAs, 1/4 of the per-thread work, no
// transposes (each thread reads/writes a contiguous 128-byte slice of its node's row).  Tensor work is replicated,
// which is free: the tensor pipe is <10% busy in this pass.  Same MMA/TMA warps, blob and barriers as above.
__global__ void __launch_bounds__(NT_THREADS, 1)
    k_node_update_r4(Plan p, Dims d, LayerW lw, LayerW wn, EmbedW ew, const unsigned char* __restrict__ blob, Work w,
                     int last, int ntiles) {
  constexpr int NSTRIDE = NRING;      // ring stage stride (the layer megakernel uses the edge pass's larger stage)
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char* X = smem;
  unsigned char* ring = smem + 5 * X_BLOCK;
  NodeR4Tail& T = *reinterpret_cast<NodeR4Tail*>(ring + NSTAGES * NRING);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int hid0 = d.hid0;

  if (tid == 0) {
    for (int i = 0; i < NSTAGES; ++i) { mbar_init(&T.full[i], 1); mbar_init(&T.empty[i], 1); }
    mbar_init(&T.a_ready, NT_EPI);
    mbar_init(&T.d_full, 1);
    mbar_init(&T.wbar, 1);
    mbar_init(&T.u_free, NT_EPI);
    mbar_fence_init();
    SmallWR4& s = T.sw;
    auto sz = [](int n) { return (uint32_t)((n * 4 + 15) & ~15); };
    uint32_t total = sz(1024) + sz(192) + sz(512) + sz(32) + 2 * sz(256) + sz(256) + sz(96) + sz(8) + 2 * sz(256) + sz(1);
    total += last ? sz(1024) + sz(96) + sz(d.Hin) : sz(256) + 2 * sz(32 * hid0) + 2 * sz(96);
    mbar_expect_tx(&T.wbar, total);
    auto cp = [&](float* dst, const float* src, int n) { bulk_g2s(dst, src, sz(n), &T.wbar); };
    cp(s.Wdf, lw.Wdf, 64 * 16); cp(s.Wff, lw.Wff, 64 * 3); cp(s.Wuf, lw.Wuf, 16 * 32); cp(s.bgf, lw.bgf, 32);
    cp(s.b1, lw.b1, 256); cp(s.b2, lw.b2, 256);
    cp(s.Wdp, lw.Wdp, 32 * 8); cp(s.Wfp, lw.Wfp, 32 * 3); cp(s.Wup, lw.Wup, 8); cp(s.bp, lw.bp, 256);
    cp(s.wgp, lw.Wgp, 256); cp(s.bgp, lw.bgp, 1);
    if (!last) {
      cp(s.u.nx.b0, wn.b0, 256);
      cp(s.u.nx.Wd0i, wn.Wd0i, 32 * hid0); cp(s.u.nx.Wd0j, wn.Wd0j, 32 * hid0); cp(s.u.nx.Wf0i, wn.Wf0i, 96); cp(s.u.nx.Wf0j, wn.Wf0j, 96);
    } else {
      cp(s.u.pj.pWd, ew.pWd, 32 * 32); cp(s.u.pj.pWf, ew.pWf, 96); cp(s.u.pj.pbs, ew.pbs, d.Hin);
    }
  }
  if (warp == 8) tmem_alloc(&T.tmem_ptr, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = T.tmem_ptr;

  if (warp == 8) {
    // ===================================================================== TMA producer (one lane)
    if (lane == 0) {
      uint32_t ci = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
#include "node_r4_tile_producer.inc"
      }
    }
  } else if (warp == 9) {
    // ======================================================================= MMA issuer (one lane)
    if (lane == 0) {
      const uint32_t i256 = umma_idesc_bf16(256, false), i32 = umma_idesc_bf16(32, false),
                     i32n = umma_idesc_bf16(32, true);
      const uint32_t xaddr = smem_u32(X), raddr = smem_u32(ring);
      uint32_t ci = 0, pa = 0;
      int ms = 28;
      auto wait_a = [&]() { mbar_wait_backoff(&T.a_ready, pa); pa ^= 1; tc_fence_after(); BDIFF_STAMP(ms++); };
      auto wait_w = [&]() -> uint32_t {
        const uint32_t s = ci % NSTAGES;
        mbar_wait_backoff(&T.full[s], (ci / NSTAGES) & 1);
        tc_fence_after();
        return raddr + s * NRING;
      };
      auto done_w = [&]() { umma_commit(&T.empty[ci % NSTAGES]); ++ci; };
      auto commit_d = [&]() { umma_commit(&T.d_full); BDIFF_STAMP(ms++); };
      auto gemm256 = [&](bool fresh, uint32_t dcol = NM_S) {
        for (int j = 0; j < 4; ++j) {
          const uint32_t wb = wait_w();
          for (int s = 0; s < 4; ++s)
            umma_bf16(tmem + dcol, umma_desc_sw128(xaddr + j * X_BLOCK + s * 32), umma_desc_sw128(wb + s * 32), i256,
                      fresh ? (j | s) > 0 : true);
          done_w();
        }
      };
      uint32_t pu = 0;
      auto gemm288 = [&](bool fresh_s, bool negate_u) {
        for (int j = 0; j < 4; ++j) {
          const uint32_t wb = wait_w();
          for (int s = 0; s < 4; ++s) {
            const uint64_t ad = umma_desc_sw128(xaddr + j * X_BLOCK + s * 32);
            umma_bf16(tmem + NM_S, ad, umma_desc_sw128(wb + s * 32), i256, fresh_s ? (j | s) > 0 : true);
            umma_bf16(tmem + NM_U, ad, umma_desc_sw128(wb + 256 * 128 + s * 32), negate_u ? i32n : i32,
                      negate_u ? (j | s) > 0 : true);
          }
          done_w();
        }
      };
      auto gemm_extra = [&]() {
        const uint32_t wb = wait_w();
        for (int s = 0; s < 2; ++s)
          umma_bf16(tmem + NM_S, umma_desc_sw128(xaddr + 4 * X_BLOCK + s * 32), umma_desc_sw128(wb + s * 32), i256, true);
        done_w();
      };
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
#include "node_r4_tile_mma.inc"
      }
    }
  } else {
    // ============================================== epilogue / compute warps: lane l = node, warp s = column slice
    mbar_wait(&T.wbar, 0);
    const int l = lane, s = warp, c0 = warp * 32;
    const uint32_t tl = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    const SmallWR4& sw = T.sw;
    uint32_t pd = 0;
    int es = 0;
    auto wait_d = [&]() { if (tid == 0) BDIFF_STAMP(es++); mbar_wait(&T.d_full, pd); pd ^= 1; tc_fence_after(); if (tid == 0) BDIFF_STAMP(es++); };
    auto publish = [&]() { fence_proxy_async(); tc_fence_before(); mbar_arrive(&T.a_ready); if (tid == 0) BDIFF_STAMP(es++); };
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
#include "node_r4_tile_epilogue.inc"
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem, 512);
}

}  // namespace bdiff
