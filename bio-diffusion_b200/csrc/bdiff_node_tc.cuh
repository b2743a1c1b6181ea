// bdiff_node_tc.cuh — declarations shared by the tensor-core node pass (bdiff_node_tc.cu) and the layer megakernel
// (bdiff_layers_tc.cu).
#pragma once
#include "bdiff_edge_tc.cuh"

namespace bdiff {

constexpr int NT_EPI = 256;
constexpr int NT_THREADS = NT_EPI + 64;
constexpr int NTM = 128;
constexpr int NRING = 288 * 128;
constexpr int NSTAGES = 2;
constexpr int NM_S = 0, NM_U = 256, NM_CHI = 288, NM_VDF = 384, NM_EX = 432;

struct alignas(16) SmallWN {
  float Wdf[64 * 16], Wff[64 * 3], Wuf[16 * 32], bgf[32];
  float b1[256], b2[256];
  float Wdp[32 * 8], Wfp[32 * 3], Wup[8], bp[256], wgp[256], bgp[4];
  float b0[256];
  float Wd0i[32 * 20], Wd0j[32 * 20], Wf0i[32 * 3], Wf0j[32 * 3];
  float pWd[32 * 32], pWf[32 * 3], pbs[32];
};

constexpr int R4M = 32;
constexpr int NM_S1 = 256;        // second accumulator of the row-replicated kernel (overlaps U, see G4)

// the small weights with the (mutually exclusive) next-layer / projection sets overlaid
struct alignas(16) SmallWR4 {
  float Wdf[64 * 16], Wff[64 * 3], Wuf[16 * 32], bgf[32];
  float b1[256], b2[256];
  float Wdp[32 * 8], Wfp[32 * 3], Wup[8], bp[256], wgp[256], bgp[4];
  union {
    struct { float b0[256], Wd0i[32 * 20], Wd0j[32 * 20], Wf0i[32 * 3], Wf0j[32 * 3]; } nx;
    struct { float pWd[32 * 32], pWf[32 * 3], pbs[32]; } pj;
  } u;
};

struct NodeR4Tail : TcBars {
  SmallWR4 sw;
  float4 sT[8][8 * 8];     // per-warp 8 x 32 fp32 transposition scratch (xor-swizzled 16-byte chunks)
  float sV[R4M][193];      // per node [agg_v (32x3) | chi (32x3)]; chi is replaced by chi_new in E3a
  float sVD[R4M][49];      // vector_down of the feed-forward GCP (16 x 3)
  float sVP[R4M][25];      // vector_down of the position GCP (8 x 3)
  float sDot[8][R4M];
};
constexpr size_t R4_SMEM_BYTES = 5 * (size_t)X_BLOCK + NSTAGES * (size_t)NRING + sizeof(NodeR4Tail) + 1024;

__device__ __forceinline__ void x_store8_rep4(unsigned char* X, int l, int kk, const float* v) {   // kk % 8 == 0
  const uint4 u = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
  unsigned char* q = X + (kk >> 6) * X_BLOCK + sw128_offset(l, kk & 63);
#pragma unroll
  for (int k = 0; k < 4; ++k) *reinterpret_cast<uint4*>(q + k * 4096) = u;
}
__device__ __forceinline__ void x_store1_rep4(unsigned char* X, int l, int kk, float v) {
  const __nv_bfloat16 b = __float2bfloat16_rn(v);
  unsigned char* q = X + (kk >> 6) * X_BLOCK + sw128_offset(l, kk & 63);
#pragma unroll
  for (int k = 0; k < 4; ++k) *reinterpret_cast<__nv_bfloat16*>(q + k * 4096) = b;
}

// Global <-> "lane = row" register tiles through the per-warp scratch, so that every global instruction touches 4
// rows x 128 contiguous bytes instead of 32 rows x 16 bytes (the L1 processes one line tag per cycle).
// v[32] = this lane's row (32 consecutive floats); gbase -> (row 0, first column) of the warp's 32 x 32 block.
__device__ __forceinline__ void warp_store_rows(float4* sc, const float* v, float* gbase, int ld, int lane) {
#pragma unroll
  for (int pass = 0; pass < 4; ++pass) {          // 8 rows per pass (1 KiB of scratch per warp)
    if ((lane >> 3) == pass) {
      const int r = lane & 7;
#pragma unroll
      for (int j = 0; j < 8; ++j) sc[r * 8 + (j ^ r)] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int r = 4 * i + (lane >> 3);
      const float4 t = sc[r * 8 + ((lane & 7) ^ r)];
      *reinterpret_cast<float4*>(gbase + (size_t)(pass * 8 + r) * ld + (lane & 7) * 4) = t;
    }
    __syncwarp();
  }
}
// t[8]: t[i] = float4 #(lane & 7) of row 4i + (lane >> 3) of the block (see the caller's loads)
__device__ __forceinline__ void warp_load_rows(float4* sc, float* v, const float4* t, int lane) {
#pragma unroll
  for (int pass = 0; pass < 4; ++pass) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int r = 4 * i + (lane >> 3);
      sc[r * 8 + ((lane & 7) ^ r)] = t[pass * 2 + i];
    }
    __syncwarp();
    if ((lane >> 3) == pass) {
      const int r = lane & 7;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 x = sc[r * 8 + (j ^ r)];
        v[4 * j] = x.x; v[4 * j + 1] = x.y; v[4 * j + 2] = x.z; v[4 * j + 3] = x.w;
      }
    }
    __syncwarp();
  }
}

}  // namespace bdiff
