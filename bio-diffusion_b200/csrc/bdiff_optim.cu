// bdiff_optim.cu — the optimiser side of a GCDM training step as three multi-tensor kernels:
//   adaptive gradient clipping by the recent gradient-norm history   (qm9_mol_gen_ddpm.py:1267-1304, Queue
//                                                                      src/models/__init__.py:442-466, get_grad_norm :90-113)
//   AdamW with amsgrad, lr 1e-4, weight decay 1e-12                   (configs/model/*_mol_gen_ddpm.yaml:3-8; torch 1.12
//                                                                      torch/optim/adamw.py _single_tensor_adamw)
//   EMA of the weights, decay 0.9999, every step                      (configs/callbacks/ema.yaml:5-11, src/utils/__init__.py:125-142)
// The reference does this with one small PyTorch kernel per tensor per operation (432 tensors x ~12 ops) plus two
// host syncs (float(grad_norm)); here the norm history lives on the device, nothing synchronises, and every parameter
// element is read and written exactly once per step: HBM-bound, 6 fp32 arrays in / 5 out per element.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/bdiff.h"

namespace bdiff {

constexpr int OPT_CHUNK = 16384;      // elements per CTA
constexpr int OPT_THREADS = 256;
// state words (int32 / float views of the same buffer): see include/bdiff.h (bdiff_optimizer_step)
constexpr int S_STEP = 0, S_QLEN = 1, S_QPOS = 2, S_NORM = 3, S_MAXNORM = 4, S_COEF = 5, S_CLIPPED = 6, S_QUEUE = 8;

__global__ void __launch_bounds__(OPT_THREADS) k_opt_sumsq(const bdiff_opt_tensor* __restrict__ tensors,
                                                           const int32_t* __restrict__ chunk_tensor,
                                                           const int64_t* __restrict__ chunk_start,
                                                           double* __restrict__ partial) {
  const bdiff_opt_tensor t = tensors[chunk_tensor[blockIdx.x]];
  const int64_t s = chunk_start[blockIdx.x];
  const int64_t n = min((int64_t)OPT_CHUNK, t.numel - s);
  const float* g = t.grad + s;
  float acc = 0.f;
  const int64_t n4 = (((uintptr_t)g & 15) == 0) ? n / 4 : 0;
  for (int64_t i = threadIdx.x; i < n4; i += OPT_THREADS) {
    const float4 q = reinterpret_cast<const float4*>(g)[i];
    acc = fmaf(q.x, q.x, acc); acc = fmaf(q.y, q.y, acc); acc = fmaf(q.z, q.z, acc); acc = fmaf(q.w, q.w, acc);
  }
  for (int64_t i = n4 * 4 + threadIdx.x; i < n; i += OPT_THREADS) acc = fmaf(g[i], g[i], acc);
  __shared__ double red[OPT_THREADS];
  red[threadIdx.x] = (double)acc;
  __syncthreads();
  for (int o = OPT_THREADS / 2; o > 0; o >>= 1) {          // fixed tree: deterministic
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = red[0];
}

// One CTA: total norm (fixed order), history statistics, clip coefficient, history update, step count.
__global__ void __launch_bounds__(OPT_THREADS) k_opt_control(const double* __restrict__ partial, int num_chunks,
                                                             int32_t* __restrict__ state, bdiff_opt_hyper hp) {
  __shared__ double red[OPT_THREADS];
  double acc = 0.0;
  for (int i = threadIdx.x; i < num_chunks; i += OPT_THREADS) acc += partial[i];
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int o = OPT_THREADS / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    float* fs = reinterpret_cast<float*>(state);
    const float norm = (float)sqrt(red[0]);
    float coef = 1.f, maxn = 0.f;
    int clipped = 0;
    if (hp.clip) {
      const int qlen = state[S_QLEN];
      double m = 0.0, v = 0.0;
      for (int i = 0; i < qlen; ++i) m += (double)fs[S_QUEUE + i];
      m /= (double)(qlen > 0 ? qlen : 1);
      for (int i = 0; i < qlen; ++i) { const double dlt = (double)fs[S_QUEUE + i] - m; v += dlt * dlt; }
      v /= (double)(qlen > 0 ? qlen : 1);                       // np.std: population standard deviation
      maxn = (float)(1.5 * m + 2.0 * sqrt(v));
      coef = fminf(maxn / (norm + 1e-6f), 1.f);                 // torch.nn.utils.clip_grad_norm_
      clipped = norm > maxn;
      const float push = clipped ? maxn : norm;
      const int cap = hp.queue_len < 1 ? 1 : (hp.queue_len > BDIFF_OPT_MAX_QUEUE ? BDIFF_OPT_MAX_QUEUE : hp.queue_len);
      int pos = state[S_QPOS];
      fs[S_QUEUE + pos] = push;                                  // ring buffer: statistics do not depend on the order
      state[S_QPOS] = (pos + 1) % cap;
      if (qlen < cap) state[S_QLEN] = qlen + 1;
    }
    fs[S_NORM] = norm; fs[S_MAXNORM] = maxn; fs[S_COEF] = coef;
    state[S_CLIPPED] = clipped;
    state[S_STEP] += 1;
  }
}

__global__ void __launch_bounds__(OPT_THREADS) k_opt_apply(const bdiff_opt_tensor* __restrict__ tensors,
                                                           const int32_t* __restrict__ chunk_tensor,
                                                           const int64_t* __restrict__ chunk_start,
                                                           const int32_t* __restrict__ state, bdiff_opt_hyper hp) {
  const bdiff_opt_tensor t = tensors[chunk_tensor[blockIdx.x]];
  const int64_t s = chunk_start[blockIdx.x];
  const int64_t n = min((int64_t)OPT_CHUNK, t.numel - s);
  const float coef = reinterpret_cast<const float*>(state)[S_COEF];
  const int step = state[S_STEP];                                // already incremented by k_opt_control
  const float bc1 = 1.f - powf(hp.beta1, (float)step);
  const float bc2s = sqrtf(1.f - powf(hp.beta2, (float)step));
  const float step_size = hp.lr / bc1;
  const float decay = 1.f - hp.lr * hp.weight_decay;
  float* p = t.param + s;
  const float* g = t.grad + s;
  float* m = t.exp_avg + s;
  float* v = t.exp_avg_sq + s;
  float* vm = (hp.amsgrad && t.max_exp_avg_sq) ? t.max_exp_avg_sq + s : nullptr;
  float* e = t.ema ? t.ema + s : nullptr;
  auto upd = [&](float& pi, float gi, float& mi, float& vi, float* vmi, float* ei) {
    gi *= coef;
    pi *= decay;
    mi = mi * hp.beta1 + (1.f - hp.beta1) * gi;
    vi = vi * hp.beta2 + (1.f - hp.beta2) * gi * gi;
    float vv = vi;
    if (vmi) { vv = fmaxf(*vmi, vi); *vmi = vv; }
    const float denom = sqrtf(vv) / bc2s + hp.eps;
    pi -= step_size * (mi / denom);
    if (ei) *ei = *ei - (*ei - pi) * (1.f - hp.ema_decay);
  };
  // 16-byte accesses when every array of this chunk is 16-byte aligned (torch allocations are; s is a multiple of 4)
  const bool vec = (((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v | (uintptr_t)vm | (uintptr_t)e) & 15) == 0;
  const int64_t n4 = vec ? n / 4 : 0;
  for (int64_t i = threadIdx.x; i < n4; i += OPT_THREADS) {
    float4 P = reinterpret_cast<float4*>(p)[i], M = reinterpret_cast<float4*>(m)[i], V = reinterpret_cast<float4*>(v)[i];
    const float4 G = reinterpret_cast<const float4*>(g)[i];
    float4 VM = vm ? reinterpret_cast<float4*>(vm)[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    float4 E = e ? reinterpret_cast<float4*>(e)[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    upd(P.x, G.x, M.x, V.x, vm ? &VM.x : nullptr, e ? &E.x : nullptr);
    upd(P.y, G.y, M.y, V.y, vm ? &VM.y : nullptr, e ? &E.y : nullptr);
    upd(P.z, G.z, M.z, V.z, vm ? &VM.z : nullptr, e ? &E.z : nullptr);
    upd(P.w, G.w, M.w, V.w, vm ? &VM.w : nullptr, e ? &E.w : nullptr);
    reinterpret_cast<float4*>(p)[i] = P; reinterpret_cast<float4*>(m)[i] = M; reinterpret_cast<float4*>(v)[i] = V;
    if (vm) reinterpret_cast<float4*>(vm)[i] = VM;
    if (e) reinterpret_cast<float4*>(e)[i] = E;
  }
  for (int64_t i = n4 * 4 + threadIdx.x; i < n; i += OPT_THREADS) {
    float pi = p[i], mi = m[i], vi = v[i];
    float vmi = vm ? vm[i] : 0.f, ei = e ? e[i] : 0.f;
    upd(pi, g[i], mi, vi, vm ? &vmi : nullptr, e ? &ei : nullptr);
    p[i] = pi; m[i] = mi; v[i] = vi;
    if (vm) vm[i] = vmi;
    if (e) e[i] = ei;
  }
}

}  // namespace bdiff

extern "C" {

int32_t bdiff_optimizer_chunk(void) { return bdiff::OPT_CHUNK; }

int32_t bdiff_optimizer_step(void* stream, const bdiff_opt_tensor* tensors_dev, const int32_t* chunk_tensor_dev,
                             const int64_t* chunk_start_dev, int32_t num_chunks, double* partial_dev,
                             int32_t* state_dev, const bdiff_opt_hyper* hyper) {
  if (!tensors_dev || !chunk_tensor_dev || !chunk_start_dev || num_chunks < 1 || !partial_dev || !state_dev || !hyper)
    return BDIFF_EINVAL;
  if (!(hyper->lr >= 0.f) || !(hyper->beta1 >= 0.f && hyper->beta1 < 1.f) || !(hyper->beta2 >= 0.f && hyper->beta2 < 1.f) ||
      !(hyper->eps >= 0.f) || !(hyper->ema_decay >= 0.f && hyper->ema_decay <= 1.f))
    return BDIFF_EINVAL;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  bdiff::k_opt_sumsq<<<num_chunks, bdiff::OPT_THREADS, 0, st>>>(tensors_dev, chunk_tensor_dev, chunk_start_dev, partial_dev);
  bdiff::k_opt_control<<<1, bdiff::OPT_THREADS, 0, st>>>(partial_dev, num_chunks, state_dev, *hyper);
  bdiff::k_opt_apply<<<num_chunks, bdiff::OPT_THREADS, 0, st>>>(tensors_dev, chunk_tensor_dev, chunk_start_dev, state_dev, *hyper);
  return cudaGetLastError() == cudaSuccess ? BDIFF_OK : BDIFF_ECUDA;
}

}  // extern "C"
