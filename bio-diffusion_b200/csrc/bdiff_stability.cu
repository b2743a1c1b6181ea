// bdiff_stability.cu — batched molecular-stability check of sampled molecules.
//
// Replaces the per-molecule Python loop of `check_molecular_stability` (/root/reference/src/datamodules/components/
// edm/__init__.py:91-124: n x n `cdist`, `get_bond_order_batch` :61-88 table lookups with the margins of
// edm/constants.py, row sums, `allowed_bonds` membership) that dominates `sample_and_analyze` once the chain is
// fast.  One CTA per molecule, thread per atom row; integer outputs, bit-exact against the oracle: the distance is
// formed with separately rounded fp32 operations (no FMA contraction) exactly like the elementwise restatement.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/bdiff.h"

namespace bdiff {

__global__ void __launch_bounds__(128) k_stability(const float* __restrict__ x, const int32_t* __restrict__ types,
                                                   const int32_t* __restrict__ mol_off, int A,
                                                   const float* __restrict__ b1, const float* __restrict__ b2,
                                                   const float* __restrict__ b3, float m1, float m2, float m3,
                                                   const uint32_t* __restrict__ allowed, int limit_one,
                                                   int32_t* __restrict__ nr_bonds, int32_t* __restrict__ nr_stable,
                                                   int32_t* __restrict__ mol_stable) {
  const int k = blockIdx.x;
  const int n0 = mol_off[k], n = mol_off[k + 1] - n0;
  __shared__ int cnt;
  if (threadIdx.x == 0) cnt = 0;
  __syncthreads();
  int mine = 0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float xi = x[(size_t)(n0 + i) * 3], yi = x[(size_t)(n0 + i) * 3 + 1], zi = x[(size_t)(n0 + i) * 3 + 2];
    const int ti = types[n0 + i];
    int sum = 0;
    for (int j = 0; j < n; ++j) {
      if (j == i) continue;                                   // np.fill_diagonal(order, 0)
      const float dx = __fsub_rn(xi, x[(size_t)(n0 + j) * 3]), dy = __fsub_rn(yi, x[(size_t)(n0 + j) * 3 + 1]),
                  dz = __fsub_rn(zi, x[(size_t)(n0 + j) * 3 + 2]);
      const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
      const float d = __fmul_rn(100.f, __fsqrt_rn(d2));       // "we change the metric"
      const int idx = ti * A + types[n0 + j];
      int order = 0;
      if (d < __fadd_rn(b1[idx], m1)) order = 1;
      if (d < __fadd_rn(b2[idx], m2)) order = 2;
      if (d < __fadd_rn(b3[idx], m3)) order = 3;
      if (limit_one && order > 1) order = 1;
      sum += order;
    }
    nr_bonds[n0 + i] = sum;
    mine += (sum < 32 && ((allowed[ti] >> sum) & 1u)) ? 1 : 0;
  }
  if (mine) atomicAdd(&cnt, mine);
  __syncthreads();
  if (threadIdx.x == 0) {
    nr_stable[k] = cnt;
    mol_stable[k] = cnt == n ? 1 : 0;
  }
}

// Bond-order matrix of `make_mol_edm` (rdkit_functions.py:276-321): E = tril(get_bond_order_batch(type_i, type_j, |x_i -
// x_j|), -1) per molecule, i.e. the bond type (0 none, 1 single, 2 double, 3 triple) of every pair i > j and 0 elsewhere,
// written as int8 into the molecule's dense [n, n] block at pair_off[k] (row-major, so that nonzero() lists the bonds in
// the order the reference adds them to the RWMol).  One CTA per molecule, threads stride over the n^2 pairs.
__global__ void __launch_bounds__(256) k_bond_orders(const float* __restrict__ x, const int32_t* __restrict__ types,
                                                     const int32_t* __restrict__ mol_off, const int64_t* __restrict__ pair_off,
                                                     int A, const float* __restrict__ b1, const float* __restrict__ b2,
                                                     const float* __restrict__ b3, float m1, float m2, float m3,
                                                     int limit_one, int8_t* __restrict__ E) {
  const int k = blockIdx.x;
  const int n0 = mol_off[k], n = mol_off[k + 1] - n0;
  int8_t* out = E + pair_off[k];
  for (int idx = threadIdx.x; idx < n * n; idx += blockDim.x) {
    const int i = idx / n, j = idx - i * n;
    int order = 0;
    if (j < i) {
      const float dx = __fsub_rn(x[(size_t)(n0 + i) * 3], x[(size_t)(n0 + j) * 3]),
                  dy = __fsub_rn(x[(size_t)(n0 + i) * 3 + 1], x[(size_t)(n0 + j) * 3 + 1]),
                  dz = __fsub_rn(x[(size_t)(n0 + i) * 3 + 2], x[(size_t)(n0 + j) * 3 + 2]);
      const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
      const float d = __fmul_rn(100.f, __fsqrt_rn(d2));
      const int t = types[n0 + i] * A + types[n0 + j];        // cartesian_prod(atom_types, atom_types): (type_i, type_j)
      if (d < __fadd_rn(b1[t], m1)) order = 1;
      if (d < __fadd_rn(b2[t], m2)) order = 2;
      if (d < __fadd_rn(b3[t], m3)) order = 3;
      if (limit_one && order > 1) order = 1;
    }
    out[idx] = (int8_t)order;
  }
}

}  // namespace bdiff

extern "C" int32_t bdiff_bond_orders(void* stream, const float* x, const int32_t* atom_types, const int32_t* mol_off,
                                     const int64_t* pair_off, int32_t num_mols, int32_t num_types, const float* bonds1,
                                     const float* bonds2, const float* bonds3, float margin1, float margin2, float margin3,
                                     int32_t limit_bonds_to_one, int8_t* bond_order) {
  if (!x || !atom_types || !mol_off || !pair_off || num_mols < 1 || num_types < 1 || !bonds1 || !bonds2 || !bonds3 || !bond_order)
    return BDIFF_EINVAL;
  bdiff::k_bond_orders<<<num_mols, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x, atom_types, mol_off, pair_off, num_types, bonds1, bonds2, bonds3, margin1, margin2, margin3, limit_bonds_to_one,
      bond_order);
  return cudaGetLastError() == cudaSuccess ? BDIFF_OK : BDIFF_ECUDA;
}

extern "C" int32_t bdiff_check_stability(void* stream, const float* x, const int32_t* atom_types, const int32_t* mol_off,
                                         int32_t num_mols, int32_t num_types, const float* bonds1, const float* bonds2,
                                         const float* bonds3, float margin1, float margin2, float margin3,
                                         const uint32_t* allowed_mask, int32_t limit_bonds_to_one, int32_t* nr_bonds,
                                         int32_t* nr_stable, int32_t* mol_stable) {
  if (!x || !atom_types || !mol_off || num_mols < 1 || num_types < 1 || !bonds1 || !bonds2 || !bonds3 || !allowed_mask ||
      !nr_bonds || !nr_stable || !mol_stable)
    return BDIFF_EINVAL;
  bdiff::k_stability<<<num_mols, 128, 0, static_cast<cudaStream_t>(stream)>>>(
      x, atom_types, mol_off, num_types, bonds1, bonds2, bonds3, margin1, margin2, margin3, allowed_mask,
      limit_bonds_to_one, nr_bonds, nr_stable, mol_stable);
  return cudaGetLastError() == cudaSuccess ? BDIFF_OK : BDIFF_ECUDA;
}
