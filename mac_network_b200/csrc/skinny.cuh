// Skinny fp32 GEMM for the latency-critical M <= 64 projections of the cell (projY, newMemory, gate, ctrlProj,
// qInput*: ops.linear at mac_cell.py:442-448, 322, 352, 363; ops.py:689):
//     Y[M<=64, N] = epilogue( Aview[M, K] @ W[K, N] )
// 33-100 MFLOP against 1-3 MB of weights: the problem is latency, not throughput.  Each CTA takes a 64 x 32 output
// tile and ONE K-slice (K / 8), loads its operands in a single round trip, multiplies from shared memory, and the 8
// CTAs of a thread-block cluster that share an output tile reduce their partial tiles over distributed shared
// memory (DSMEM) in a fixed order -- deterministic, one launch, no global scratch, no atomics.
#pragma once
#include <cooperative_groups.h>
#include "common.cuh"
#include "sgemm.cuh"

namespace mac {
namespace cg = cooperative_groups;

constexpr int SK_CLUSTER = 8;
constexpr int SK_BM = 64, SK_BN = 32, SK_THREADS = 128;

static __global__ void __launch_bounds__(SK_THREADS) skinny_gemm_kernel(const SgemmParams p, int ks, int dbg) {
  extern __shared__ __align__(16) float sk_smem[];
  constexpr int XLD = SK_BM + 4;                // 16-byte aligned rows of the transposed A slice
  float* ws = sk_smem;                          // [ks][SK_BN]       W slice
  float* part = ws + (size_t)ks * SK_BN;        // [SK_BM][SK_BN]    this CTA's partial tile
  float* xs = part + SK_BM * SK_BN;             // [ks][XLD]         A slice, transposed
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();   // K-slice index (cluster spans gridDim.x)
  const int tid = threadIdx.x;
  const int n0 = blockIdx.y * SK_BN;
  const int k0 = rank * ks;

  // ---- operand load in batches of 8 float4 per thread: all requests of a batch are in flight before the first
  //      shared-memory store of the batch (ks = 64 -> one round trip for A and one for W)
  {
    constexpr int BATCH = 8;
    const int nx = SK_BM * (ks / 4), nw = ks * (SK_BN / 4);
    for (int base = 0; base < nx && !(dbg & 1); base += BATCH * SK_THREADS) {
      float4 rx[BATCH];
#pragma unroll
      for (int i = 0; i < BATCH; ++i) {
        const int f = base + tid + i * SK_THREADS;
        if (f < nx) rx[i] = sg_load_a(p, f % SK_BM, k0 + (f / SK_BM) * 4);   // lane <-> row: conflict-free transposed stores
      }
#pragma unroll
      for (int i = 0; i < BATCH; ++i) {
        const int f = base + tid + i * SK_THREADS;
        if (f < nx) {
          const int row = f % SK_BM, kq = f / SK_BM;
          xs[(kq * 4 + 0) * XLD + row] = rx[i].x;
          xs[(kq * 4 + 1) * XLD + row] = rx[i].y;
          xs[(kq * 4 + 2) * XLD + row] = rx[i].z;
          xs[(kq * 4 + 3) * XLD + row] = rx[i].w;
        }
      }
    }
    for (int base = 0; base < nw && !(dbg & 1); base += BATCH * SK_THREADS) {
      float4 rw[BATCH];
#pragma unroll
      for (int i = 0; i < BATCH; ++i) {
        const int f = base + tid + i * SK_THREADS;
        if (f < nw) {
          const int kr = f / (SK_BN / 4), n = n0 + (f % (SK_BN / 4)) * 4;
          rw[i] = (n < p.N) ? __ldg(reinterpret_cast<const float4*>(p.W + (size_t)(k0 + kr) * p.ldw + n))
                            : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
#pragma unroll
      for (int i = 0; i < BATCH; ++i) {
        const int f = base + tid + i * SK_THREADS;
        if (f < nw) *reinterpret_cast<float4*>(ws + (size_t)(f / (SK_BN / 4)) * SK_BN + (f % (SK_BN / 4)) * 4) = rw[i];
      }
    }
  }
  __syncthreads();

  // ---- 64 x 32 x ks product, 4 x 4 outputs per thread (2 LDS.128 per 16 FMAs: shared-memory bandwidth is the limit)
  const int rg = tid >> 3, cgp = tid & 7;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  if (!(dbg & 2)) {
#pragma unroll 4
    for (int k = 0; k < ks; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(xs + k * XLD + rg * 4);
      const float4 w = *reinterpret_cast<const float4*>(ws + k * SK_BN + cgp * 4);
      acc[0][0] = fmaf(a.x, w.x, acc[0][0]); acc[0][1] = fmaf(a.x, w.y, acc[0][1]);
      acc[0][2] = fmaf(a.x, w.z, acc[0][2]); acc[0][3] = fmaf(a.x, w.w, acc[0][3]);
      acc[1][0] = fmaf(a.y, w.x, acc[1][0]); acc[1][1] = fmaf(a.y, w.y, acc[1][1]);
      acc[1][2] = fmaf(a.y, w.z, acc[1][2]); acc[1][3] = fmaf(a.y, w.w, acc[1][3]);
      acc[2][0] = fmaf(a.z, w.x, acc[2][0]); acc[2][1] = fmaf(a.z, w.y, acc[2][1]);
      acc[2][2] = fmaf(a.z, w.z, acc[2][2]); acc[2][3] = fmaf(a.z, w.w, acc[2][3]);
      acc[3][0] = fmaf(a.w, w.x, acc[3][0]); acc[3][1] = fmaf(a.w, w.y, acc[3][1]);
      acc[3][2] = fmaf(a.w, w.z, acc[3][2]); acc[3][3] = fmaf(a.w, w.w, acc[3][3]);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
    *reinterpret_cast<float4*>(part + (rg * 4 + i) * SK_BN + cgp * 4) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
  cluster.sync();

  // ---- DSMEM reduction: CTA `rank` owns rows [8*rank, 8*rank+8) of the tile; fixed summation order
#pragma unroll
  for (int h = 0; h < (SK_BM / SK_CLUSTER) * SK_BN / SK_THREADS; ++h) {
    const int e = h * SK_THREADS + tid;
    const int rr = rank * (SK_BM / SK_CLUSTER) + e / SK_BN, c = e % SK_BN;
    float s = 0.f;
#pragma unroll
    for (int z = 0; z < SK_CLUSTER; ++z) {
      const float* remote = (dbg & 4) ? part : cluster.map_shared_rank(part, z);
      s += remote[rr * SK_BN + c];
    }
    const int m = rr, n = n0 + c;
    if (m < p.M && n < p.N) {
      float t = s + p.bias_const + (p.bias ? __ldg(p.bias + n) : 0.f);
      const size_t o = (size_t)m * p.ldy + n;
      if (p.epi == EPI_GATE) {
        const float z = sigmoid_f(t);
        if (p.gate_z) p.gate_z[o] = z;
        p.Y[o] = p.gnew[o] * z + p.gold[o] * (1.f - z);
      } else {
        t = apply_act(p.act, t);
        p.Y[o] = p.accumulate ? p.Y[o] + t : t;
      }
    }
  }
  cluster.sync();      // keep every CTA's shared memory alive until all remote reads are done
}

// usable when M <= 64, K splits evenly into 8 slices of whole float4s, and the slice fits in shared memory
inline bool skinny_ok(const SgemmParams& p) {
  if (p.M > SK_BM || (p.epi != EPI_BIAS_ACT && p.epi != EPI_GATE) || p.a_mode != A_SEGS) return false;
  if (p.K % (SK_CLUSTER * 4)) return false;
  const int ks = p.K / SK_CLUSTER;
  if (ks > 256) return false;
  return true;
}

inline int skinny_launch(const SgemmParams& p, cudaStream_t stream) {
  const int ks = p.K / SK_CLUSTER;
  const size_t smem = ((size_t)ks * (SK_BM + 4) + (size_t)ks * SK_BN + SK_BM * SK_BN) * sizeof(float);
  MAC_CUDA_TRY(cudaFuncSetAttribute(skinny_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(SK_CLUSTER, (p.N + SK_BN - 1) / SK_BN, 1);
  cfg.blockDim = dim3(SK_THREADS, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = SK_CLUSTER;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int dbg = 0;
  if (const char* e = getenv("MAC_SK_DEBUG")) dbg = atoi(e);
  MAC_CUDA_TRY(cudaLaunchKernelEx(&cfg, skinny_gemm_kernel, p, ks, dbg));
  MAC_LAUNCH_CHECK();
  return MAC_OK;
}

}  // namespace mac
