// The cell's M <= 128 projections on tensor cores (ops.linear at mac_cell.py:442-448 qInput / qInput{i}, 322 ctrlProj,
// 352 newMemory, 363 gate; ops.py:689 projY):
//
//     Y[M, N] = epilogue( concat_k(x_0 .. x_{nseg-1})[M, K] @ W[K, N] )        M <= 128 (the batch), fp32 in, fp32 out
//
// These are latency problems (33-100 MFLOP against 0.5-2 MB of weights), and they sit on the recurrent state path, so
// they must not lose the state's precision to bf16: the activations are split on the fly into bf16 hi + lo parts and the
// weights are pre-split the same way (mac_pack_weight_bf16_split), and three tcgen05 products accumulate into one fp32
// accumulator in tensor memory,
//     D = A_hi B_hi + A_lo B_hi + A_hi B_lo        (the dropped A_lo B_lo term is ~2^-18 relative),
// which keeps the result at fp32-class accuracy (~1e-5) while the MACs run on the tensor pipe (UMMA 128 x BN x 16).
// With wt_lo == NULL it is a plain single-pass bf16 product.
//
// One CTA per BN output columns (BN = 32 / 64 -> 16..96 CTAs pull the weights from L2 in parallel).  Warp roles:
//   warp 0      TMA producer: weight k-blocks [BN x 64] (hi and lo) through a 4-stage ring
//   warp 1      TMEM allocator + MMA issuer
//   warps 2..9  workers: load the fp32 activation k-block [M x 64] with 16-byte loads -- the loads of k-block i+2 are in
//               flight while k-block i is split into bf16 hi / lo and written in the 128-byte-swizzled K-major layout
//               tcgen05 reads (4-stage ring) -- then warps 2..5 run the epilogue (bias, activation / write gate, column
//               split of the folded write unit), thread == output row.
// (First version: 4 worker warps, loads issued per k-block: 18.7 us at K = 512 -- one L2 round trip per k-block on the
// critical path; the fp32 cluster kernel took 9.5 us.)
#pragma once
#include "tc_gemm.cuh"

namespace mac {

constexpr int ST_A_STAGES = 4;
constexpr int ST_B_STAGES = 6;
constexpr int ST_WORKERS = 256;
constexpr int ST_THREADS = 64 + ST_WORKERS;
constexpr int ST_F4 = 128 * 16 / ST_WORKERS;  // float4 groups per worker thread and k-block at M = 128
constexpr int ST_A_TILE = 128 * 128;          // [128 rows x 64 bf16]

struct SkinnyTcParams {
  const float* a[4];
  int ak[4];
  int lda[4];
  int nseg;
  int M, N, K;
  int split;               // 1: hi/lo three-pass product
  const float* bias;       // [N] or NULL
  float bias_const;
  int act;
  float* Y;
  int ldy;
  float* Y2;               // columns >= n_split go to Y2[m, n - n_split] (same ldy) when Y2 != NULL
  int n_split;
  const float* gnew;       // write gate (mac_cell.py:358-367) when != NULL: z = sigmoid(t); Y = gnew*z + gold*(1-z)
  const float* gold;
  float* gate_z;
};

template <int BN>
struct StCfg {
  static constexpr int B_TILE = BN * 128;                                 // [BN rows x 64 bf16]
  static constexpr int A_BYTES = ST_A_STAGES * 2 * ST_A_TILE;             // hi + lo per stage
  static constexpr int B_BYTES = ST_B_STAGES * 2 * B_TILE;
  static constexpr int SMEM_BYTES = A_BYTES + B_BYTES + 1024 /*align*/ + 256 /*barriers*/;
};

__device__ __forceinline__ uint32_t bf16_bits_rn(float x) {
  return (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(x));
}

template <int BN>
__global__ void __launch_bounds__(ST_THREADS, 1)
skinny_tc_kernel(const __grid_constant__ CUtensorMap map_hi, const __grid_constant__ CUtensorMap map_lo, const SkinnyTcParams p) {
  using C = StCfg<BN>;
  extern __shared__ unsigned char smem_dyn[];
  const uint32_t base_u32 = smem_u32(smem_dyn);
  const uint32_t pad = (1024u - (base_u32 & 1023u)) & 1023u;
  unsigned char* a_tiles = smem_dyn + pad;                       // [stage][hi | lo]
  unsigned char* b_tiles = a_tiles + C::A_BYTES;                 // [stage][hi | lo]
  uint64_t* bars = reinterpret_cast<uint64_t*>(b_tiles + C::B_BYTES);
  uint64_t* a_ready = bars;                        // [4] workers -> MMA (8 warp arrivals)
  uint64_t* a_empty = bars + 4;                    // [4] MMA -> workers
  uint64_t* b_full = bars + 8;                     // [6] TMA -> MMA
  uint64_t* b_empty = bars + 14;                   // [6] MMA -> TMA
  uint64_t* done = bars + 20;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 21);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * BN;
  const int kblocks = p.K / TC_BK;
  const int passes_bytes = (p.split ? 2 : 1) * C::B_TILE;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_hi);
    if (p.split) tma_prefetch_desc(&map_lo);
#pragma unroll
    for (int i = 0; i < ST_A_STAGES; ++i) {
      mbar_init(&a_ready[i], ST_WORKERS / 32);
      mbar_init(&a_empty[i], 1);
    }
#pragma unroll
    for (int i = 0; i < ST_B_STAGES; ++i) {
      mbar_init(&b_full[i], 1);
      mbar_init(&b_empty[i], 1);
    }
    mbar_init(done, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr, BN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    if (elect_one()) {
      for (int kb = 0; kb < kblocks; ++kb) {
        const int s = kb % ST_B_STAGES, n = kb / ST_B_STAGES;
        mbar_wait(&b_empty[s], (n & 1) ^ 1);
        mbar_expect_tx(&b_full[s], passes_bytes);
        unsigned char* dst = b_tiles + s * 2 * C::B_TILE;
        tma_load_2d(dst, &map_hi, kb * TC_BK, n0, &b_full[s]);
        if (p.split) tma_load_2d(dst + C::B_TILE, &map_lo, kb * TC_BK, n0, &b_full[s]);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    constexpr uint32_t idesc = make_idesc_bf16(128, BN);
    for (int kb = 0; kb < kblocks; ++kb) {
      const int sa = kb % ST_A_STAGES, na = kb / ST_A_STAGES;
      const int sb = kb % ST_B_STAGES, nb = kb / ST_B_STAGES;
      mbar_wait(&a_ready[sa], na & 1);
      mbar_wait(&b_full[sb], nb & 1);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t sa_u = smem_u32(a_tiles + sa * 2 * ST_A_TILE), sb_u = smem_u32(b_tiles + sb * 2 * C::B_TILE);
        const uint64_t a_hi = make_sw128_kmajor_desc(sa_u), a_lo = make_sw128_kmajor_desc(sa_u + ST_A_TILE);
        const uint64_t b_hi = make_sw128_kmajor_desc(sb_u), b_lo = make_sw128_kmajor_desc(sb_u + C::B_TILE);
#pragma unroll
        for (int k = 0; k < TC_BK / 16; ++k) umma_bf16(tmem_base, a_hi + 2 * k, b_hi + 2 * k, idesc, (kb | k) ? 1u : 0u);
        if (p.split) {
#pragma unroll
          for (int k = 0; k < TC_BK / 16; ++k) umma_bf16(tmem_base, a_lo + 2 * k, b_hi + 2 * k, idesc, 1u);
#pragma unroll
          for (int k = 0; k < TC_BK / 16; ++k) umma_bf16(tmem_base, a_hi + 2 * k, b_lo + 2 * k, idesc, 1u);
        }
        umma_commit(&a_empty[sa]);
        umma_commit(&b_empty[sb]);
        if (kb == kblocks - 1) umma_commit(done);
      }
      __syncwarp();
    }
  } else {
    // ===================================================== workers: activation split, then the epilogue
    const int wt = threadIdx.x - 64;                 // 0..255
    const int nf4 = p.M * 16;                        // float4 groups of one [M x 64] k-block
    // k-block -> (segment pointer at that k, leading dimension)
    auto kb_src = [&](int kb, int& ld) -> const float* {
      int k0 = kb * TC_BK, sg = 0;
      while (sg + 1 < p.nseg && k0 >= p.ak[sg]) { k0 -= p.ak[sg]; ++sg; }
      ld = p.lda[sg];
      return p.a[sg] + k0;
    };
    auto load_kb = [&](int kb, float4 (&v)[ST_F4]) {
      int ld;
      const float* src = kb_src(kb, ld);
#pragma unroll
      for (int i = 0; i < ST_F4; ++i) {
        const int e = wt + ST_WORKERS * i;
        if (e < nf4) {
          const uint4 u = ldg_nc_v4(src + (size_t)(e >> 4) * ld + (e & 15) * 4);
          v[i] = make_float4(__uint_as_float(u.x), __uint_as_float(u.y), __uint_as_float(u.z), __uint_as_float(u.w));
        }
      }
    };
    auto store_kb = [&](int kb, const float4 (&v)[ST_F4]) {
      const int sa = kb % ST_A_STAGES, na = kb / ST_A_STAGES;
      mbar_wait(&a_empty[sa], (na & 1) ^ 1);
      unsigned char* t_hi = a_tiles + sa * 2 * ST_A_TILE;
      unsigned char* t_lo = t_hi + ST_A_TILE;
#pragma unroll
      for (int i = 0; i < ST_F4; ++i) {
        const int e = wt + ST_WORKERS * i;
        if (e < nf4) {
          const int row = e >> 4, f4 = e & 15;
          const uint32_t off = row * 128 + (((f4 >> 1) ^ (row & 7)) << 4) + (f4 & 1) * 8;
          const uint32_t h0 = bf16_bits_rn(v[i].x), h1 = bf16_bits_rn(v[i].y), h2 = bf16_bits_rn(v[i].z), h3 = bf16_bits_rn(v[i].w);
          *reinterpret_cast<uint2*>(t_hi + off) = make_uint2(h0 | (h1 << 16), h2 | (h3 << 16));
          if (p.split) {
            const uint32_t l0 = bf16_bits_rn(v[i].x - __uint_as_float(h0 << 16)), l1 = bf16_bits_rn(v[i].y - __uint_as_float(h1 << 16));
            const uint32_t l2 = bf16_bits_rn(v[i].z - __uint_as_float(h2 << 16)), l3 = bf16_bits_rn(v[i].w - __uint_as_float(h3 << 16));
            *reinterpret_cast<uint2*>(t_lo + off) = make_uint2(l0 | (l1 << 16), l2 | (l3 << 16));
          }
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&a_ready[sa]);
    };
    // three register buffers rotate: k-blocks i+1 and i+2 are in flight while k-block i is converted
    float4 v0[ST_F4], v1[ST_F4], v2[ST_F4];
    load_kb(0, v0);
    if (kblocks > 1) load_kb(1, v1);
    for (int kb = 0; kb < kblocks; kb += 3) {
      if (kb + 2 < kblocks) load_kb(kb + 2, v2);
      store_kb(kb, v0);
      if (kb + 1 < kblocks) {
        if (kb + 3 < kblocks) load_kb(kb + 3, v0);
        store_kb(kb + 1, v1);
      }
      if (kb + 2 < kblocks) {
        if (kb + 4 < kblocks) load_kb(kb + 4, v1);
        store_kb(kb + 2, v2);
      }
    }
    if (warp >= 6) goto workers_done;                // the epilogue needs one warp per TMEM lane quarter: warps 2..5
    // ---- epilogue: thread == output row (TMEM lane)
    const int row = (warp & 3) * 32 + lane;
    const uint32_t tlane = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    mbar_wait(done, 0);
    tc_fence_after();
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 16) {
      uint32_t r[16];
      tmem_ld16(tlane + c0, r);
      tmem_ld_wait();
      if (row < p.M) {
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          const int n = n0 + c0 + j;
          float t[4];
#pragma unroll
          for (int q = 0; q < 4; ++q)
            t[q] = __uint_as_float(r[j + q]) + (p.bias ? __ldg(p.bias + n + q) : 0.f) + p.bias_const;
          const size_t o = (size_t)row * p.ldy + n;
          if (p.gnew) {
            const float4 gn = *reinterpret_cast<const float4*>(p.gnew + o), go = *reinterpret_cast<const float4*>(p.gold + o);
            const float z0 = sigmoid_f(t[0]), z1 = sigmoid_f(t[1]), z2 = sigmoid_f(t[2]), z3 = sigmoid_f(t[3]);
            if (p.gate_z) *reinterpret_cast<float4*>(p.gate_z + o) = make_float4(z0, z1, z2, z3);
            *reinterpret_cast<float4*>(p.Y + o) = make_float4(gn.x * z0 + go.x * (1.f - z0), gn.y * z1 + go.y * (1.f - z1),
                                                              gn.z * z2 + go.z * (1.f - z2), gn.w * z3 + go.w * (1.f - z3));
          } else {
            const float4 y4 = make_float4(apply_act(p.act, t[0]), apply_act(p.act, t[1]), apply_act(p.act, t[2]),
                                          apply_act(p.act, t[3]));
            float* dst = (p.Y2 && n >= p.n_split) ? p.Y2 + (size_t)row * p.ldy + (n - p.n_split) : p.Y + o;
            *reinterpret_cast<float4*>(dst) = y4;
          }
        }
      }
    }
    tc_fence_before();
  }
workers_done:
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, BN);
  }
}

// wt_hi / wt_lo: bf16 [N, K] (K-major) halves of the fp32 weight (mac_pack_weight_bf16_split); wt_lo == NULL: single pass
inline int skinny_tc_launch(SkinnyTcParams p, const void* wt_hi, const void* wt_lo, cudaStream_t stream) {
  if (p.M <= 0 || p.M > 128 || p.N <= 0 || p.K <= 0 || p.nseg < 1 || p.nseg > 4) return MAC_ERR_INVALID;
  if ((p.N % 32) || (p.K % TC_BK) || (p.ldy & 3) || (p.Y2 && (p.n_split % 32))) return MAC_ERR_UNSUPPORTED;
  int ksum = 0;
  for (int i = 0; i < p.nseg; ++i) {
    if (!p.a[i] || p.ak[i] <= 0 || (p.ak[i] % TC_BK) || (p.lda[i] & 3)) return MAC_ERR_UNSUPPORTED;
    if (!mac_aligned16(p.a[i])) return MAC_ERR_ALIGN;
    ksum += p.ak[i];
  }
  if (ksum != p.K || !wt_hi || !p.Y || !mac_aligned16(p.Y) || !mac_aligned16(wt_hi)) return MAC_ERR_INVALID;
  p.split = wt_lo ? 1 : 0;
  const int BN = (p.N % 64 == 0 && p.N >= 1024) ? 64 : 32;
  CUtensorMap mhi, mlo;
  int st = make_tmap_2d(&mhi, wt_hi, 1, (uint64_t)p.N, (uint64_t)p.K, (uint64_t)p.K * 2, (uint32_t)BN, TC_BK, 1);
  if (st != MAC_OK) return st;
  if (wt_lo) {
    st = make_tmap_2d(&mlo, wt_lo, 1, (uint64_t)p.N, (uint64_t)p.K, (uint64_t)p.K * 2, (uint32_t)BN, TC_BK, 1);
    if (st != MAC_OK) return st;
  } else {
    mlo = mhi;
  }
  static bool attr_set[2] = {false, false};
  if (BN == 64) {
    auto kern = skinny_tc_kernel<64>;
    if (!attr_set[0]) {
      MAC_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, StCfg<64>::SMEM_BYTES));
      attr_set[0] = true;
    }
    kern<<<p.N / 64, ST_THREADS, StCfg<64>::SMEM_BYTES, stream>>>(mhi, mlo, p);
  } else {
    auto kern = skinny_tc_kernel<32>;
    if (!attr_set[1]) {
      MAC_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, StCfg<32>::SMEM_BYTES));
      attr_set[1] = true;
    }
    kern<<<p.N / 32, ST_THREADS, StCfg<32>::SMEM_BYTES, stream>>>(mhi, mlo, p);
  }
  MAC_LAUNCH_CHECK();
  return MAC_OK;
}

// fp32 [K, N] (in, out) weight -> bf16 hi and lo halves, both [N, K] (out, in): W = hi + lo + O(2^-17 |W|)
__global__ void pack_weight_bf16_split_kernel(const float* __restrict__ W, __nv_bfloat16* __restrict__ hi,
                                              __nv_bfloat16* __restrict__ lo, int K, int N) {
  __shared__ float tile[32][33];
  const int k0 = blockIdx.y * 32, n0 = blockIdx.x * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int k = k0 + i, n = n0 + threadIdx.x;
    tile[i][threadIdx.x] = (k < K && n < N) ? W[(size_t)k * N + n] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int n = n0 + i, k = k0 + threadIdx.x;
    if (n < N && k < K) {
      const float w = tile[threadIdx.x][i];
      const __nv_bfloat16 h = __float2bfloat16_rn(w);
      hi[(size_t)n * K + k] = h;
      lo[(size_t)n * K + k] = __float2bfloat16_rn(w - __bfloat162float(h));
    }
  }
}

}  // namespace mac
