// bf16 tensor-core GEMM for sm_100a: tcgen05.mma (UMMA 128 x BN x 16, cta_group::1) with fp32 accumulators in
// TMEM, operands staged by TMA (cp.async.bulk.tensor, 128-byte swizzle) through a 4-stage mbarrier ring, and the
// read unit's elementwise work fused into the TMEM->register epilogue.
//
//   C[M,N] = epilogue( A[M,K] @ Wt[N,K]^T )      A, Wt bf16 row-major with K contiguous ("K-major" both)
//
// Warp roles (192 threads, persistent over output tiles, one CTA per SM):
//   warp 0      TMA producer          (one elected lane issues the A and B boxes of each k-block)
//   warp 1      TMEM allocator + MMA issuer (one elected lane issues tcgen05.mma / tcgen05.commit)
//   warps 2..9  epilogue              (tcgen05.ld 32 lanes x 32 columns per instruction; thread == output row;
//                                      two warps per TMEM lane quarter, each takes half of the tile's columns)
// Two accumulator buffers (2 x BN TMEM columns) let the epilogue of tile i overlap the MMAs of tile i+1.
//
// Epilogues (the read-unit chain of mac_cell.py:230-266 / ops.py:668-725, see mac_b200.h):
//   TC_EPI_P       P = acc + bx            -> bf16 P and bf16 P*y[b]            (ops.py:688, 694-703)
//   TC_EPI_ACT     act(acc + b)            -> bf16                               (mac_cell.py:236-238)
//   TC_EPI_LOGITS  I1 = acc + bm2; t = ELU(I1 * control[b]); (dropout); parts[m, ntile] = sum_n t * wr[n]
//                                                                                (ops.py:325-328, mac_cell.py:248-266)
//   TC_EPI_F32     act(acc + b)            -> fp32                               (generic ops.linear)
//   TC_EPI_ADDACT  act(acc + b + add[m,n]) -> bf16, add = bf16 [M, N]           (eval-mode read: step-invariant half of
//                                                                                 the memKbProj concat, mac_cell.py:236-238)
//   TC_EPI_ACT_SPLIT  x = act(acc + b + addf[m,n]) (addf fp32, optional) -> bf16 hi at out0[m, n] and bf16 lo = x - hi at
//                     out0[m, N + n] (ldo = 2N): the A operand of the next split-bf16 ("tc32") product
#pragma once
#include "common.cuh"
#include "tmap.cuh"
#include <stdlib.h>

namespace mac {

enum { TC_EPI_P = 0, TC_EPI_ACT = 1, TC_EPI_LOGITS = 2, TC_EPI_F32 = 3, TC_EPI_ADDACT = 4, TC_EPI_ACT_SPLIT = 5 };

struct TcGemmParams {
  int M, N, K;
  int kblocks0;            // k-blocks (of 64) served by tensor map a0; the rest come from a1 (concat along K)
  int epi, act;
  const float* bias;       // [N] or NULL
  __nv_bfloat16* out0;     // bf16 output 0 (P / act / I1-save), may be NULL for TC_EPI_LOGITS
  __nv_bfloat16* out1;     // bf16 output 1 (P*y)
  float* outf;             // fp32 output (TC_EPI_F32)
  const __nv_bfloat16* add;   // TC_EPI_ADDACT: pre-activation addend [M, ldo]
  const float* addf;          // TC_EPI_ACT_SPLIT: optional fp32 pre-activation addend [M, ldaf]
  int ldaf;
  int ldo;
  const float* y;          // [B, N] row scale for TC_EPI_P
  const float* ctrl;       // [B, N] for TC_EPI_LOGITS
  const float* wr;         // [N]
  float* parts;            // [M, gridN]
  int rows_per_batch;
  uint32_t e_thresh;       // dropout on the logits' input (0 = none)
  float e_scale;
  uint64_t seed;
  int e_site, step;
  int ksplit;              // split-K (TC_EPI_F32 only, single-CTA kernel): the K range is cut into `ksplit` equal slices, each an
  long long split_stride;  //   extra "tile" writing its fp32 partial to outf + slice * split_stride; 0 / 1 = off
  int debug;               // MAC_TC_DEBUG (profiling experiments only): 1 skip epilogue math/stores, 2 skip MMA, 4 skip TMA
};

// ------------------------------------------------------------------ tcgen05 PTX wrappers
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]; single-thread issue
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
// mbarrier arrive when all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor (cute::UMMA::SmemDescriptor) for a K-major tile stored as rows of 128 bytes with
// the 128-byte swizzle (what TMA SWIZZLE_128B writes for a [rows x 64 bf16] box):
//   start address >> 4 | LBO (ignored for swizzled K-major; 1) | SBO = 1024 B between 8-row groups | version 1 |
//   layout type 2 (SWIZZLE_128B).  The tile base must be 1024-byte aligned (base_offset 0).
__device__ __forceinline__ uint64_t make_sw128_kmajor_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);   // bits [0,14)
  d |= (uint64_t)1 << 16;                        // leading byte offset (16-B units), bits [16,30)
  d |= (uint64_t)(1024 >> 4) << 32;              // stride byte offset, bits [32,46)
  d |= (uint64_t)1 << 46;                        // version = 1 (Blackwell), bits [46,48)
  d |= (uint64_t)2 << 61;                        // layout type SWIZZLE_128B, bits [61,64)
  return d;
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): D fp32, A/B bf16, both K-major, dense, M x N
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

constexpr int TC_BK = 64;                 // 64 bf16 = 128 B = one swizzle atom row
constexpr int TC_MAX_VEC = 4;              // a 128-row tile spans at most this many samples on the smem-parameter path
constexpr int TC_EPI_WARPS = 8;            // cta_group::2 kernel: two per TMEM lane quarter
constexpr int TC_THREADS = 64 + 32 * TC_EPI_WARPS;
constexpr int TC_EW = 16;                  // single-CTA kernel: four epilogue warps per TMEM lane quarter
constexpr int TC_THREADS1 = 64 + 32 * TC_EW;

// Tile shapes.  The measured per-SM TMA fill rate is ~64 B/clk (profiles/r1/NOTES.md): a k-block of a BM x BN tile
// brings (BM + BN) * 128 B and feeds BM * BN / 64 MMA cycles, so
//   128 x 128 -> 128 B/clk needed (2x fill-bound), 128 x 256 -> 96 B/clk (1.5x), 256 x 256 -> 64 B/clk (balanced).
// BM = 256 is two M=128 UMMAs per k-step that share the B tile in shared memory and own one accumulator each
// (2 x 256 TMEM columns, not double-buffered); BM = 128 keeps two accumulator buffers so the epilogue of tile i
// overlaps the MMAs of tile i+1.
template <int BM, int BN>
struct TcCfg {
  static constexpr int NMS = BM / 128;                      // M sub-tiles (UMMA M = 128 each)
  static constexpr int STAGES = BM == 256 ? 3 : (BN == 256 ? 4 : 6);
  static constexpr int A_BYTES = BM * TC_BK * 2;            // 16 / 32 KB
  static constexpr int B_BYTES = BN * TC_BK * 2;            // 16 / 32 KB
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int NBUF = BM == 256 ? 1 : 2;            // accumulator buffers
  static constexpr int BUF_COLS = NMS * BN;
  static constexpr int TMEM_COLS = 512 / (BM == 128 && BN == 128 ? 2 : 1);   // NBUF * BUF_COLS, a power of two
  static constexpr int STG_WORDS = 32 * 8;                  // per epilogue warp: 32 rows x 8 words (32 B), XOR-swizzled
  static constexpr int PAR_ROWS = 2 + TC_MAX_VEC;           // bias, wr, and up to TC_MAX_VEC per-sample rows (y / control)
  static constexpr int PAR_WORDS = NBUF * PAR_ROWS * BN;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/ +
                                    TC_EW * STG_WORDS * 4 + PAR_WORDS * 4;
};

// fast ELU for the tensor-core path: x > 0 ? x : exp(x) - 1 with the SFU exponential (abs error ~1e-7 near 0,
// far below the bf16 rounding of the stored activations)
// (ex2.approx.ftz directly: without -ftz, __expf expands to a ~10-instruction denormal-safe sequence per element, which
// made the 128 x 512 epilogues of the fused read step issue-bound)
__device__ __forceinline__ float ex2_ftz(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float elu_fast(float x) {
  const float e = ex2_ftz(x * 1.4426950408889634f) - 1.f;
  return x > 0.f ? x : e;
}
template <int ACT>
__device__ __forceinline__ float act_ct(float x) {
  if constexpr (ACT == MAC_ACT_TANH) return tanhf(x);
  else if constexpr (ACT == MAC_ACT_SIGMOID) return 1.f / (1.f + __expf(-x));
  else if constexpr (ACT == MAC_ACT_ELU) return elu_fast(x);
  else if constexpr (ACT == MAC_ACT_RELU) return fmaxf(x, 0.f);
  else return x;
}
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}

template <int BM, int BN, int EPI, int ACT>
__global__ void __launch_bounds__(TC_THREADS1, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap map_a0, const __grid_constant__ CUtensorMap map_a1,
               const __grid_constant__ CUtensorMap map_b, const TcGemmParams p) {
  using C = TcCfg<BM, BN>;
  constexpr int TC_BM = BM;
  extern __shared__ unsigned char smem_dyn[];
  // 1024-byte aligned operand ring
  const uint32_t base_u32 = smem_u32(smem_dyn);
  const uint32_t pad = (1024u - (base_u32 & 1023u)) & 1023u;
  unsigned char* tiles = smem_dyn + pad;
  uint64_t* bars = reinterpret_cast<uint64_t*>(tiles + C::STAGES * C::STAGE_BYTES);
  uint64_t* full = bars;                       // [STAGES]  TMA -> MMA
  uint64_t* empty = bars + C::STAGES;          // [STAGES]  MMA -> TMA
  uint64_t* tfull = bars + 2 * C::STAGES;      // [2]       MMA -> epilogue
  uint64_t* tempty = tfull + 2;                // [2]       epilogue -> MMA
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty + 2);
  uint32_t* stg_all = reinterpret_cast<uint32_t*>(reinterpret_cast<unsigned char*>(bars) + 256);
  float* par_all = reinterpret_cast<float*>(stg_all + TC_EW * C::STG_WORDS);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_tiles = (p.M + TC_BM - 1) / TC_BM;
  const int n_tiles = p.N / BN;
  const int ksplit = p.ksplit > 1 ? p.ksplit : 1;
  const int mn_tiles = m_tiles * n_tiles;
  const int num_tiles = mn_tiles * ksplit;                 // tile t = (K slice t / mn_tiles, output tile t % mn_tiles)
  const int kblocks = p.K / TC_BK / ksplit;                // k-blocks per slice

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a0);
    tma_prefetch_desc(&map_a1);
    tma_prefetch_desc(&map_b);
#pragma unroll
    for (int i = 0; i < C::STAGES; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    mbar_init(&tfull[0], 1);
    mbar_init(&tfull[1], 1);
    mbar_init(&tempty[0], TC_EW);
    mbar_init(&tempty[1], TC_EW);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr, C::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ===================================================== TMA producer
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        const int tt = t % mn_tiles, kb0 = (t / mn_tiles) * kblocks;
        const int mt = tt / n_tiles, nt = tt % n_tiles;
        for (int kb = kb0; kb < kb0 + kblocks; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          unsigned char* sa = tiles + stage * C::STAGE_BYTES;
          unsigned char* sb = sa + C::A_BYTES;
          if (p.debug & 4) {
            mbar_arrive(&full[stage]);
            if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
            continue;
          }
          mbar_expect_tx(&full[stage], C::STAGE_BYTES);
          if (kb < p.kblocks0)
            tma_load_2d(sa, &map_a0, kb * TC_BK, mt * TC_BM, &full[stage]);
          else
            tma_load_2d(sa, &map_a1, (kb - p.kblocks0) * TC_BK, mt * TC_BM, &full[stage]);
          tma_load_2d(sb, &map_b, kb * TC_BK, nt * BN, &full[stage]);
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer
    constexpr uint32_t idesc = make_idesc_bf16(128, BN);
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
      const int acc = it % C::NBUF;
      const uint32_t acc_phase = (it / C::NBUF) & 1;
      mbar_wait(&tempty[acc], acc_phase ^ 1);        // epilogue has drained this accumulator buffer
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + acc * C::BUF_COLS;
      for (int kb = 0; kb < kblocks; ++kb) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t sa = smem_u32(tiles + stage * C::STAGE_BYTES);
          const uint64_t adesc = make_sw128_kmajor_desc(sa);
          const uint64_t bdesc = make_sw128_kmajor_desc(sa + C::A_BYTES);
#pragma unroll
          for (int k = 0; k < TC_BK / 16; ++k) {
            // advance 16 bf16 = 32 bytes along K inside the 128-byte swizzle atom: +2 in the 16-B address field;
            // the second M sub-tile's rows start 16 KB (= 1024 in 16-B units) further
#pragma unroll
            for (int ms = 0; ms < C::NMS; ++ms)
              if (!(p.debug & 2))
                umma_bf16(tmem_d + ms * BN, adesc + 2 * k + ms * 1024, bdesc + 2 * k, idesc, (kb | k) ? 1u : 0u);
          }
          umma_commit(&empty[stage]);                // frees the smem slot when these MMAs retire
          if (kb == kblocks - 1) umma_commit(&tfull[acc]);
        }
        __syncwarp();
        if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    // ===================================================== epilogue: TC_EW = 16 warps, four per TMEM lane quarter
    // thread == one output row of the tile (TMEM lane); the warp walks its share of the tile's columns 16 at a time
    // (one tcgen05.ld 32x32b.x16).  With two warps per scheduler (the 8-warp version) the ELU/pack/store chain was
    // issue-latency bound and its ~4 us per 128x256 tile was the largest non-MMA term of the K=512 projections;
    // four warps per scheduler overlap those chains.
    const int q = warp & 3;
    const int ew = warp - 2;
    uint32_t* stg = stg_all + ew * C::STG_WORDS;    // 1 KB patch: 32 rows x 32 B
    const int grp = ew >> 2;                        // 0..3
    const int ms = C::NMS == 2 ? (grp >> 1) : 0;    // BM=256: which M sub-tile
    constexpr int CW = C::NMS == 2 ? BN / 2 : BN / 4;                 // columns per warp
    constexpr int NSUB = C::NMS == 2 ? 2 : 4;                         // logit partial sums per (row, n-tile)
    const int sub = C::NMS == 2 ? (grp & 1) : grp;
    const int c_begin = sub * CW, c_end = c_begin + CW;
    // Row-per-thread registers -> coalesced 128-bit global stores through the warp's private patch.  A patch row is
    // 32 B = two 16-B groups; group g of row r lives at slot (g ^ ((r >> 2) & 1)), which makes both the row-wise
    // writes (thread == row) and the read-back (lane l -> group l & 1 of row i*16 + l/2) bank-conflict free.  One
    // store instruction then covers 16 rows x 32 B (whole sectors).
    uint4* s4 = reinterpret_cast<uint4*>(stg);
    const int sw = (lane >> 2) & 1;
    const int rb_g = lane & 1;
    auto store_bf16_chunk = [&](const uint32_t (&w)[8], __nv_bfloat16* out, int row0, int col0) {
      __syncwarp();
      s4[lane * 2 + (0 ^ sw)] = make_uint4(w[0], w[1], w[2], w[3]);
      s4[lane * 2 + (1 ^ sw)] = make_uint4(w[4], w[5], w[6], w[7]);
      __syncwarp();
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int rr = i * 16 + (lane >> 1);
        const uint4 v = s4[rr * 2 + (rb_g ^ ((rr >> 2) & 1))];
        if (row0 + rr < p.M) *reinterpret_cast<uint4*>(out + (size_t)(row0 + rr) * p.ldo + col0 + 8 * rb_g) = v;
      }
    };
    auto store_f32_chunk = [&](const float (&w)[16], float* out, int row0, int col0) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {            // two 8-column halves through the 1 KB patch
        __syncwarp();
        s4[lane * 2 + (0 ^ sw)] = make_uint4(__float_as_uint(w[8 * h]), __float_as_uint(w[8 * h + 1]),
                                             __float_as_uint(w[8 * h + 2]), __float_as_uint(w[8 * h + 3]));
        s4[lane * 2 + (1 ^ sw)] = make_uint4(__float_as_uint(w[8 * h + 4]), __float_as_uint(w[8 * h + 5]),
                                             __float_as_uint(w[8 * h + 6]), __float_as_uint(w[8 * h + 7]));
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const int rr = i * 16 + (lane >> 1);
          const uint4 v = s4[rr * 2 + (rb_g ^ ((rr >> 2) & 1))];
          if (row0 + rr < p.M)
            *reinterpret_cast<uint4*>(out + (size_t)(row0 + rr) * p.ldo + col0 + 8 * h + 4 * rb_g) = v;
        }
      }
    };
    const int etid = threadIdx.x - 64;              // 0..511 among the epilogue threads
    const float* vec_src = (EPI == TC_EPI_P) ? p.y : (EPI == TC_EPI_LOGITS ? p.ctrl : nullptr);
    int it = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
      const int tt = t % mn_tiles;
      const int mt = tt / n_tiles, nt = tt % n_tiles;
      float* const outf_t = p.outf + (size_t)(t / mn_tiles) * (size_t)p.split_stride;      // split-K partial (slice 0: outf)
      const int acc = it % C::NBUF;
      const uint32_t acc_phase = (it / C::NBUF) & 1;
      // ---- stage this tile's parameters in shared memory while the MMAs of the tile are still running:
      //      row 0 bias[BN], row 1 wr[BN], rows 2.. the y / control vectors of the samples this tile touches
      float* par = par_all + acc * (C::PAR_ROWS * BN);
      const int b_lo = (mt * TC_BM) / p.rows_per_batch;
      const int last_row = min(p.M, (mt + 1) * TC_BM) - 1;
      const int nvec = vec_src ? (last_row / p.rows_per_batch - b_lo + 1) : 0;
      const bool vec_smem = nvec <= TC_MAX_VEC;
      {
        const int nb = nt * BN;
        for (int i = etid; i < BN; i += 32 * TC_EW) {
          par[i] = p.bias ? __ldg(p.bias + nb + i) : 0.f;
          if constexpr (EPI == TC_EPI_LOGITS) par[BN + i] = __ldg(p.wr + nb + i);
        }
        if (vec_src && vec_smem)
          for (int i = etid; i < nvec * BN; i += 32 * TC_EW)
            par[2 * BN + i] = __ldg(vec_src + (size_t)(b_lo + i / BN) * p.N + nb + (i % BN));
        asm volatile("bar.sync 1, %0;" ::"n"(32 * TC_EW) : "memory");
      }
      const int row0 = mt * TC_BM + ms * 128 + q * 32;
      const int row = row0 + lane;
      const bool row_ok = row < p.M;
      const int bidx = row_ok ? row / p.rows_per_batch : b_lo;
      const float* vrow = vec_smem ? par + (2 + bidx - b_lo) * BN : nullptr;   // this row's y / control vector (tile-local)
      const uint32_t taddr = tmem_base + acc * C::BUF_COLS + ms * BN + ((uint32_t)(q * 32) << 16);
      // TC_EPI_ADDACT: the bf16 addend of up to ADD_WIN chunks, in the coalesced (read-back) mapping.  It does not
      // depend on the accumulator, so the loads are issued before the wait on it and overlap this tile's MMAs
      // (with a one-chunk look-ahead the epilogue paid one DRAM/L2 round trip per chunk: 24.7 -> 21.1 us -> see NOTES).
      constexpr int NCH = CW / 16;
      constexpr int ADD_WIN = (EPI == TC_EPI_ADDACT) ? (NCH < 4 ? NCH : 4) : 1;
      uint4 addq[ADD_WIN][2];
      auto load_addend = [&](int slot, int c0) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const int rr = i * 16 + (lane >> 1);
          addq[slot][i] = (row0 + rr < p.M)
              ? __ldg(reinterpret_cast<const uint4*>(p.add + (size_t)(row0 + rr) * p.ldo + nt * BN + c0 + 8 * rb_g))
              : make_uint4(0u, 0u, 0u, 0u);
        }
      };
      if constexpr (EPI == TC_EPI_ADDACT) {
#pragma unroll
        for (int k = 0; k < ADD_WIN; ++k) load_addend(k, c_begin + 16 * k);
      }
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      float part = 0.f;
      // the ADDACT variant is fully unrolled so that the addend window is indexed statically
#pragma unroll (EPI == TC_EPI_ADDACT ? NCH : 1)
      for (int ch = 0; ch < NCH; ++ch) {
        const int c0 = c_begin + 16 * ch;
        if (p.debug & 1) break;
        uint32_t r[16];
        if (p.debug & 8) {
#pragma unroll
          for (int j = 0; j < 16; ++j) r[j] = (uint32_t)(c0 + j);
        } else {
          tmem_ld16(taddr + c0, r);
          tmem_ld_wait();
        }
        const int n0 = nt * BN + c0;
        const float4* bias4 = reinterpret_cast<const float4*>(par + c0);
        if constexpr (EPI == TC_EPI_P) {
          const float4* y4 = vec_smem ? reinterpret_cast<const float4*>(vrow + c0)
                                      : reinterpret_cast<const float4*>(p.y + (size_t)bidx * p.N + n0);
          uint32_t w0[8], w1[8];
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            const float4 b0 = bias4[j / 4], y0 = y4[j / 4];
            const float x0 = __uint_as_float(r[j]) + b0.x, x1 = __uint_as_float(r[j + 1]) + b0.y;
            const float x2 = __uint_as_float(r[j + 2]) + b0.z, x3 = __uint_as_float(r[j + 3]) + b0.w;
            w0[j / 2] = pack_bf16(x0, x1);
            w0[j / 2 + 1] = pack_bf16(x2, x3);
            w1[j / 2] = pack_bf16(x0 * y0.x, x1 * y0.y);
            w1[j / 2 + 1] = pack_bf16(x2 * y0.z, x3 * y0.w);
          }
          store_bf16_chunk(w0, p.out0, row0, n0);
          store_bf16_chunk(w1, p.out1, row0, n0);
        } else if constexpr (EPI == TC_EPI_ACT) {
          uint32_t w0[8];
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            const float4 b0 = bias4[j / 4];
            w0[j / 2] = pack_bf16(act_ct<ACT>(__uint_as_float(r[j]) + b0.x), act_ct<ACT>(__uint_as_float(r[j + 1]) + b0.y));
            w0[j / 2 + 1] = pack_bf16(act_ct<ACT>(__uint_as_float(r[j + 2]) + b0.z), act_ct<ACT>(__uint_as_float(r[j + 3]) + b0.w));
          }
          store_bf16_chunk(w0, p.out0, row0, n0);
        } else if constexpr (EPI == TC_EPI_ADDACT) {
          // addend chunk -> patch (coalesced mapping) -> own row back into registers
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            const int rr = i * 16 + (lane >> 1);
            s4[rr * 2 + (rb_g ^ ((rr >> 2) & 1))] = addq[ch % ADD_WIN][i];
          }
          __syncwarp();
          const uint4 qa = s4[lane * 2 + (0 ^ sw)], qb = s4[lane * 2 + (1 ^ sw)];
          if (ch + ADD_WIN < NCH) load_addend(ch % ADD_WIN, c0 + 16 * ADD_WIN);   // refill the slot just consumed
          const uint32_t qw[8] = {qa.x, qa.y, qa.z, qa.w, qb.x, qb.y, qb.z, qb.w};
          uint32_t w0[8];
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            const float4 b0 = bias4[j / 4];
            // bf16 -> fp32 is a 16-bit shift
            const float q0 = __uint_as_float(qw[j / 2] << 16), q1 = __uint_as_float(qw[j / 2] & 0xffff0000u);
            const float q2 = __uint_as_float(qw[j / 2 + 1] << 16), q3 = __uint_as_float(qw[j / 2 + 1] & 0xffff0000u);
            w0[j / 2] = pack_bf16(act_ct<ACT>(__uint_as_float(r[j]) + b0.x + q0), act_ct<ACT>(__uint_as_float(r[j + 1]) + b0.y + q1));
            w0[j / 2 + 1] = pack_bf16(act_ct<ACT>(__uint_as_float(r[j + 2]) + b0.z + q2), act_ct<ACT>(__uint_as_float(r[j + 3]) + b0.w + q3));
          }
          store_bf16_chunk(w0, p.out0, row0, n0);
        } else if constexpr (EPI == TC_EPI_ACT_SPLIT) {
          const float* arow = p.addf ? p.addf + (size_t)(row_ok ? row : 0) * p.ldaf + n0 : nullptr;
          uint32_t wh[8], wl[8];
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            const float4 b0 = bias4[j / 4];
            const float4 a0 = arow ? *reinterpret_cast<const float4*>(arow + j) : make_float4(0.f, 0.f, 0.f, 0.f);
            const float x0 = act_ct<ACT>(__uint_as_float(r[j]) + b0.x + a0.x), x1 = act_ct<ACT>(__uint_as_float(r[j + 1]) + b0.y + a0.y);
            const float x2 = act_ct<ACT>(__uint_as_float(r[j + 2]) + b0.z + a0.z), x3 = act_ct<ACT>(__uint_as_float(r[j + 3]) + b0.w + a0.w);
            const uint32_t h01 = pack_bf16(x0, x1), h23 = pack_bf16(x2, x3);
            wh[j / 2] = h01;
            wh[j / 2 + 1] = h23;
            wl[j / 2] = pack_bf16(x0 - __uint_as_float(h01 << 16), x1 - __uint_as_float(h01 & 0xffff0000u));
            wl[j / 2 + 1] = pack_bf16(x2 - __uint_as_float(h23 << 16), x3 - __uint_as_float(h23 & 0xffff0000u));
          }
          store_bf16_chunk(wh, p.out0, row0, n0);
          store_bf16_chunk(wl, p.out0, row0, p.N + n0);
        } else if constexpr (EPI == TC_EPI_F32) {
          float w0[16];
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            const float4 b0 = bias4[j / 4];
            w0[j] = act_ct<ACT>(__uint_as_float(r[j]) + b0.x);
            w0[j + 1] = act_ct<ACT>(__uint_as_float(r[j + 1]) + b0.y);
            w0[j + 2] = act_ct<ACT>(__uint_as_float(r[j + 2]) + b0.z);
            w0[j + 3] = act_ct<ACT>(__uint_as_float(r[j + 3]) + b0.w);
          }
          store_f32_chunk(w0, outf_t, row0, n0);
        } else {  // TC_EPI_LOGITS
          const float4* c4 = vec_smem ? reinterpret_cast<const float4*>(vrow + c0)
                                      : reinterpret_cast<const float4*>(p.ctrl + (size_t)bidx * p.N + n0);
          const float4* w4 = reinterpret_cast<const float4*>(par + BN + c0);
          uint32_t w0[8];
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            const float4 b0 = bias4[j / 4], cc = c4[j / 4], ww = w4[j / 4];
            const float i0 = __uint_as_float(r[j]) + b0.x, i1 = __uint_as_float(r[j + 1]) + b0.y;
            const float i2 = __uint_as_float(r[j + 2]) + b0.z, i3 = __uint_as_float(r[j + 3]) + b0.w;
            w0[j / 2] = pack_bf16(i0, i1);
            w0[j / 2 + 1] = pack_bf16(i2, i3);
            float t0 = elu_fast(i0 * cc.x), t1 = elu_fast(i1 * cc.y), t2 = elu_fast(i2 * cc.z), t3 = elu_fast(i3 * cc.w);
            if (p.e_thresh) {
              const uint64_t e = (uint64_t)row * (uint64_t)p.N + (uint64_t)(n0 + j);
              const Philox4 rr = philox4x32_10(p.seed, e >> 2, (uint32_t)p.e_site, (uint32_t)p.step);
              t0 = ((rr.x >> 8) >= p.e_thresh) ? t0 * p.e_scale : 0.f;
              t1 = ((rr.y >> 8) >= p.e_thresh) ? t1 * p.e_scale : 0.f;
              t2 = ((rr.z >> 8) >= p.e_thresh) ? t2 * p.e_scale : 0.f;
              t3 = ((rr.w >> 8) >= p.e_thresh) ? t3 * p.e_scale : 0.f;
            }
            part = fmaf(t0, ww.x, part);
            part = fmaf(t1, ww.y, part);
            part = fmaf(t2, ww.z, part);
            part = fmaf(t3, ww.w, part);
          }
          if (p.out0) store_bf16_chunk(w0, p.out0, row0, n0);       // I1 kept for backward
        }
      }
      if constexpr (EPI == TC_EPI_LOGITS) {
        // NSUB partial sums per (row, n-tile): one per epilogue warp that shares the row
        if (row_ok) p.parts[((size_t)row * n_tiles + nt) * NSUB + sub] = part;
      }
      // release the accumulator buffer to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
    }
  }

  // ---- teardown
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

// =====================================================================================================
// cta_group::2 variant: a CTA PAIR (2-CTA cluster, one TPC) computes a 256 x 256 tile with ONE tcgen05.mma issued by
// the leader CTA.  Each CTA stages its own 128 rows of A and only HALF of the B tile (128 of the 256 W rows) and keeps
// its 128 x 256 half of the accumulator in its own TMEM, so per MMA each SM's shared memory supplies 8 KB instead of
// 12 KB -- the measured SS-mode operand-read limit (~64 B/clk/SM) stops throttling the tensor core.
//   * TMA loads are issued by both CTAs into their own smem but complete on the LEADER's `full` barrier
//     (cp.async.bulk.tensor ... .cta_group::2, barrier address with the peer bit cleared).
//   * the leader's tcgen05.commit multicasts to both CTAs' `empty` / `tfull` barriers.
//   * each CTA's epilogue drains its own TMEM half; every epilogue warp of BOTH CTAs arrives on the leader's `tempty`.
// =====================================================================================================
__device__ __forceinline__ void tmem_alloc2(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma2_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void umma2_commit_mc(uint64_t* bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(cta_mask)
               : "memory");
}
// TMA load into this CTA's smem, transaction bytes counted on the LEADER CTA's barrier (peer bit 24 cleared)
__device__ __forceinline__ void tma2_load_2d(void* smem_dst, const void* tmap, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
          smem_u32(smem_dst)),
      "l"(tmap), "r"(c0), "r"(c1), "r"(smem_u32(bar) & 0xFEFFFFFFu)
      : "memory");
}
// arrive on the barrier at the same offset in CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t cta) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}"
      ::"r"(smem_u32(bar)), "r"(cta)
      : "memory");
}
__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_barrier() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

struct Tc2Cfg {
  static constexpr int BN = 256;
  static constexpr int STAGES = 5;
  static constexpr int A_BYTES = 128 * TC_BK * 2;            // this CTA's 128 rows of A
  static constexpr int B_BYTES = (BN / 2) * TC_BK * 2;       // this CTA's half of the B tile
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;      // 32 KB per CTA per k-block
  static constexpr int TMEM_COLS = 512;                      // two 256-column accumulator buffers
  static constexpr int STG_WORDS = 32 * 16;
  static constexpr int PAR_ROWS = 2 + TC_MAX_VEC;
  static constexpr int PAR_WORDS = 2 * PAR_ROWS * BN;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256 + TC_EPI_WARPS * STG_WORDS * 4 + PAR_WORDS * 4;
};

template <int EPI, int ACT>
__global__ void __launch_bounds__(TC_THREADS, 1)
tc2_gemm_kernel(const __grid_constant__ CUtensorMap map_a0, const __grid_constant__ CUtensorMap map_a1,
                const __grid_constant__ CUtensorMap map_b, const TcGemmParams p) {
  using C = Tc2Cfg;
  constexpr int BN = C::BN;
  extern __shared__ unsigned char smem_dyn[];
  const uint32_t base_u32 = smem_u32(smem_dyn);
  const uint32_t pad = (1024u - (base_u32 & 1023u)) & 1023u;
  unsigned char* tiles = smem_dyn + pad;
  uint64_t* bars = reinterpret_cast<uint64_t*>(tiles + C::STAGES * C::STAGE_BYTES);
  uint64_t* full = bars;                       // [STAGES]  (leader's are the live ones)
  uint64_t* empty = bars + C::STAGES;          // [STAGES]  per CTA
  uint64_t* tfull = bars + 2 * C::STAGES;      // [2]       per CTA
  uint64_t* tempty = tfull + 2;                // [2]       leader's are the live ones
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty + 2);
  uint32_t* stg_all = reinterpret_cast<uint32_t*>(reinterpret_cast<unsigned char*>(bars) + 256);
  float* par_all = reinterpret_cast<float*>(stg_all + TC_EPI_WARPS * C::STG_WORDS);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_rank();                     // 0 = leader
  const int m_pairs = (p.M + 255) / 256;
  const int n_tiles = p.N / BN;
  const int num_tiles = m_pairs * n_tiles;
  const int kblocks = p.K / TC_BK;
  const int pair0 = blockIdx.x >> 1, pair_stride = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a0);
    tma_prefetch_desc(&map_a1);
    tma_prefetch_desc(&map_b);
#pragma unroll
    for (int i = 0; i < C::STAGES; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    mbar_init(&tfull[0], 1);
    mbar_init(&tfull[1], 1);
    mbar_init(&tempty[0], 2 * TC_EPI_WARPS);
    mbar_init(&tempty[1], 2 * TC_EPI_WARPS);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc2(tmem_ptr, C::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  cluster_barrier();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ===================================================== TMA producer (both CTAs)
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = pair0; t < num_tiles; t += pair_stride) {
        const int mp = t / n_tiles, nt = t % n_tiles;
        const int row_a = mp * 256 + (int)rank * 128;
        const int row_b = nt * BN + (int)rank * (BN / 2);
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          unsigned char* sa = tiles + stage * C::STAGE_BYTES;
          unsigned char* sb = sa + C::A_BYTES;
          if (rank == 0) mbar_expect_tx(&full[stage], 2 * C::STAGE_BYTES);   // both CTAs' bytes land on the leader's barrier
          if (kb < p.kblocks0)
            tma2_load_2d(sa, &map_a0, kb * TC_BK, row_a, &full[stage]);
          else
            tma2_load_2d(sa, &map_a1, (kb - p.kblocks0) * TC_BK, row_a, &full[stage]);
          tma2_load_2d(sb, &map_b, kb * TC_BK, row_b, &full[stage]);
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer (leader CTA only)
    if (rank == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(256, BN);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int t = pair0; t < num_tiles; t += pair_stride, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&tempty[acc], acc_phase ^ 1);      // both CTAs' epilogues have drained this buffer
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BN;
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t sa = smem_u32(tiles + stage * C::STAGE_BYTES);
            const uint64_t adesc = make_sw128_kmajor_desc(sa);
            const uint64_t bdesc = make_sw128_kmajor_desc(sa + C::A_BYTES);
#pragma unroll
            for (int k = 0; k < TC_BK / 16; ++k) umma2_bf16(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) ? 1u : 0u);
            umma2_commit_mc(&empty[stage], 0x3);     // both CTAs may refill this slot
            if (kb == kblocks - 1) umma2_commit_mc(&tfull[acc], 0x3);
          }
          __syncwarp();
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else {
    // ===================================================== epilogue (both CTAs; own 128 rows)
    const int q = warp & 3;
    uint32_t* stg = stg_all + (warp - 2) * C::STG_WORDS;
    const int grp = (warp - 2) >> 2;
    const int c_begin = grp * (BN / 2), c_end = (grp + 1) * (BN / 2);
    auto store_bf16_chunk = [&](const uint32_t (&w)[16], __nv_bfloat16* out, int row0, int col0) {
      __syncwarp();
      uint4* s4 = reinterpret_cast<uint4*>(stg);
#pragma unroll
      for (int g = 0; g < 4; ++g) s4[lane * 4 + (g ^ (lane & 3))] = make_uint4(w[4 * g], w[4 * g + 1], w[4 * g + 2], w[4 * g + 3]);
      __syncwarp();
      const int g = lane & 3;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int rr = i * 8 + (lane >> 2);
        const uint4 v = s4[rr * 4 + (g ^ (rr & 3))];
        if (row0 + rr < p.M) *reinterpret_cast<uint4*>(out + (size_t)(row0 + rr) * p.ldo + col0 + 8 * g) = v;
      }
    };
    auto store_f32_chunk = [&](const float (&w)[32], float* out, int row0, int col0) {
      float4* s4 = reinterpret_cast<float4*>(stg);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        __syncwarp();
#pragma unroll
        for (int g = 0; g < 4; ++g)
          s4[lane * 4 + (g ^ (lane & 3))] = make_float4(w[16 * h + 4 * g], w[16 * h + 4 * g + 1], w[16 * h + 4 * g + 2], w[16 * h + 4 * g + 3]);
        __syncwarp();
        const int g = lane & 3;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int rr = i * 8 + (lane >> 2);
          const float4 v = s4[rr * 4 + (g ^ (rr & 3))];
          if (row0 + rr < p.M) *reinterpret_cast<float4*>(out + (size_t)(row0 + rr) * p.ldo + col0 + 16 * h + 4 * g) = v;
        }
      }
    };
    const int etid = threadIdx.x - 64;
    const float* vec_src = (EPI == TC_EPI_P) ? p.y : (EPI == TC_EPI_LOGITS ? p.ctrl : nullptr);
    int it = 0;
    for (int t = pair0; t < num_tiles; t += pair_stride, ++it) {
      const int mp = t / n_tiles, nt = t % n_tiles;
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int tile_row0 = mp * 256 + (int)rank * 128;
      float* par = par_all + acc * (C::PAR_ROWS * BN);
      const int b_lo = min(tile_row0, p.M - 1) / p.rows_per_batch;
      const int last_row = max(min(p.M, tile_row0 + 128) - 1, min(tile_row0, p.M - 1));
      const int nvec = vec_src ? (last_row / p.rows_per_batch - b_lo + 1) : 0;
      const bool vec_smem = nvec <= TC_MAX_VEC;
      {
        const int nb = nt * BN;
        for (int i = etid; i < BN; i += 32 * TC_EPI_WARPS) {
          par[i] = p.bias ? __ldg(p.bias + nb + i) : 0.f;
          if constexpr (EPI == TC_EPI_LOGITS) par[BN + i] = __ldg(p.wr + nb + i);
        }
        if (vec_src && vec_smem)
          for (int i = etid; i < nvec * BN; i += 32 * TC_EPI_WARPS)
            par[2 * BN + i] = __ldg(vec_src + (size_t)(b_lo + i / BN) * p.N + nb + (i % BN));
        asm volatile("bar.sync 1, %0;" ::"n"(32 * TC_EPI_WARPS) : "memory");
      }
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      const int row0 = tile_row0 + q * 32;
      const int row = row0 + lane;
      const bool row_ok = row < p.M;
      const int bidx = row_ok ? row / p.rows_per_batch : b_lo;
      const float* vrow = vec_smem ? par + (2 + bidx - b_lo) * BN : nullptr;
      const uint32_t taddr = tmem_base + acc * BN + ((uint32_t)(q * 32) << 16);
      float part = 0.f;
#pragma unroll 1
      for (int c0 = c_begin; c0 < c_end; c0 += 32) {
        uint32_t r[32];
        tmem_ld32(taddr + c0, r);
        tmem_ld_wait();
        const int n0 = nt * BN + c0;
        const float4* bias4 = reinterpret_cast<const float4*>(par + c0);
        if constexpr (EPI == TC_EPI_P) {
          const float4* y4 = vec_smem ? reinterpret_cast<const float4*>(vrow + c0)
                                      : reinterpret_cast<const float4*>(p.y + (size_t)bidx * p.N + n0);
          uint32_t w0[16], w1[16];
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 b0 = bias4[j / 4], y0 = y4[j / 4];
            const float x0 = __uint_as_float(r[j]) + b0.x, x1 = __uint_as_float(r[j + 1]) + b0.y;
            const float x2 = __uint_as_float(r[j + 2]) + b0.z, x3 = __uint_as_float(r[j + 3]) + b0.w;
            w0[j / 2] = pack_bf16(x0, x1);
            w0[j / 2 + 1] = pack_bf16(x2, x3);
            w1[j / 2] = pack_bf16(x0 * y0.x, x1 * y0.y);
            w1[j / 2 + 1] = pack_bf16(x2 * y0.z, x3 * y0.w);
          }
          store_bf16_chunk(w0, p.out0, row0, n0);
          store_bf16_chunk(w1, p.out1, row0, n0);
        } else if constexpr (EPI == TC_EPI_ACT) {
          uint32_t w0[16];
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 b0 = bias4[j / 4];
            w0[j / 2] = pack_bf16(act_ct<ACT>(__uint_as_float(r[j]) + b0.x), act_ct<ACT>(__uint_as_float(r[j + 1]) + b0.y));
            w0[j / 2 + 1] = pack_bf16(act_ct<ACT>(__uint_as_float(r[j + 2]) + b0.z), act_ct<ACT>(__uint_as_float(r[j + 3]) + b0.w));
          }
          store_bf16_chunk(w0, p.out0, row0, n0);
        } else if constexpr (EPI == TC_EPI_F32) {
          float w0[32];
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 b0 = bias4[j / 4];
            w0[j] = act_ct<ACT>(__uint_as_float(r[j]) + b0.x);
            w0[j + 1] = act_ct<ACT>(__uint_as_float(r[j + 1]) + b0.y);
            w0[j + 2] = act_ct<ACT>(__uint_as_float(r[j + 2]) + b0.z);
            w0[j + 3] = act_ct<ACT>(__uint_as_float(r[j + 3]) + b0.w);
          }
          store_f32_chunk(w0, p.outf, row0, n0);
        } else {
          const float4* c4 = vec_smem ? reinterpret_cast<const float4*>(vrow + c0)
                                      : reinterpret_cast<const float4*>(p.ctrl + (size_t)bidx * p.N + n0);
          const float4* w4 = reinterpret_cast<const float4*>(par + BN + c0);
          uint32_t w0[16];
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 b0 = bias4[j / 4], cc = c4[j / 4], ww = w4[j / 4];
            const float i0 = __uint_as_float(r[j]) + b0.x, i1 = __uint_as_float(r[j + 1]) + b0.y;
            const float i2 = __uint_as_float(r[j + 2]) + b0.z, i3 = __uint_as_float(r[j + 3]) + b0.w;
            w0[j / 2] = pack_bf16(i0, i1);
            w0[j / 2 + 1] = pack_bf16(i2, i3);
            float t0 = elu_fast(i0 * cc.x), t1 = elu_fast(i1 * cc.y), t2 = elu_fast(i2 * cc.z), t3 = elu_fast(i3 * cc.w);
            if (p.e_thresh) {
              const uint64_t e = (uint64_t)row * (uint64_t)p.N + (uint64_t)(n0 + j);
              const Philox4 rr = philox4x32_10(p.seed, e >> 2, (uint32_t)p.e_site, (uint32_t)p.step);
              t0 = ((rr.x >> 8) >= p.e_thresh) ? t0 * p.e_scale : 0.f;
              t1 = ((rr.y >> 8) >= p.e_thresh) ? t1 * p.e_scale : 0.f;
              t2 = ((rr.z >> 8) >= p.e_thresh) ? t2 * p.e_scale : 0.f;
              t3 = ((rr.w >> 8) >= p.e_thresh) ? t3 * p.e_scale : 0.f;
            }
            part = fmaf(t0, ww.x, part);
            part = fmaf(t1, ww.y, part);
            part = fmaf(t2, ww.z, part);
            part = fmaf(t3, ww.w, part);
          }
          if (p.out0) store_bf16_chunk(w0, p.out0, row0, n0);
        }
      }
      if constexpr (EPI == TC_EPI_LOGITS) {
        if (row_ok) p.parts[((size_t)row * n_tiles + nt) * 2 + grp] = part;
      }
      // release the accumulator buffer: every epilogue warp of both CTAs arrives on the LEADER's barrier
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (rank == 0) mbar_arrive(&tempty[acc]);
        else mbar_arrive_remote(&tempty[acc], 0);
      }
    }
  }

  // ---- teardown: nobody leaves while the partner may still signal it or read its shared memory
  tc_fence_before();
  __syncthreads();
  cluster_barrier();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, C::TMEM_COLS);
  }
}

// ------------------------------------------------------------------ host side
inline int tc_num_sms() {
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  }
  return sms;
}

template <int BM, int BN, int EPI, int ACT>
inline int tc_gemm_launch_t(const CUtensorMap& ma0, const CUtensorMap& ma1, const CUtensorMap& mb,
                            const TcGemmParams& p, cudaStream_t stream) {
  auto kern = tc_gemm_kernel<BM, BN, EPI, ACT>;
  MAC_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfg<BM, BN>::SMEM_BYTES));
  const int tiles = ((p.M + BM - 1) / BM) * (p.N / BN) * (p.ksplit > 1 ? p.ksplit : 1);
  // persistent grid balanced over the rounds: same makespan as one CTA per SM, but e.g. 196 tiles run as 98 CTAs x 2
  // tiles (epilogue of the first overlaps the MMAs of the second) and leave 50 SMs to concurrent streams
  const int rounds = (tiles + tc_num_sms() - 1) / tc_num_sms();
  const int grid = (tiles + rounds - 1) / rounds;
  kern<<<grid, TC_THREADS1, TcCfg<BM, BN>::SMEM_BYTES, stream>>>(ma0, ma1, mb, p);
  MAC_LAUNCH_CHECK();
  return MAC_OK;
}

template <int BM, int BN>
inline int tc_gemm_dispatch(const CUtensorMap& ma0, const CUtensorMap& ma1, const CUtensorMap& mb,
                            const TcGemmParams& p, cudaStream_t stream) {
  switch (p.epi) {
    case TC_EPI_P: return tc_gemm_launch_t<BM, BN, TC_EPI_P, MAC_ACT_NON>(ma0, ma1, mb, p, stream);
    case TC_EPI_ADDACT:
      if (p.act == MAC_ACT_ELU) return tc_gemm_launch_t<BM, BN, TC_EPI_ADDACT, MAC_ACT_ELU>(ma0, ma1, mb, p, stream);
      return MAC_ERR_UNSUPPORTED;
    case TC_EPI_LOGITS: return tc_gemm_launch_t<BM, BN, TC_EPI_LOGITS, MAC_ACT_NON>(ma0, ma1, mb, p, stream);
    case TC_EPI_ACT_SPLIT:
      if (p.act == MAC_ACT_ELU) return tc_gemm_launch_t<BM, BN, TC_EPI_ACT_SPLIT, MAC_ACT_ELU>(ma0, ma1, mb, p, stream);
      if (p.act == MAC_ACT_NON) return tc_gemm_launch_t<BM, BN, TC_EPI_ACT_SPLIT, MAC_ACT_NON>(ma0, ma1, mb, p, stream);
      return MAC_ERR_UNSUPPORTED;
    case TC_EPI_ACT:
      if (p.act == MAC_ACT_ELU) return tc_gemm_launch_t<BM, BN, TC_EPI_ACT, MAC_ACT_ELU>(ma0, ma1, mb, p, stream);
      if (p.act == MAC_ACT_NON) return tc_gemm_launch_t<BM, BN, TC_EPI_ACT, MAC_ACT_NON>(ma0, ma1, mb, p, stream);
      return MAC_ERR_UNSUPPORTED;
    case TC_EPI_F32:
      switch (p.act) {
        case MAC_ACT_NON: return tc_gemm_launch_t<BM, BN, TC_EPI_F32, MAC_ACT_NON>(ma0, ma1, mb, p, stream);
        case MAC_ACT_TANH: return tc_gemm_launch_t<BM, BN, TC_EPI_F32, MAC_ACT_TANH>(ma0, ma1, mb, p, stream);
        case MAC_ACT_SIGMOID: return tc_gemm_launch_t<BM, BN, TC_EPI_F32, MAC_ACT_SIGMOID>(ma0, ma1, mb, p, stream);
        case MAC_ACT_ELU: return tc_gemm_launch_t<BM, BN, TC_EPI_F32, MAC_ACT_ELU>(ma0, ma1, mb, p, stream);
        case MAC_ACT_RELU: return tc_gemm_launch_t<BM, BN, TC_EPI_F32, MAC_ACT_RELU>(ma0, ma1, mb, p, stream);
      }
      return MAC_ERR_UNSUPPORTED;
  }
  return MAC_ERR_UNSUPPORTED;
}

template <int EPI, int ACT>
inline int tc2_gemm_launch_t(const CUtensorMap& ma0, const CUtensorMap& ma1, const CUtensorMap& mb,
                             const TcGemmParams& p, cudaStream_t stream) {
  auto kern = tc2_gemm_kernel<EPI, ACT>;
  MAC_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Tc2Cfg::SMEM_BYTES));
  const int tiles = ((p.M + 255) / 256) * (p.N / Tc2Cfg::BN);
  const int max_pairs = tc_num_sms() / 2;
  const int rounds = (tiles + max_pairs - 1) / max_pairs;
  const int pairs = (tiles + rounds - 1) / rounds;        // balanced: 98 pair-tiles -> 49 pairs x 2 tiles (98 SMs)
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * pairs, 1, 1);
  cfg.blockDim = dim3(TC_THREADS, 1, 1);
  cfg.dynamicSmemBytes = Tc2Cfg::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  MAC_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, ma0, ma1, mb, p));
  MAC_LAUNCH_CHECK();
  return MAC_OK;
}

inline int tc2_gemm_dispatch(const CUtensorMap& ma0, const CUtensorMap& ma1, const CUtensorMap& mb,
                             const TcGemmParams& p, cudaStream_t stream) {
  switch (p.epi) {
    case TC_EPI_P: return tc2_gemm_launch_t<TC_EPI_P, MAC_ACT_NON>(ma0, ma1, mb, p, stream);
    case TC_EPI_LOGITS: return tc2_gemm_launch_t<TC_EPI_LOGITS, MAC_ACT_NON>(ma0, ma1, mb, p, stream);
    case TC_EPI_ACT:
      if (p.act == MAC_ACT_ELU) return tc2_gemm_launch_t<TC_EPI_ACT, MAC_ACT_ELU>(ma0, ma1, mb, p, stream);
      if (p.act == MAC_ACT_NON) return tc2_gemm_launch_t<TC_EPI_ACT, MAC_ACT_NON>(ma0, ma1, mb, p, stream);
      return MAC_ERR_UNSUPPORTED;
    case TC_EPI_F32:
      if (p.act == MAC_ACT_NON) return tc2_gemm_launch_t<TC_EPI_F32, MAC_ACT_NON>(ma0, ma1, mb, p, stream);
      if (p.act == MAC_ACT_ELU) return tc2_gemm_launch_t<TC_EPI_F32, MAC_ACT_ELU>(ma0, ma1, mb, p, stream);
      if (p.act == MAC_ACT_TANH) return tc2_gemm_launch_t<TC_EPI_F32, MAC_ACT_TANH>(ma0, ma1, mb, p, stream);
      return MAC_ERR_UNSUPPORTED;
  }
  return MAC_ERR_UNSUPPORTED;
}

// Tile shape: minimise rounds-over-the-SMs x measured cycles per k-block of one tile.  SS-mode UMMAs are fed ~55-64 B/clk
// from shared memory (profiles/r1/NOTES.md): a 128x256x16 UMMA takes ~224 cycles, a 128x128x16 one ~149, so per
// k-block: 256x256 -> 1792, 128x256 -> 896 (fill needs 768), 128x128 -> ~620.  Ties go to 128x256: with the balanced
// persistent grid it runs as 2 tiles per CTA and hides the first tile's epilogue under the second tile's MMAs
// (measured 21.2k vs 20.9k (256x256) vs 19.4k (128x128) reasoning-steps/s at the headline shape).  Returns BM*1000 + BN.
inline int tc_pick_tile(int M, int N) {
  const int sms = tc_num_sms();
  auto cost = [&](int bm, int bn, int per_tile) {
    if (N % bn) return 1 << 30;
    const int tiles = ((M + bm - 1) / bm) * (N / bn);
    return ((tiles + sms - 1) / sms) * per_tile;
  };
  int best = 128256, bc = cost(128, 256, 896);
  const int c3 = cost(256, 256, 1792), c1 = cost(128, 128, 620);
  if (c3 < bc) { best = 256256; bc = c3; }
  if (c1 < bc) { best = 128128; bc = c1; }
  return best;
}

// A = [a0 (K0 cols) | a1 (K1 cols)] bf16 row-major (ld = own K), Wt bf16 [N, K0+K1]
inline int tc_gemm_launch(const void* a0, int K0, const void* a1, int K1, const void* wt, TcGemmParams p,
                          cudaStream_t stream, int* nparts_per_row = nullptr, int ldw = 0, int lda0 = 0, int lda1 = 0) {
  if (ldw == 0) ldw = K0 + K1;                  // row pitch of Wt in elements (> K: a column block of a wider weight)
  if (lda0 == 0) lda0 = K0;                     // row pitch of the A segments (> K: a column block of a wider matrix)
  if (lda1 == 0) lda1 = K1;
  if (p.M <= 0 || p.N <= 0 || (p.N % 128) || (K0 % TC_BK) || (K1 % TC_BK) || K0 <= 0) return MAC_ERR_UNSUPPORTED;
  if (!mac_aligned16(a0) || !mac_aligned16(wt)) return MAC_ERR_ALIGN;
  int tile = tc_pick_tile(p.M, p.N);
  if (const char* e = getenv("MAC_TC_TILE")) {
    const int v = atoi(e);
    if ((v == 128128 || v == 128256 || v == 256256) && p.N % (v % 1000) == 0) tile = v;
  }
  if (const char* e = getenv("MAC_TC_DEBUG")) p.debug = atoi(e);
  // cta_group::2 pair kernel (opt-in while it is being qualified): MAC_TC_PAIR=1
  bool pair = false;
  if (const char* e = getenv("MAC_TC_PAIR"))
    pair = atoi(e) != 0 && (p.N % 256 == 0) && p.M > 128 && p.epi != TC_EPI_ADDACT && p.epi != TC_EPI_ACT_SPLIT &&
           lda0 == K0 && lda1 == K1 && p.ksplit <= 1;
  if (pair) {
    if (nparts_per_row) *nparts_per_row = (p.N / 256) * 2;
    p.K = K0 + K1;
    p.kblocks0 = K0 / TC_BK;
    CUtensorMap ma0, ma1, mb;
    int st = make_tmap_2d(&ma0, a0, 1, (uint64_t)p.M, (uint64_t)K0, (uint64_t)K0 * 2, 128, TC_BK, 1);
    if (st != MAC_OK) return st;
    if (K1 > 0) {
      st = make_tmap_2d(&ma1, a1, 1, (uint64_t)p.M, (uint64_t)K1, (uint64_t)K1 * 2, 128, TC_BK, 1);
      if (st != MAC_OK) return st;
    } else {
      ma1 = ma0;
    }
    st = make_tmap_2d(&mb, wt, 1, (uint64_t)p.N, (uint64_t)p.K, (uint64_t)ldw * 2, 128, TC_BK, 1);
    if (st != MAC_OK) return st;
    return tc2_gemm_dispatch(ma0, ma1, mb, p, stream);
  }
  const int BM = tile / 1000, BN = tile % 1000;
  if (nparts_per_row) *nparts_per_row = (p.N / BN) * (BM == 256 ? 2 : 4);
  p.K = K0 + K1;
  p.kblocks0 = K0 / TC_BK;
  CUtensorMap ma0, ma1, mb;
  int st = make_tmap_2d(&ma0, a0, 1, (uint64_t)p.M, (uint64_t)K0, (uint64_t)lda0 * 2, (uint32_t)BM, TC_BK, 1);
  if (st != MAC_OK) return st;
  if (K1 > 0) {
    st = make_tmap_2d(&ma1, a1, 1, (uint64_t)p.M, (uint64_t)K1, (uint64_t)lda1 * 2, (uint32_t)BM, TC_BK, 1);
    if (st != MAC_OK) return st;
  } else {
    ma1 = ma0;
  }
  st = make_tmap_2d(&mb, wt, 1, (uint64_t)p.N, (uint64_t)p.K, (uint64_t)ldw * 2, (uint32_t)BN, TC_BK, 1);
  if (st != MAC_OK) return st;
  if (tile == 256256) return tc_gemm_dispatch<256, 256>(ma0, ma1, mb, p, stream);
  if (tile == 128256) return tc_gemm_dispatch<128, 256>(ma0, ma1, mb, p, stream);
  return tc_gemm_dispatch<128, 128>(ma0, ma1, mb, p, stream);
}

// fp32 [K, N] (in, out) weight -> bf16 [N, K] (out, in): the K-major B operand of the forward GEMMs
__global__ void pack_weight_bf16_kernel(const float* __restrict__ W, __nv_bfloat16* __restrict__ Wt, int K, int N) {
  __shared__ float tile[32][33];
  const int k0 = blockIdx.y * 32, n0 = blockIdx.x * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int k = k0 + i, n = n0 + threadIdx.x;
    tile[i][threadIdx.x] = (k < K && n < N) ? W[(size_t)k * N + n] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int n = n0 + i, k = k0 + threadIdx.x;
    if (n < N && k < K) Wt[(size_t)n * K + k] = __float2bfloat16_rn(tile[threadIdx.x][i]);
  }
}

// Activations for a tensor-core weight gradient: fp32 X [K rows, N cols] -> bf16 X^T [N, K] (the contraction index K = B*N rows
// becomes the K-major direction of both wgrad operands), optionally also the row-major bf16 copy (the A operand of the matching
// data gradient) from the same read.  64x64 tiles: 256-byte row reads, 128-byte row writes.
//   MODE 0 plain; 1 x * rowvec[row / rows_per_batch, col] (P * y, ops.py:694-703); 2 dropout(x) with the forward's Philox stream
//   (one draw per aligned column quad, element index row*N + col: mac_dropout_fwd's numbering).      N % 4 == 0, K % 2 == 0.
template <int MODE>
__global__ void __launch_bounds__(256) pack_t_bf16_kernel(const float* __restrict__ X, __nv_bfloat16* __restrict__ Xt,
                                                         __nv_bfloat16* __restrict__ Xrm, int K, int N,
                                                         const float* __restrict__ rowvec, int rows_per_batch, uint32_t thresh,
                                                         float scale, uint64_t seed, int site, int step) {
  __shared__ float tile[64][65];                              // [col][row]
  const int k0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  const int tq = threadIdx.x & 15, tr = threadIdx.x >> 4;     // 16 column quads x 16 rows per pass
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    const int kk = tr + 16 * p;
    const int k = k0 + kk, n = n0 + tq * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (k < K && n < N) {
      v = *reinterpret_cast<const float4*>(X + (size_t)k * N + n);
      if (MODE == 1) {
        const float4 y = *reinterpret_cast<const float4*>(rowvec + (size_t)(k / rows_per_batch) * N + n);
        v.x *= y.x; v.y *= y.y; v.z *= y.z; v.w *= y.w;
      }
      if (MODE == 2) {
        const uint64_t e = (uint64_t)k * (uint64_t)N + (uint64_t)n;
        const Philox4 r = philox4x32_10(seed, e >> 2, (uint32_t)site, (uint32_t)step);
        v.x = ((r.x >> 8) >= thresh) ? v.x * scale : 0.f;
        v.y = ((r.y >> 8) >= thresh) ? v.y * scale : 0.f;
        v.z = ((r.z >> 8) >= thresh) ? v.z * scale : 0.f;
        v.w = ((r.w >> 8) >= thresh) ? v.w * scale : 0.f;
      }
      if (Xrm) *reinterpret_cast<uint2*>(Xrm + (size_t)k * N + n) = make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
    }
    tile[tq * 4 + 0][kk] = v.x;
    tile[tq * 4 + 1][kk] = v.y;
    tile[tq * 4 + 2][kk] = v.z;
    tile[tq * 4 + 3][kk] = v.w;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int r = warp; r < 64; r += 8) {
    const int n = n0 + r, k = k0 + 2 * lane;
    if (n < N && k + 1 < K)
      *reinterpret_cast<uint32_t*>(Xt + (size_t)n * K + k) = pack_bf16(tile[r][2 * lane], tile[r][2 * lane + 1]);
  }
}

struct PackTArgs {
  const float* rowvec = nullptr;
  int rows_per_batch = 1;
  uint32_t thresh = 0;
  float scale = 1.f;
  uint64_t seed = 0;
  int site = 0, step = 0;
};

inline int pack_t_bf16_launch(int mode, const float* X, void* Xt, void* Xrm, int K, int N, const PackTArgs& a,
                              cudaStream_t stream) {
  if (!X || !Xt || K <= 0 || N <= 0 || (N & 3) || (K & 1)) return MAC_ERR_INVALID;
  dim3 grid((N + 63) / 64, (K + 63) / 64);
  __nv_bfloat16* t = reinterpret_cast<__nv_bfloat16*>(Xt);
  __nv_bfloat16* r = reinterpret_cast<__nv_bfloat16*>(Xrm);
  if (mode == 0)
    pack_t_bf16_kernel<0><<<grid, 256, 0, stream>>>(X, t, r, K, N, nullptr, 1, 0u, 1.f, 0, 0, 0);
  else if (mode == 1)
    pack_t_bf16_kernel<1><<<grid, 256, 0, stream>>>(X, t, r, K, N, a.rowvec, a.rows_per_batch, 0u, 1.f, 0, 0, 0);
  else
    pack_t_bf16_kernel<2><<<grid, 256, 0, stream>>>(X, t, r, K, N, nullptr, 1, a.thresh, a.scale, a.seed, a.site, a.step);
  MAC_LAUNCH_CHECK();
  return MAC_OK;
}

// bf16 -> fp32 widening of up to three equally long slabs in one launch (the activations the tensor-core training forward
// leaves in bf16, read in fp32 by the backward kernels)
__global__ void __launch_bounds__(256) widen3_bf16_kernel(const uint4* __restrict__ s0, const uint4* __restrict__ s1,
                                                         const uint4* __restrict__ s2, float4* __restrict__ d0,
                                                         float4* __restrict__ d1, float4* __restrict__ d2, long long n8) {
  const uint4* s = blockIdx.y == 0 ? s0 : blockIdx.y == 1 ? s1 : s2;
  float4* d = blockIdx.y == 0 ? d0 : blockIdx.y == 1 ? d1 : d2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    const uint4 v = s[i];
    d[2 * i] = make_float4(__uint_as_float(v.x << 16), __uint_as_float(v.x & 0xffff0000u), __uint_as_float(v.y << 16),
                           __uint_as_float(v.y & 0xffff0000u));
    d[2 * i + 1] = make_float4(__uint_as_float(v.z << 16), __uint_as_float(v.z & 0xffff0000u), __uint_as_float(v.w << 16),
                               __uint_as_float(v.w & 0xffff0000u));
  }
}

// training mode: bf16 copy of dropout(KB) for this step (ops.py:678), same Philox stream as the fp32 path
__global__ void dropout_cast_bf16_kernel(const float4* __restrict__ x, uint2* __restrict__ out, uint32_t thresh,
                                         float scale, uint64_t seed, int site, int step, long long n4) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const float4 v = x[i];
  const Philox4 r = philox4x32_10(seed, (uint64_t)i, (uint32_t)site, (uint32_t)step);
  const float a = ((r.x >> 8) >= thresh) ? v.x * scale : 0.f, b = ((r.y >> 8) >= thresh) ? v.y * scale : 0.f;
  const float c = ((r.z >> 8) >= thresh) ? v.z * scale : 0.f, d = ((r.w >> 8) >= thresh) ? v.w * scale : 0.f;
  out[i] = make_uint2(pack_bf16(a, b), pack_bf16(c, d));
}

// extra workspace of the bf16 read chain: P, P*y, H (+ I1 when saving for backward), each [B*N, d] bf16
inline size_t tc_read_extra_workspace_bytes(int B, int N, int d) {
  return (size_t)5 * (((size_t)B * N * d * 2 + 1023) & ~(size_t)1023) + 1024;     // P, P*y, H, I1, dropout(KB)
}

// The read unit's three projections on tensor cores (see mac_read_fwd in mac_b200.h).
inline int tc_read_chain(const float* kb_f32, const void* kb_bf16, const float* y, const float* control, const mac_read_weights* w,
                         uint32_t thr, float scale, uint64_t seed, int step, float* /*P_f32*/, float* /*H_f32*/,
                         float* /*I1_f32*/, float* parts, int* nparts, void* ws, size_t ws_bytes, int B, int N, int d,
                         bool save, cudaStream_t stream) {
  if (!kb_bf16 || !w->Wx_bf16 || !w->Wm_bf16 || !w->Wm2_bf16) return MAC_ERR_INVALID;
  if (d % 128) return MAC_ERR_UNSUPPORTED;
  if (ws_bytes < tc_read_extra_workspace_bytes(B, N, d)) return MAC_ERR_WORKSPACE;
  const int M = B * N;
  const size_t slab = (((size_t)M * d * 2 + 1023) & ~(size_t)1023);
  char* base = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(ws) + 1023) & ~(uintptr_t)1023);
  __nv_bfloat16* P = reinterpret_cast<__nv_bfloat16*>(base);
  __nv_bfloat16* PY = reinterpret_cast<__nv_bfloat16*>(base + slab);
  __nv_bfloat16* H = reinterpret_cast<__nv_bfloat16*>(base + 2 * slab);
  __nv_bfloat16* I1 = save ? reinterpret_cast<__nv_bfloat16*>(base + 3 * slab) : nullptr;
  if (thr != 0) {
    // A operand of the first projection = dropout(KB) (ops.py:678): TMA cannot mask in flight, so make this step's copy
    __nv_bfloat16* kbd = reinterpret_cast<__nv_bfloat16*>(base + 4 * slab);
    const long long n4 = (long long)M * d / 4;
    dropout_cast_bf16_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, stream>>>(
        reinterpret_cast<const float4*>(kb_f32), reinterpret_cast<uint2*>(kbd), thr, scale, seed, MAC_SITE_READ_KB, step, n4);
    MAC_LAUNCH_CHECK();
    kb_bf16 = kbd;
  }
  TcGemmParams p{};
  p.M = M; p.N = d; p.rows_per_batch = N; p.ldo = d; p.seed = seed; p.step = step;
  // P = KB @ Wx + bx ; also P*y
  p.epi = TC_EPI_P; p.bias = w->bx; p.out0 = P; p.out1 = PY; p.y = y;
  int st = tc_gemm_launch(kb_bf16, d, nullptr, 0, w->Wx_bf16, p, stream);
  if (st != MAC_OK) return st;
  // H = ELU([P*y, P] @ Wm + bm)
  p.epi = TC_EPI_ACT; p.act = MAC_ACT_ELU; p.bias = w->bm; p.out0 = H; p.out1 = nullptr;
  st = tc_gemm_launch(PY, d, P, d, w->Wm_bf16, p, stream);
  if (st != MAC_OK) return st;
  // logits parts = sum_n ELU((H @ Wm2 + bm2) * control) * wr
  p.epi = TC_EPI_LOGITS; p.bias = w->bm2; p.out0 = I1; p.ctrl = control; p.wr = w->wr; p.parts = parts;
  p.e_thresh = thr; p.e_scale = scale; p.e_site = MAC_SITE_READ_INTER;
  st = tc_gemm_launch(H, d, nullptr, 0, w->Wm2_bf16, p, stream, nparts);
  if (st != MAC_OK) return st;
  return MAC_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// Eval-mode read (readDropout == 1): dropout(KB) == KB at every step and the read weights are shared over the steps
// (mac_cell.py:209-277 is called with the same variables for each of the netLength cells), so
//   P = KB @ Wx + bx                (ops.py:688)                       and
//   Q = P @ Wm[d:2d, :] + bm        (the projected-KB half of the [P*y, P] concat, mac_cell.py:236-238)
// do not depend on the step.  They are computed once per forward into `inv` = [P | Q] (bf16), and each step runs
//   PY = P * y_b ;  H = ELU(PY @ Wm[0:d, :] + Q) ;  logits = ... (unchanged)
// i.e. 2 x d instead of 4 x d MACs per knowledge-base element and step.
inline size_t tc_read_invariant_bytes(int B, int N, int d) {
  // [P | Q] bf16 slabs + the per-tile softmax partials of the packed fused read step (read_step.cuh: B x 3 x (512 + 2) floats)
  return (size_t)2 * (((size_t)B * N * d * 2 + 1023) & ~(size_t)1023) + 1024 + (size_t)B * 3 * (512 + 2) * sizeof(float) + 1024;
}

// PY[m, :] = P[m, :] * y[m / rows_per_batch, :]   (8 bf16 per thread)
__global__ void scale_rows_bf16_kernel(const uint4* __restrict__ P, const float* __restrict__ y, uint4* __restrict__ out,
                                       int rows_per_batch, int d, long long n8) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n8) return;
  const long long e = i * 8;
  const int c = (int)(e % d);
  const long long b = (e / d) / rows_per_batch;
  const uint4 v = P[i];
  const float4 y0 = __ldg(reinterpret_cast<const float4*>(y + b * d + c));
  const float4 y1 = __ldg(reinterpret_cast<const float4*>(y + b * d + c + 4));
  auto lo = [](uint32_t w) { return __uint_as_float(w << 16); };
  auto hi = [](uint32_t w) { return __uint_as_float(w & 0xffff0000u); };
  out[i] = make_uint4(pack_bf16(lo(v.x) * y0.x, hi(v.x) * y0.y), pack_bf16(lo(v.y) * y0.z, hi(v.y) * y0.w),
                      pack_bf16(lo(v.z) * y1.x, hi(v.z) * y1.y), pack_bf16(lo(v.w) * y1.z, hi(v.w) * y1.w));
}

inline char* tc_align1k(void* p) {
  return reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(p) + 1023) & ~(uintptr_t)1023);
}

inline int tc_read_invariant(const void* kb_bf16, const mac_read_weights* w, void* inv, size_t inv_bytes, int B, int N,
                             int d, cudaStream_t stream) {
  if (!kb_bf16 || !w->Wx_bf16 || !w->Wm_bf16 || !inv) return MAC_ERR_INVALID;
  if (d % 128) return MAC_ERR_UNSUPPORTED;
  if (inv_bytes < tc_read_invariant_bytes(B, N, d)) return MAC_ERR_WORKSPACE;
  const int M = B * N;
  const size_t slab = (((size_t)M * d * 2 + 1023) & ~(size_t)1023);
  char* base = tc_align1k(inv);
  __nv_bfloat16* P = reinterpret_cast<__nv_bfloat16*>(base);
  __nv_bfloat16* Q = reinterpret_cast<__nv_bfloat16*>(base + slab);
  TcGemmParams p{};
  p.M = M; p.N = d; p.rows_per_batch = N; p.ldo = d;
  p.epi = TC_EPI_ACT; p.act = MAC_ACT_NON; p.bias = w->bx; p.out0 = P;
  int st = tc_gemm_launch(kb_bf16, d, nullptr, 0, w->Wx_bf16, p, stream);
  if (st != MAC_OK) return st;
  p.bias = w->bm; p.out0 = Q;
  return tc_gemm_launch(P, d, nullptr, 0, reinterpret_cast<const __nv_bfloat16*>(w->Wm_bf16) + d, p, stream, nullptr, 2 * d);
}

inline int tc_read_chain_inv(const void* inv, const float* y, const float* control, const mac_read_weights* w, float* parts,
                             int* nparts, void* ws, size_t ws_bytes, int B, int N, int d, cudaStream_t stream) {
  if (!inv || !w->Wm_bf16 || !w->Wm2_bf16) return MAC_ERR_INVALID;
  if (d % 128) return MAC_ERR_UNSUPPORTED;
  if (ws_bytes < tc_read_extra_workspace_bytes(B, N, d)) return MAC_ERR_WORKSPACE;
  const int M = B * N;
  const size_t slab = (((size_t)M * d * 2 + 1023) & ~(size_t)1023);
  const char* ibase = tc_align1k(const_cast<void*>(inv));
  const __nv_bfloat16* P = reinterpret_cast<const __nv_bfloat16*>(ibase);
  const __nv_bfloat16* Q = reinterpret_cast<const __nv_bfloat16*>(ibase + slab);
  char* base = tc_align1k(ws);
  __nv_bfloat16* PY = reinterpret_cast<__nv_bfloat16*>(base + slab);
  __nv_bfloat16* H = reinterpret_cast<__nv_bfloat16*>(base + 2 * slab);
  const long long n8 = (long long)M * d / 8;
  scale_rows_bf16_kernel<<<(unsigned)((n8 + 255) / 256), 256, 0, stream>>>(
      reinterpret_cast<const uint4*>(P), y, reinterpret_cast<uint4*>(PY), N, d, n8);
  MAC_LAUNCH_CHECK();
  TcGemmParams p{};
  p.M = M; p.N = d; p.rows_per_batch = N; p.ldo = d;
  int st;
  static const bool q_hoist = !(getenv("MAC_READ_QHOIST") && atoi(getenv("MAC_READ_QHOIST")) == 0);
  if (q_hoist) {
    // H = ELU(PY @ Wm[0:d] + Q)      (bm is inside Q)
    p.epi = TC_EPI_ADDACT; p.act = MAC_ACT_ELU; p.bias = nullptr; p.out0 = H; p.add = Q;
    st = tc_gemm_launch(PY, d, nullptr, 0, w->Wm_bf16, p, stream, nullptr, 2 * d);
  } else {
    // experiment switch: keep the concatenated K = 2d form, H = ELU([PY, P] @ Wm + bm), with only P hoisted
    p.epi = TC_EPI_ACT; p.act = MAC_ACT_ELU; p.bias = w->bm; p.out0 = H;
    st = tc_gemm_launch(PY, d, P, d, w->Wm_bf16, p, stream);
  }
  if (st != MAC_OK) return st;
  p.epi = TC_EPI_LOGITS; p.act = MAC_ACT_NON; p.bias = w->bm2; p.out0 = nullptr; p.add = nullptr;
  p.ctrl = control; p.wr = w->wr; p.parts = parts;
  return tc_gemm_launch(H, d, nullptr, 0, w->Wm2_bf16, p, stream, nparts);
}


// ---------------------------------------------------------------------------------------------------------------
// Weight gradients on tensor cores: dW[in, out] (+)= X^T[in, M] @ G^T[out, M]^T with K = M = B*N (12 544 at the headline shape)
// and only (in/128) x (out/256) = 8-16 output tiles: one K loop of 196 k-blocks per CTA left 130+ SMs idle (44.7 us per
// launch, 12 % of a tensor-core training step).  Split-K: the K range is cut into S slices, every (slice, tile) is a tile of the
// persistent kernel and writes its fp32 partial; the partials are summed in slice order (deterministic) into dW.
__global__ void splitk_accum_kernel(const float4* __restrict__ part, float4* __restrict__ dst, long long n4, int S,
                                    long long stride4, int accumulate) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  float4 a = accumulate ? dst[i] : make_float4(0.f, 0.f, 0.f, 0.f);
  for (int s_ = 0; s_ < S; ++s_) {
    const float4 v = part[i + s_ * stride4];
    a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
  }
  dst[i] = a;
}

// largest S <= 28 that divides K / 64, keeps >= 2 k-blocks per slice and at most two waves of tiles
inline int tc_pick_ksplit(int K, int out_tiles) {
  const int kblocks = K / TC_BK;
  int best = 1;
  for (int S = 2; S <= 28; ++S)
    if (kblocks % S == 0 && kblocks / S >= 2 && out_tiles * S <= 2 * tc_num_sms()) best = S;
  return best;
}
inline size_t tc_wgrad_partial_bytes(int in_dim, int out_dim) { return (size_t)28 * in_dim * out_dim * 4; }

// dW[in, out] += xT[in, K] @ gT[out, K]^T   (both bf16, K-major); `partial` holds tc_wgrad_partial_bytes(in, out)
inline int tc_wgrad_splitk(const void* xT, const void* gT, float* dW, float* partial, int in_dim, int out_dim, int K,
                           cudaStream_t stream) {
  if ((in_dim % 128) || (out_dim % 128) || (K % TC_BK)) return MAC_ERR_UNSUPPORTED;
  TcGemmParams p{};
  p.M = in_dim; p.N = out_dim; p.act = MAC_ACT_NON; p.bias = nullptr; p.ldo = out_dim; p.rows_per_batch = 1;
  p.epi = TC_EPI_F32; p.outf = partial;
  const int out_tiles = (in_dim / 128) * (out_dim / (out_dim % 256 == 0 ? 256 : 128));
  const int S = tc_pick_ksplit(K, out_tiles);
  p.ksplit = S; p.split_stride = (long long)in_dim * out_dim;
  int st = tc_gemm_launch(xT, K, nullptr, 0, gT, p, stream);
  if (st != MAC_OK) return st;
  const long long n4 = (long long)in_dim * out_dim / 4;
  splitk_accum_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, stream>>>(reinterpret_cast<const float4*>(partial),
                                                                       reinterpret_cast<float4*>(dW), n4, S, n4, 1);
  MAC_LAUNCH_CHECK();
  return MAC_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// "tc32": the read unit's three [B*N, .] projections as SPLIT-bf16 products on the tensor cores, for the <= 1e-4 parity bar
// of BASELINE.json (the plain bf16 path is at ~1e-3).  Every fp32 operand x is carried as two bf16 values x_hi = bf16(x),
// x_lo = bf16(x - x_hi) and the product uses three of the four partial products, accumulated in ONE fp32 accumulator:
//     A W  ~=  A_hi W_hi + A_lo W_hi + A_hi W_lo                      (A_lo W_lo ~ 2^-16 relative is dropped)
// With A' = [A_hi | A_lo] ([M, 2K]) and W' = [W_hi | W_hi | W_lo] ([N, 3K], K-major) this is the existing two-segment GEMM:
// segment 0 = A'[:, 0:2K] against W'[:, 0:2K], segment 1 = A'[:, 0:K] against W'[:, 2K:3K] -- K triples, nothing else
// changes.  Producers write their outputs directly as hi | lo pairs (TC_EPI_ACT_SPLIT) or as fp32 (TC_EPI_F32).
// Inference form only (P, Q hoisted, mac_read_invariant): inv = [P fp32 | Q fp32 | scratch [M, 2d] bf16].
// ---------------------------------------------------------------------------------------------------------------
// out[m, c] = hi(x[m, c] * y[m / rows_per_batch, c]), out[m, d + c] = lo(...)     (y == NULL: plain split); 4 columns / thread
__global__ void split_rows_kernel(const float4* __restrict__ x, const float* __restrict__ y, uint2* __restrict__ out,
                                  int rows_per_batch, int d, long long n4) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const int d4 = d / 4;
  const long long m = i / d4;
  const int c4 = (int)(i - m * d4);
  float4 v = x[i];
  if (y) {
    const float4 s = __ldg(reinterpret_cast<const float4*>(y + (m / rows_per_batch) * d) + c4);
    v.x *= s.x; v.y *= s.y; v.z *= s.z; v.w *= s.w;
  }
  const uint32_t h01 = pack_bf16(v.x, v.y), h23 = pack_bf16(v.z, v.w);
  const uint32_t l01 = pack_bf16(v.x - __uint_as_float(h01 << 16), v.y - __uint_as_float(h01 & 0xffff0000u));
  const uint32_t l23 = pack_bf16(v.z - __uint_as_float(h23 << 16), v.w - __uint_as_float(h23 & 0xffff0000u));
  uint2* row = out + m * (2 * d4);
  row[c4] = make_uint2(h01, h23);
  row[d4 + c4] = make_uint2(l01, l23);
}
inline int split_rows_launch(const float* x, const float* y, void* out, int rows_per_batch, int d, long long M,
                             cudaStream_t stream) {
  const long long n4 = M * d / 4;
  split_rows_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, stream>>>(reinterpret_cast<const float4*>(x), y,
                                                                      reinterpret_cast<uint2*>(out), rows_per_batch, d, n4);
  MAC_LAUNCH_CHECK();
  return MAC_OK;
}

// fp32 W[K, N] (rows k0 .. k0+K of a wider [*, N] weight are passed as W + k0*N) -> bf16 Wt3[N, 3K] = [hi | hi | lo]
__global__ void pack_weight_split3_kernel(const float* __restrict__ W, __nv_bfloat16* __restrict__ Wt3, int K, int N) {
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
      __nv_bfloat16* row = Wt3 + (size_t)n * 3 * K;
      row[k] = h;
      row[K + k] = h;
      row[2 * K + k] = __float2bfloat16_rn(w - __bfloat162float(h));
    }
  }
}

inline size_t tc3_slab(int B, int N, int d) { return (((size_t)B * N * 2 * d * 2 + 1023) & ~(size_t)1023); }   // [M, 2d] bf16
inline size_t tc3_invariant_bytes(int B, int N, int d) {
  return (size_t)2 * B * N * d * 4 + tc3_slab(B, N, d) + 2048;
}
inline size_t tc3_extra_workspace_bytes(int B, int N, int d) { return 2 * tc3_slab(B, N, d) + 1024; }

// C = epilogue(A W) with A' = [A_hi | A_lo] ([M, 2K] bf16) and W' = [W_hi | W_hi | W_lo] ([N, 3K] bf16)
inline int tc3_gemm(const void* a_split, int K, const void* wt3, TcGemmParams p, cudaStream_t stream, int* nparts = nullptr) {
  return tc_gemm_launch(a_split, 2 * K, a_split, K, wt3, p, stream, nparts, 3 * K, 2 * K, 2 * K);
}

inline int tc3_read_invariant(const float* kb, const mac_read_weights* w, void* inv, size_t inv_bytes, int B, int N, int d,
                              cudaStream_t stream) {
  if (!kb || !w->Wx_s3 || !w->Wmb_s3 || !inv) return MAC_ERR_INVALID;
  if (d % 128) return MAC_ERR_UNSUPPORTED;
  if (inv_bytes < tc3_invariant_bytes(B, N, d)) return MAC_ERR_WORKSPACE;
  const int M = B * N;
  float* P = reinterpret_cast<float*>(inv);
  float* Q = P + (size_t)M * d;
  void* scratch = tc_align1k(Q + (size_t)M * d);
  TcGemmParams p{};
  p.M = M; p.N = d; p.rows_per_batch = N; p.ldo = d; p.epi = TC_EPI_F32; p.act = MAC_ACT_NON;
  int st = split_rows_launch(kb, nullptr, scratch, N, d, M, stream);
  if (st != MAC_OK) return st;
  p.bias = w->bx; p.outf = P;
  st = tc3_gemm(scratch, d, w->Wx_s3, p, stream);                       // P = KB @ Wx + bx            (ops.py:688)
  if (st != MAC_OK) return st;
  st = split_rows_launch(P, nullptr, scratch, N, d, M, stream);
  if (st != MAC_OK) return st;
  p.bias = w->bm; p.outf = Q;
  return tc3_gemm(scratch, d, w->Wmb_s3, p, stream);                    // Q = P @ Wm[d:2d] + bm       (mac_cell.py:236-238)
}

inline int tc3_read_chain_inv(const void* inv, const float* y, const float* control, const mac_read_weights* w, float* parts,
                              int* nparts, void* ws, size_t ws_bytes, int B, int N, int d, cudaStream_t stream) {
  if (!inv || !w->Wma_s3 || !w->Wm2_s3) return MAC_ERR_INVALID;
  if (d % 128) return MAC_ERR_UNSUPPORTED;
  if (ws_bytes < tc3_extra_workspace_bytes(B, N, d)) return MAC_ERR_WORKSPACE;
  const int M = B * N;
  const float* P = reinterpret_cast<const float*>(inv);
  const float* Q = P + (size_t)M * d;
  char* base = tc_align1k(ws);
  void* PYs = base;
  __nv_bfloat16* Hs = reinterpret_cast<__nv_bfloat16*>(base + tc3_slab(B, N, d));
  int st = split_rows_launch(P, y, PYs, N, d, M, stream);                // (P * y_b) as hi | lo       (ops.py:694-703)
  if (st != MAC_OK) return st;
  TcGemmParams p{};
  p.M = M; p.N = d; p.rows_per_batch = N;
  p.epi = TC_EPI_ACT_SPLIT; p.act = MAC_ACT_ELU; p.bias = nullptr; p.addf = Q; p.ldaf = d; p.out0 = Hs; p.ldo = 2 * d;
  st = tc3_gemm(PYs, d, w->Wma_s3, p, stream);                           // H = ELU((P*y) @ Wm[0:d] + Q) as hi | lo
  if (st != MAC_OK) return st;
  p.epi = TC_EPI_LOGITS; p.act = MAC_ACT_NON; p.bias = w->bm2; p.addf = nullptr; p.out0 = nullptr; p.ldo = d;
  p.ctrl = control; p.wr = w->wr; p.parts = parts;
  return tc3_gemm(Hs, d, w->Wm2_s3, p, stream, nparts);                  // logits partial sums         (mac_cell.py:248-266)
}

}  // namespace mac
