// K7 on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), sm_100a only.
//
//   f = theta_hat^T phi_hat  ->  sim = rowmax f  ->  P = softmax_j(f / T)  ->  y = P V      (NonlocalNet.py:477-498)
//
// fp32-class accuracy from low-precision MMAs by operand splitting: x = hi + lo with hi, lo exactly
// representable in the MMA input type (tf32: 2 x 11 significant bits; bf16: 2 x 8), and
//   f ~= hi_a.hi_b + hi_a.lo_b + lo_a.hi_b           (lo.lo dropped: 2^-22 resp. 2^-16 relative)
// all three accumulated into the same fp32 TMEM tile.  A third format, FP16X3, splits x * 2^14 into two fp16 planes
// (theta_hat / phi_hat are unit vectors, so the fixed power-of-two scale is exact and cannot overflow): the same
// 2 x 11 significant bits as tf32 at twice the MMA rate and half the operand bytes; scores come out times 2^28.
// FP16X3 is the default (|df| ~ 5e-7, argmax identical to fp64 on every test); DVC_MATH_TF32X3 (|df| ~ 1e-7) and
// DVC_MATH_BF16X3 (|df| ~ 2e-6) are selectable.  By default two CTAs on adjacent query tiles run as a pair
// (tcgen05.mma.cta_group::2, each staging half of the reference tile; Cfg<2> below); the single-CTA form:
//
// Kernel structure (one CTA = 128 query rows x a range of 256-column tiles of reference positions):
//   warp 0      TMA producer: per k-block (128 bytes of K) loads A_hi, A_lo [128 x 128B] and B_hi, B_lo [256 x 128B]
//               with SWIZZLE_128B into a 2-stage shared-memory ring (96 KB per stage), mbarrier expect_tx.
//   warp 1      TMEM owner + MMA issuer: one elected thread issues 4 k-steps x 3 tcgen05.mma (M128 x N256) per
//               stage into one of two 256-column fp32 accumulators (all 512 TMEM columns), tcgen05.commit
//               releases the smem stage and, after the last k-block, publishes the score tile.
//   warps 2..5  epilogue: thread t owns query row t (TMEM lane t): tcgen05.ld 32 columns at a time, running
//               (max, argmax) or online softmax (max, sum, 3 colour sums) entirely in registers -- no
//               cross-thread reduction; overlaps the MMAs of the next tile through the double-buffered TMEM.
// The N x N score matrix never leaves the SM.  Column-range splits (grid.z) balance the 148 SMs; a small merge
// kernel combines the per-split row statistics.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cudaTypedefs.h>
#include <math.h>

#include <mutex>

#include "corr_tc.cuh"
#include "tc_common.cuh"

namespace dvc {

// ------------------------------------------------------------------------------------------------
// tensor-map encoding through the driver entry point (no link-time libcuda dependency)
// ------------------------------------------------------------------------------------------------
int encode_tmap_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows, uint32_t box_cols,
                   int elem_bytes, int swizzle_bytes) {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  });
  if (!fn) return -1;
  const cuuint64_t gdim[2] = {cols, rows};
  const cuuint64_t gstride[1] = {cols * (uint64_t)elem_bytes};
  const cuuint32_t box[2] = {box_cols, box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUtensorMapDataType dt = elem_bytes == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  CUresult r = fn(out, dt, 2, const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : (int)r;
}

namespace {

constexpr int BM = 128;        // query rows per CTA (= TMEM lanes)
constexpr int BN = 256;        // reference positions per score tile (= TMEM columns per accumulator)
// CL = 2: CTA pairs (tcgen05.mma.cta_group::2) on two adjacent 128-row query tiles.  The three MMAs per k-step
// re-read both operands from shared memory, which binds a single CTA to the 128 B/clk shared-memory port (96 KB of TMA
// writes + 144 KB of MMA reads per 1536 tensor cycles = 156 B/clk; ncu: tensor pipe 74-82 %).  In a pair each CTA
// stages its own query rows and HALF of the reference-position tile (104 B/clk), and three 64 KB stages fit.
template <int CL>
struct Cfg {
  static constexpr int STAGES = CL == 2 ? 3 : 2;
  static constexpr int A_BYTES = BM * 128, B_BYTES = (BN / CL) * 128;
  static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;  // 98304 / 65536
  static constexpr int V_RING_BYTES = 2 * BN * 16;  // softmax epilogue: the V rows of the tile in flight, two slots
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*alignment slack*/ + 256 /*barriers*/ + V_RING_BYTES;
};
// warp 0 TMA, warp 1 MMA, then 8 epilogue warps: two per TMEM lane quarter, each owning 128 of the tile's 256 columns --
// twice the threads to hide the tcgen05.ld / exp2 / FMA latency chains behind (with 4 warps the exact argmax kernel ran at
// 39 % tensor-pipe activity under ncu, the 8-warp softmax kernel at 71 %)
template <bool SOFTMAX>
struct Epi {
  static constexpr int WARPS = 8;
  static constexpr int HALVES = WARPS / 4;
  static constexpr int NTHREADS = 64 + 32 * WARPS;
};

// packed fp32 pair arithmetic (Blackwell FFMA2): acc.{x,y} += a.{x,y} * b.{x,y}
__device__ __forceinline__ void fma2(unsigned long long& acc, unsigned long long a, unsigned long long b) {
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(b));
}
__device__ __forceinline__ unsigned long long pack2(float x, float y) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(x), "f"(y));
  return r;
}
__device__ __forceinline__ float2 unpack2(unsigned long long v) {
  float2 r;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
  return r;
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// 1-D bulk copy global -> this CTA's shared memory, completing on a local mbarrier
__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(tc::smem_u32(smem_dst)),
               "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(tc::smem_u32(bar))
               : "memory");
}

// per (split part, row) partial statistics.  Softmax: running max m, sum s of exp weights, weighted colour sums a*.
// Argmax: max m, lowest column idx attaining it, s = NUMBER of columns whose score equals m bit for bit and a* = the sum
// of their V rows -- the reference's fp32 softmax(f / 1e-10) averages the V rows of bit-equal maxima (duplicated
// exemplar columns: letterbox bars, flat regions), NonlocalNet.py:486-497.
struct SplitOut {
  float m, s, a0, a1, a2;
  int idx;
  float pad0, pad1;
};

struct TcParams {
  int NA, NB, B, Bphi, C;
  int tiles_per_split;
  float sc;  // log2(e) / T
  const float* row_sc;  // optional per-query-row log2(e) / T_i (overrides sc): the contextual loss normalises every row by its own minimum distance
  float out_scale;  // scores in TMEM are true scores / out_scale (2^-28 for pre-scaled fp16 operands, else 1)
  const float4* V;
  SplitOut* part;  // [nsplit][B*NA]
};

// ---- operand split: rows [R][C] fp32 -> hi / lo planes -----------------------------------------------
__device__ __forceinline__ float tf32_rna(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}

template <int FMT>
__global__ void __launch_bounds__(256) split_planes_kernel(const float* __restrict__ src, void* __restrict__ hi,
                                                           void* __restrict__ lo, size_t n4) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(src) + i);
    const float x[4] = {v.x, v.y, v.z, v.w};
    if constexpr (FMT == 2) {
      __half h[4], l[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float xs = x[j] * 16384.0f;
        h[j] = __float2half_rn(xs);
        l[j] = __float2half_rn(xs - __half2float(h[j]));
      }
      reinterpret_cast<uint2*>(hi)[i] = make_uint2(
          (uint32_t)__half_as_ushort(h[0]) | ((uint32_t)__half_as_ushort(h[1]) << 16),
          (uint32_t)__half_as_ushort(h[2]) | ((uint32_t)__half_as_ushort(h[3]) << 16));
      reinterpret_cast<uint2*>(lo)[i] = make_uint2(
          (uint32_t)__half_as_ushort(l[0]) | ((uint32_t)__half_as_ushort(l[1]) << 16),
          (uint32_t)__half_as_ushort(l[2]) | ((uint32_t)__half_as_ushort(l[3]) << 16));
    } else if constexpr (FMT == 0) {
      float h[4], l[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) h[j] = tf32_rna(x[j]), l[j] = tf32_rna(x[j] - h[j]);
      reinterpret_cast<float4*>(hi)[i] = make_float4(h[0], h[1], h[2], h[3]);
      reinterpret_cast<float4*>(lo)[i] = make_float4(l[0], l[1], l[2], l[3]);
    } else {
      __nv_bfloat16 h[4], l[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        h[j] = __float2bfloat16_rn(x[j]);
        l[j] = __float2bfloat16_rn(x[j] - __bfloat162float(h[j]));
      }
      reinterpret_cast<uint2*>(hi)[i] = make_uint2(
          (uint32_t)__bfloat16_as_ushort(h[0]) | ((uint32_t)__bfloat16_as_ushort(h[1]) << 16),
          (uint32_t)__bfloat16_as_ushort(h[2]) | ((uint32_t)__bfloat16_as_ushort(h[3]) << 16));
      reinterpret_cast<uint2*>(lo)[i] = make_uint2(
          (uint32_t)__bfloat16_as_ushort(l[0]) | ((uint32_t)__bfloat16_as_ushort(l[1]) << 16),
          (uint32_t)__bfloat16_as_ushort(l[2]) | ((uint32_t)__bfloat16_as_ushort(l[3]) << 16));
    }
  }
}

// ---- screened T -> 0 path: operand preparation ---------------------------------------------------------
// One warp per row of [R][256]: the fp16 hi plane of x * 2^14 (the only operand of the screening pass) and the exact
// Euclidean norm of what the plane drops, d = x - hi * 2^-14, which bounds the screening error of every score of that
// row: |f - f_screen| <= |d_a . b| + |a_hi . d_b| <= ||d_a|| ||b|| + ||a_hi|| ||d_b||  (Cauchy-Schwarz).
// nd[row] = ||d_row|| (rounded up), nh[row] = ||hi_row * 2^-14|| (rounded up); *nd_max = max over rows (float bits).
__global__ void __launch_bounds__(256) screen_planes_kernel(const float* __restrict__ src, __half* __restrict__ hi, float* __restrict__ nd,
                                                            float* __restrict__ nh, unsigned int* __restrict__ nd_max,
                                                            unsigned int* __restrict__ nh_max, int R) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= R) return;
  const float4* sp = reinterpret_cast<const float4*>(src + (size_t)warp * 256);
  float sd = 0.f, sh = 0.f;
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const float4 v = __ldg(sp + k * 32 + lane);
    const float x[4] = {v.x, v.y, v.z, v.w};
    unsigned short hb[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const __half h = __float2half_rn(x[j] * 16384.0f);
      hb[j] = __half_as_ushort(h);
      const float hf = __half2float(h) * 6.103515625e-05f;  // exact: a power-of-two scale
      const float d = x[j] - hf;                            // exact (Sterbenz / few significant bits)
      sd = fmaf(d, d, sd), sh = fmaf(hf, hf, sh);
    }
    reinterpret_cast<uint2*>(hi + (size_t)warp * 256)[k * 32 + lane] =
        make_uint2((uint32_t)hb[0] | ((uint32_t)hb[1] << 16), (uint32_t)hb[2] | ((uint32_t)hb[3] << 16));
  }
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) sd += __shfl_xor_sync(0xffffffffu, sd, off), sh += __shfl_xor_sync(0xffffffffu, sh, off);
  if (lane == 0) {
    // round the norms up generously (fp32 summation error of 256 non-negative terms is < 2^-15 relative)
    const float a = sqrtf(sd) * 1.0001f + 1e-12f, b = sqrtf(sh) * 1.0001f;
    nd[warp] = a, nh[warp] = b;
    atomicMax(nd_max, __float_as_uint(a));
    atomicMax(nh_max, __float_as_uint(b));
  }
}

// ---- main kernel ---------------------------------------------------------------------------------------
template <int FMT, bool SOFTMAX, int CL>
__global__ void __launch_bounds__(Epi<SOFTMAX>::NTHREADS, 1)
    corr_tc_kernel(const __grid_constant__ CUtensorMap tmAh, const __grid_constant__ CUtensorMap tmAl,
                   const __grid_constant__ CUtensorMap tmBh, const __grid_constant__ CUtensorMap tmBl, const TcParams p) {
  constexpr bool TF32 = (FMT == 0);
  constexpr int KB = TF32 ? 32 : 64;       // K elements per 128-byte k-block
  constexpr int UMMA_K_BYTES = 32;         // one MMA consumes 32 bytes of K (8 tf32 / 16 bf16)
  constexpr uint32_t IDESC = tc::umma_idesc(FMT == 0 ? 2u : (FMT == 1 ? 1u : 0u), BM * CL, BN);  // tf32 / bf16 / f16
  constexpr int STAGES = Cfg<CL>::STAGES, A_BYTES = Cfg<CL>::A_BYTES, B_BYTES = Cfg<CL>::B_BYTES;
  constexpr int STAGE_BYTES = Cfg<CL>::STAGE_BYTES;
  const int crank = (CL == 2) ? (int)tc::cluster_ctarank() : 0;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* full = bars;                 // [STAGES]   TMA -> MMA
  uint64_t* empty = bars + STAGES;       // [STAGES]   MMA -> TMA
  uint64_t* tfull = bars + 2 * STAGES;   // [2]        MMA -> epilogue
  uint64_t* tempty = bars + 2 * STAGES + 2;  // [2]    epilogue -> MMA
  uint64_t* vfull = bars + 2 * STAGES + 4;   // [2]    bulk copy of the tile's V rows -> epilogue (softmax)
  uint64_t* vempty = bars + 2 * STAGES + 6;  // [2]    this CTA's epilogue -> its producer (softmax)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 8);
  float4* v_ring = reinterpret_cast<float4*>(smem + STAGES * STAGE_BYTES + 256);  // [2][BN]
  constexpr int EPI_WARPS = Epi<SOFTMAX>::WARPS;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y;
  const int bphi = (p.Bphi == 1) ? 0 : b;
  const int m0 = blockIdx.x * BM;
  const int ntiles_all = (p.NB + BN - 1) / BN;
  const int t0 = blockIdx.z * p.tiles_per_split;
  const int t1 = min(t0 + p.tiles_per_split, ntiles_all);
  const int ntiles = max(t1 - t0, 0);
  const int nkb = p.C / KB;

  if (threadIdx.x == 0) {
    tc::tma_prefetch_desc(&tmAh);
    tc::tma_prefetch_desc(&tmAl);
    tc::tma_prefetch_desc(&tmBh);
    tc::tma_prefetch_desc(&tmBl);
    for (int i = 0; i < STAGES; ++i) tc::mbar_init(&full[i], 1), tc::mbar_init(&empty[i], 1);
    for (int i = 0; i < 2; ++i) {
      tc::mbar_init(&tfull[i], 1), tc::mbar_init(&tempty[i], EPI_WARPS * CL);  // pair: both epilogues
      tc::mbar_init(&vfull[i], 1), tc::mbar_init(&vempty[i], EPI_WARPS);
    }
    tc::fence_barrier_init();
  }
  if (CL == 2) tc::cluster_sync_all();  // both CTAs are resident before the pair allocation
  if (warp == 1) {
    if (CL == 2) {
      tc::tmem_alloc_pair(tmem_slot, 512);
      tc::tmem_relinquish_pair();
    } else {
      tc::tmem_alloc(tmem_slot, 512);
      tc::tmem_relinquish();
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (CL == 2) tc::cluster_sync_all();  // the peer's barriers are initialised before any remote arrive reaches them
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================= TMA producer (warp-convergent loop, one elected lane issues) =================
    {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = 0; t < ntiles; ++t) {
        const int col0 = bphi * p.NB + (t0 + t) * BN;
        for (int kb = 0; kb < nkb; ++kb) {
          tc::mbar_wait(&empty[stage], phase ^ 1);
          if (tc::elect_one()) {
            uint8_t* st = smem + stage * STAGE_BYTES;
            if (CL == 1) {
              tc::mbar_arrive_expect_tx(&full[stage], STAGE_BYTES);
              tc::tma_load_2d(st, &tmAh, &full[stage], kb * KB, b * p.NA + m0);
              tc::tma_load_2d(st + A_BYTES, &tmAl, &full[stage], kb * KB, b * p.NA + m0);
              tc::tma_load_2d(st + 2 * A_BYTES, &tmBh, &full[stage], kb * KB, col0);
              tc::tma_load_2d(st + 2 * A_BYTES + B_BYTES, &tmBl, &full[stage], kb * KB, col0);
            } else {  // my query rows and my half of the reference positions; the leader's barrier counts all bytes
              if (crank == 0) tc::mbar_arrive_expect_tx(&full[stage], 2 * STAGE_BYTES);
              const int hrow = crank * (BN / 2);
              tc::tma_load_2d_pair(st, &tmAh, &full[stage], kb * KB, b * p.NA + m0);
              tc::tma_load_2d_pair(st + A_BYTES, &tmAl, &full[stage], kb * KB, b * p.NA + m0);
              tc::tma_load_2d_pair(st + 2 * A_BYTES, &tmBh, &full[stage], kb * KB, col0 + hrow);
              tc::tma_load_2d_pair(st + 2 * A_BYTES + B_BYTES, &tmBl, &full[stage], kb * KB, col0 + hrow);
            }
          }
          __syncwarp();
          if (++stage == STAGES) stage = 0, phase ^= 1;
        }
        if (SOFTMAX) {
          // the V rows of this tile's columns for my own epilogue.  Issued after the tile's last k-block: by then the
          // MMAs of this tile have started, so the slot (freed together with the accumulator of tile t-2) is free and
          // the wait below never delays the operand prefetch.
          const int buf = t & 1;
          tc::mbar_wait(&vempty[buf], ((t >> 1) & 1) ^ 1);
          if (tc::elect_one()) {
            const int c0 = (t0 + t) * BN;
            const uint32_t bytes = (uint32_t)min(BN, p.NB - c0) * 16u;
            tc::mbar_arrive_expect_tx(&vfull[buf], bytes);
            bulk_load_1d(v_ring + buf * BN, p.V + (size_t)bphi * p.NB + c0, bytes, &vfull[buf]);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (warp-convergent loop, one elected lane issues; pair: leader CTA only) =====
    if (CL == 1 || crank == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = 0; t < ntiles; ++t) {
        const int buf = t & 1;
        const uint32_t acc_phase = (t >> 1) & 1;
        tc::mbar_wait(&tempty[buf], acc_phase ^ 1);
        tc::tc_fence_after();
        const uint32_t d = tmem_base + buf * BN;
        for (int kb = 0; kb < nkb; ++kb) {
          tc::mbar_wait(&full[stage], phase);
          tc::tc_fence_after();
          const uint32_t sa = tc::smem_u32(smem + stage * STAGE_BYTES);
          const uint64_t dAh = tc::umma_desc_k128(sa), dAl = tc::umma_desc_k128(sa + A_BYTES);
          const uint64_t dBh = tc::umma_desc_k128(sa + 2 * A_BYTES), dBl = tc::umma_desc_k128(sa + 2 * A_BYTES + B_BYTES);
          if (tc::elect_one()) {
#pragma unroll
            for (int kk = 0; kk < 128 / UMMA_K_BYTES; ++kk) {
              const uint64_t adv = (uint64_t)((kk * UMMA_K_BYTES) >> 4);  // start-address field is in 16-byte units
              // small cross terms first, the dominant hi.hi term last
              if (CL == 1) {
                tc::umma_ss<TF32>(d, dAl + adv, dBh + adv, IDESC, (kb | kk) ? 1u : 0u);
                tc::umma_ss<TF32>(d, dAh + adv, dBl + adv, IDESC, 1u);
                tc::umma_ss<TF32>(d, dAh + adv, dBh + adv, IDESC, 1u);
              } else {
                tc::umma_ss_pair<TF32>(d, dAl + adv, dBh + adv, IDESC, (kb | kk) ? 1u : 0u);
                tc::umma_ss_pair<TF32>(d, dAh + adv, dBl + adv, IDESC, 1u);
                tc::umma_ss_pair<TF32>(d, dAh + adv, dBh + adv, IDESC, 1u);
              }
            }
            if (CL == 1) {
              tc::umma_commit(&empty[stage]);  // smem stage reusable once these MMAs have read it
              if (kb == nkb - 1) tc::umma_commit(&tfull[buf]);
            } else {
              tc::umma_commit_pair_mc(&empty[stage], 3);
              if (kb == nkb - 1) tc::umma_commit_pair_mc(&tfull[buf], 3);
            }
          }
          __syncwarp();
          if (++stage == STAGES) stage = 0, phase ^= 1;
        }
      }
    }
  } else {
    // ================= epilogue: one query row per thread (softmax: per thread and column half) =================
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;  // which 128 columns of every 256-column tile
    constexpr int COLS = BN / Epi<SOFTMAX>::HALVES;     // columns per thread and tile
    const int row_local = q * 32 + lane;
    const int row = m0 + row_local;
    // exponent scale in units of the (possibly pre-scaled) TMEM scores
    const float sck = (p.row_sc ? __ldg(p.row_sc + (size_t)b * p.NA + min(m0 + q * 32 + lane, p.NA - 1)) : p.sc) * p.out_scale;
    float run_m = -INFINITY;
    // argmax: number of bit-equal maxima and the sum of their V rows; softmax: (a0, a1) and (a2, s) as packed pairs
    float cnt = 0.f, t0s = 0.f, t1s = 0.f, t2s = 0.f;
    unsigned long long acc01 = 0ull, acc2s = 0ull;
    int run_i = 0;
    const float4* __restrict__ Vg = p.V + (size_t)bphi * p.NB;
    for (int t = 0; t < ntiles; ++t) {
      const int buf = t & 1;
      const uint32_t acc_phase = (t >> 1) & 1;
      tc::mbar_wait(&tfull[buf], acc_phase);
      if (SOFTMAX) tc::mbar_wait(&vfull[buf], acc_phase);
      tc::tc_fence_after();
      const int colbase = (t0 + t) * BN + half * COLS;
      const float4* Vs = v_ring + buf * BN + half * COLS;
#pragma unroll 1
      for (int c0 = 0; c0 < COLS / 32; c0 += 2) {
        if (colbase + c0 * 32 >= p.NB) break;  // warp-uniform
        // two loads in flight per wait: a tcgen05.ld that is waited for alone exposes its whole latency
        uint32_t r2[2][32];
        __syncwarp();  // tcgen05.ld is .sync.aligned: the warp must be converged
        tc::tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + buf * BN + half * COLS + c0 * 32, r2[0]);
        tc::tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + buf * BN + half * COLS + c0 * 32 + 32, r2[1]);
        tc::tmem_ld_wait();
#pragma unroll
      for (int cj = 0; cj < 2; ++cj) {
        const int c = c0 + cj;
        const int cb = colbase + c * 32;
        if (cb >= p.NB) continue;  // warp-uniform
        uint32_t (&r)[32] = r2[cj];
        const int nvalid = min(32, p.NB - cb);
        float cm = -INFINITY;
        if (nvalid == 32) {
#pragma unroll
          for (int i = 0; i < 32; ++i) cm = fmaxf(cm, __uint_as_float(r[i]));
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (i < nvalid) cm = fmaxf(cm, __uint_as_float(r[i]));
        }
        if (!SOFTMAX) {
          if (cm >= run_m) {  // rare once the running maximum has settled
            if (cm > run_m) run_m = cm, cnt = 0.f, t0s = t1s = t2s = 0.f;
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (i < nvalid && __uint_as_float(r[i]) == cm) {
                if (cnt == 0.f) run_i = cb + i;  // lowest index attaining the maximum
                const float4 v = __ldg(Vg + cb + i);
                cnt += 1.f, t0s += v.x, t1s += v.y, t2s += v.z;
              }
          }
        } else {
          if (cm > run_m) {
            const float sc_old = (run_m == -INFINITY) ? 0.f : ex2_approx((run_m - cm) * sck);
            const unsigned long long sc2 = pack2(sc_old, sc_old);
            unsigned long long z = 0ull;
            fma2(z, acc01, sc2), acc01 = z, z = 0ull;
            fma2(z, acc2s, sc2), acc2s = z;
            run_m = cm;
          }
          // weights e = 2^((f - m) * log2(e) / T); V rows come from shared memory as (L, a | b, 1): two packed FMAs
          // accumulate (a0, a1) and (a2, sum of weights)
          if (nvalid == 32) {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const float e = ex2_approx((__uint_as_float(r[i]) - run_m) * sck);
              const ulonglong2 vv = *reinterpret_cast<const ulonglong2*>(Vs + c * 32 + i);
              const unsigned long long e2 = pack2(e, e);
              fma2(acc01, e2, vv.x), fma2(acc2s, e2, vv.y);
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (i < nvalid) {
                const float e = ex2_approx((__uint_as_float(r[i]) - run_m) * sck);
                const ulonglong2 vv = *reinterpret_cast<const ulonglong2*>(Vs + c * 32 + i);
                const unsigned long long e2 = pack2(e, e);
                fma2(acc01, e2, vv.x), fma2(acc2s, e2, vv.y);
              }
          }
        }
      }  // cj
      }
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (CL == 1)
          tc::mbar_arrive(&tempty[buf]);
        else
          tc::mbar_arrive_leader(&tempty[buf]);
        if (SOFTMAX) tc::mbar_arrive(&vempty[buf]);
      }
    }
    if (row < p.NA) {
      SplitOut o;
      o.m = run_m * p.out_scale, o.idx = run_i, o.pad0 = o.pad1 = 0.f;
      if (SOFTMAX) {
        const float2 x01 = unpack2(acc01), x2s = unpack2(acc2s);
        o.s = x2s.y, o.a0 = x01.x, o.a1 = x01.y, o.a2 = x2s.x;
      } else {
        o.s = cnt, o.a0 = t0s, o.a1 = t1s, o.a2 = t2s;
      }
      p.part[((size_t)(blockIdx.z * Epi<SOFTMAX>::HALVES + half) * p.B + b) * p.NA + row] = o;
    }
  }

  tc::tc_fence_before();
  __syncthreads();
  if (CL == 2) tc::cluster_sync_all();  // the leader's MMAs read the peer's shared memory until the very end
  if (warp == 1) {
    tc::tc_fence_after();
    if (CL == 2)
      tc::tmem_dealloc_pair(tmem_base, 512);
    else
      tc::tmem_dealloc(tmem_base, 512);
  }
}

// ---- screened T -> 0 path: one fp16 pass + exact re-scoring of the candidates ----------------------------------------
// At T <= 2e-10 only the row maximum matters (one-hot softmax), and a single hi.hi pass locates it up to the
// rigorous error eps_i of row i (screen_planes_kernel): every column j with f_screen(i, j) >= max_j f_screen - 2 eps_i is a
// CANDIDATE, all others are provably not the maximum.  The screening kernel keeps up to SCREEN_K candidates per (row,
// column-range split) while it streams the tiles -- one third of the MMA work and half of the operand bytes of the
// 3-pass kernel; corr_rescore_kernel then evaluates the candidates exactly in fp32 on the CUDA cores (a few per row) and
// produces (sim, argmax, mean V of bit-equal maxima).  A list that overflows marks its (row, split) for brute force.
constexpr int SCREEN_K = 16;        // candidates kept per (row, column-range split, column half)
constexpr int SCREEN_EPI_WARPS = 8;  // two per TMEM lane quarter, each owning 128 of a tile's 256 columns
constexpr int SCREEN_HALVES = SCREEN_EPI_WARPS / 4;

// The query tile (128 rows x 256 channels of fp16 = 64 KB) stays RESIDENT in shared memory for the CTA's whole sweep over
// the reference positions; only the reference tiles stream through the ring.  The kernel is bound by the L2 -> SM fabric
// (ncu: l1tex__m_xbar2l1tex_read_bytes at 4.6 TB/s with both operands streamed; the exact 3-pass kernel pulls 6.6 TB/s, the
// most any kernel of this library gets out of the L2), so halving the bytes per tile is what shortens it.
template <int CL>
struct ScreenCfg {
  static constexpr int STAGES = CL == 2 ? 8 : 4;
  static constexpr int A_BYTES = BM * 128, B_BYTES = (BN / CL) * 128;
  static constexpr int A_RES_BYTES = 4 * A_BYTES;        // all four k-blocks of the query tile (C = 256)
  static constexpr int STAGE_BYTES = B_BYTES;            // 16384 / 32768
  // candidate lists of the epilogue threads: [SCREEN_K][epilogue threads] column indices -- slot k of thread t lives at
  // [k][t], so dynamic slot indices never conflict on a bank and never touch local memory
  static constexpr int LIST_BYTES = SCREEN_K * SCREEN_EPI_WARPS * 32 * 4;
  static constexpr int SMEM_BYTES = A_RES_BYTES + STAGES * STAGE_BYTES + 1024 + 256 + LIST_BYTES;
};
constexpr int SCREEN_THREADS = 64 + 32 * SCREEN_EPI_WARPS;

struct ScreenParams {
  int NA, NB, B, Bphi, C;
  int tiles_per_split;
  const float* nd_a;   // [B*NA]   ||dropped part|| of every query row
  const float* nh_a;   // [B*NA]   ||hi part||
  const unsigned int* nd_b_max;  // float bits: max over reference rows of ||dropped part||
  const unsigned int* nh_b_max;  //             max over reference rows of ||hi part||
  float* pm;     // [parts][B*NA]            screening maximum of the part (true-score units)
  int* pcnt;     // [parts][B*NA]            number of candidates, or -1: overflow (brute-force the part's columns)
  int* pidx;     // [parts][B*NA][SCREEN_K]  candidate columns
};

// 2 * eps_i in true-score units (see screen_planes_kernel); 64 * 2^-24 covers the truncating TMEM accumulation
__device__ __forceinline__ float screen_threshold(float nd_a, float nh_a, float nd_b, float nh_b) {
  return 2.f * (nd_a * (nh_b + nd_b) + nh_a * nd_b + 4e-6f) * 1.001f;
}

template <int CL>
__global__ void __launch_bounds__(SCREEN_THREADS, 1)
    corr_screen_kernel(const __grid_constant__ CUtensorMap tmAh, const __grid_constant__ CUtensorMap tmBh, const ScreenParams p) {
  using C = ScreenCfg<CL>;
  constexpr int KB = 64;
  constexpr uint32_t IDESC = tc::umma_idesc(0u, BM * CL, BN);
  constexpr int STAGES = C::STAGES, A_BYTES = C::A_BYTES, STAGE_BYTES = C::STAGE_BYTES, A_RES = C::A_RES_BYTES;
  const int crank = (CL == 2) ? (int)tc::cluster_ctarank() : 0;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* ring = smem + A_RES;  // the resident query tile comes first, then the ring of reference tiles
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring + STAGES * STAGE_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + STAGES;
  uint64_t* tfull = bars + 2 * STAGES;
  uint64_t* tempty = bars + 2 * STAGES + 2;
  uint64_t* afull = bars + 2 * STAGES + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 5);
  int* s_ci = reinterpret_cast<int*>(ring + STAGES * STAGE_BYTES + 256);  // [SCREEN_K][epilogue threads] candidate columns

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y;
  const int bphi = (p.Bphi == 1) ? 0 : b;
  const int m0 = blockIdx.x * BM;
  const int ntiles_all = (p.NB + BN - 1) / BN;
  const int t0 = blockIdx.z * p.tiles_per_split;
  const int ntiles = max(min(t0 + p.tiles_per_split, ntiles_all) - t0, 0);
  const int nkb = p.C / KB;

  if (threadIdx.x == 0) {
    tc::tma_prefetch_desc(&tmAh);
    tc::tma_prefetch_desc(&tmBh);
    for (int i = 0; i < STAGES; ++i) tc::mbar_init(&full[i], 1), tc::mbar_init(&empty[i], 1);
    for (int i = 0; i < 2; ++i) tc::mbar_init(&tfull[i], 1), tc::mbar_init(&tempty[i], SCREEN_EPI_WARPS * CL);
    tc::mbar_init(afull, 1);
    tc::fence_barrier_init();
  }
  if (CL == 2) tc::cluster_sync_all();
  if (warp == 1) {
    if (CL == 2) {
      tc::tmem_alloc_pair(tmem_slot, 512);
      tc::tmem_relinquish_pair();
    } else {
      tc::tmem_alloc(tmem_slot, 512);
      tc::tmem_relinquish();
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (CL == 2) tc::cluster_sync_all();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    int stage = 0;
    uint32_t phase = 0;
    if (ntiles > 0 && tc::elect_one()) {  // the query tile: loaded once, resident for the whole sweep
      if (CL == 1) {
        tc::mbar_arrive_expect_tx(afull, A_RES);
        for (int kb = 0; kb < nkb; ++kb) tc::tma_load_2d(smem + kb * A_BYTES, &tmAh, afull, kb * KB, b * p.NA + m0);
      } else {
        if (crank == 0) tc::mbar_arrive_expect_tx(afull, 2 * A_RES);
        for (int kb = 0; kb < nkb; ++kb) tc::tma_load_2d_pair(smem + kb * A_BYTES, &tmAh, afull, kb * KB, b * p.NA + m0);
      }
    }
    __syncwarp();
    for (int t = 0; t < ntiles; ++t) {
      const int col0 = bphi * p.NB + (t0 + t) * BN;
      for (int kb = 0; kb < nkb; ++kb) {
        tc::mbar_wait(&empty[stage], phase ^ 1);
        if (tc::elect_one()) {
          uint8_t* st = ring + stage * STAGE_BYTES;
          if (CL == 1) {
            tc::mbar_arrive_expect_tx(&full[stage], STAGE_BYTES);
            tc::tma_load_2d(st, &tmBh, &full[stage], kb * KB, col0);
          } else {
            if (crank == 0) tc::mbar_arrive_expect_tx(&full[stage], 2 * STAGE_BYTES);
            tc::tma_load_2d_pair(st, &tmBh, &full[stage], kb * KB, col0 + crank * (BN / 2));
          }
        }
        __syncwarp();
        if (++stage == STAGES) stage = 0, phase ^= 1;
      }
    }
  } else if (warp == 1) {
    if (CL == 1 || crank == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = 0; t < ntiles; ++t) {
        const int buf = t & 1;
        if (t == 0) tc::mbar_wait(afull, 0);
        tc::mbar_wait(&tempty[buf], ((t >> 1) & 1) ^ 1);
        tc::tc_fence_after();
        const uint32_t d = tmem_base + buf * BN;
        for (int kb = 0; kb < nkb; ++kb) {
          tc::mbar_wait(&full[stage], phase);
          tc::tc_fence_after();
          const uint64_t dA = tc::umma_desc_k128(tc::smem_u32(smem + kb * A_BYTES));
          const uint64_t dB = tc::umma_desc_k128(tc::smem_u32(ring + stage * STAGE_BYTES));
          if (tc::elect_one()) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              const uint64_t adv = (uint64_t)((kk * 32) >> 4);
              if (CL == 1)
                tc::umma_ss<false>(d, dA + adv, dB + adv, IDESC, (kb | kk) ? 1u : 0u);
              else
                tc::umma_ss_pair<false>(d, dA + adv, dB + adv, IDESC, (kb | kk) ? 1u : 0u);
            }
            if (CL == 1) {
              tc::umma_commit(&empty[stage]);
              if (kb == nkb - 1) tc::umma_commit(&tfull[buf]);
            } else {
              tc::umma_commit_pair_mc(&empty[stage], 3);
              if (kb == nkb - 1) tc::umma_commit_pair_mc(&tfull[buf], 3);
            }
          }
          __syncwarp();
          if (++stage == STAGES) stage = 0, phase ^= 1;
        }
      }
    }
  } else {
    // ================= epilogue: one query row x one column half per thread =================
    // Per tile a thread drains its 128 columns with four tcgen05.ld in flight before one wait (a load that is waited for
    // alone exposes its full latency: the first version, one x32 load per wait, spent 7.5 K cycles per tile on a 2 K-cycle
    // MMA tile -- as did the exact 3-pass kernel's epilogue, which hid behind its 6 K cycles of MMAs), then 16 three-input
    // maxima per 32-column chunk.  A chunk whose maximum comes within the threshold of the row's running maximum (a
    // "record" or a near-tie: ~ln(#chunks) times per row, but for SOME lane of a warp in about every second chunk) builds
    // a 32-bit mask of its qualifying columns and appends their indices to the thread's list in shared memory
    // ([slot][thread]: no bank conflicts, no local memory).  Values are not kept: a record that beats the previous maximum
    // by more than the threshold disqualifies the whole list at once (every entry is <= the previous maximum); otherwise
    // the old entries stay -- at worst a few extra candidates for the exact re-scoring, never a missing one.
    const int q = warp & 3, half = (warp - 2) >> 2;
    constexpr int COLS = BN / SCREEN_HALVES;   // 128
    constexpr int LT = SCREEN_EPI_WARPS * 32;  // list stride
    const int et = threadIdx.x - 64;           // epilogue thread index
    const int row = m0 + q * 32 + lane;
    const size_t grow = (size_t)b * p.NA + min(row, p.NA - 1);
    // candidate threshold in TMEM units (scores there are true scores * 2^28)
    const float thr = screen_threshold(__ldg(p.nd_a + grow), __ldg(p.nh_a + grow), __uint_as_float(__ldg(p.nd_b_max)),
                                       __uint_as_float(__ldg(p.nh_b_max))) * 268435456.0f;
    float run_m = -INFINITY;
    int cnt = 0;  // entries appended (only the first SCREEN_K are stored: cnt > SCREEN_K = overflow)
    for (int t = 0; t < ntiles; ++t) {
      const int buf = t & 1;
      tc::mbar_wait(&tfull[buf], (t >> 1) & 1);
      tc::tc_fence_after();
      const int colbase = (t0 + t) * BN + half * COLS;
      if (colbase < p.NB) {  // warp-uniform
        uint32_t r[COLS / 32][32];
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * BN + half * COLS;
        __syncwarp();
#pragma unroll
        for (int c = 0; c < COLS / 32; ++c) tc::tmem_ld_32x32(taddr + c * 32, r[c]);
        tc::tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < COLS / 32; ++c) {
          const int cb = colbase + c * 32;
          const int nvalid = p.NB - cb;  // columns at or beyond NB hold zero-filled (TMA) operands: mask them
          if (nvalid < 32) {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (i >= nvalid) r[c][i] = __float_as_uint(-INFINITY);
          }
          float cm = -INFINITY;
#pragma unroll
          for (int i = 0; i < 32; ++i) cm = fmaxf(cm, __uint_as_float(r[c][i]));
          if (cm - thr > run_m) cnt = 0;  // a record that disqualifies every earlier entry (all <= the old maximum)
          run_m = fmaxf(run_m, cm);
          const float lim = run_m - thr;
          if (cm >= lim && cm > -INFINITY) {
            uint32_t mask = 0;
#pragma unroll
            for (int i = 0; i < 32; ++i) mask |= (__uint_as_float(r[c][i]) >= lim ? 1u : 0u) << i;
            while (mask) {
              const int i = __ffs(mask) - 1;
              mask &= mask - 1;
              if (cnt < SCREEN_K) s_ci[cnt * LT + et] = cb + i;
              ++cnt;
            }
            if (cnt > SCREEN_K) cnt = SCREEN_K + 1;  // overflow (sticky until a disqualifying record clears the list)
          }
        }
      }
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (CL == 1)
          tc::mbar_arrive(&tempty[buf]);
        else
          tc::mbar_arrive_leader(&tempty[buf]);
      }
    }
    if (row < p.NA) {
      const size_t o = ((size_t)(blockIdx.z * SCREEN_HALVES + half) * p.B + b) * p.NA + row;
      p.pm[o] = run_m * 3.725290298461914e-09f;
      const bool overflow = cnt > SCREEN_K;
      if (!overflow)
        for (int j = 0; j < cnt; ++j) p.pidx[o * SCREEN_K + j] = s_ci[j * LT + et];
      p.pcnt[o] = overflow ? -1 : cnt;
    }
  }

  tc::tc_fence_before();
  __syncthreads();
  if (CL == 2) tc::cluster_sync_all();
  if (warp == 1) {
    tc::tc_fence_after();
    if (CL == 2)
      tc::tmem_dealloc_pair(tmem_base, 512);
    else
      tc::tmem_dealloc(tmem_base, 512);
  }
}

// Exact re-scoring: one warp per query row.  fp32 dot products of the ORIGINAL fp32 operands (lane l owns dimensions
// 4l..4l+3 and 128+4l..128+4l+3, eight FMAs, then a butterfly sum that leaves the same bits in every lane), running
// (max, lowest index, number of bit-equal maxima, sum of their V rows) exactly like the 3-pass kernel's epilogue.
__global__ void __launch_bounds__(256) corr_rescore_kernel(const float* __restrict__ theta, const float* __restrict__ phi,
                                                           const float4* __restrict__ V, const ScreenParams p, int nparts,
                                                           float4* __restrict__ y, float* __restrict__ sim, int* __restrict__ argmax,
                                                           const CorrPeers peers) {
  const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int rows = p.B * p.NA;
  if (r >= rows) return;
  const int b = r / p.NA;
  const int bphi = (p.Bphi == 1) ? 0 : b;
  const float* ph = phi + (size_t)bphi * p.NB * 256;
  const float4* Vg = V + (size_t)bphi * p.NB;
  const float4 a0 = __ldg(reinterpret_cast<const float4*>(theta + (size_t)r * 256) + lane);
  const float4 a1 = __ldg(reinterpret_cast<const float4*>(theta + (size_t)r * 256) + 32 + lane);
  auto score = [&](int col) {
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(ph + (size_t)col * 256) + lane);
    const float4 b1 = __ldg(reinterpret_cast<const float4*>(ph + (size_t)col * 256) + 32 + lane);
    float s0 = a0.x * b0.x, s1 = a1.x * b1.x;
    s0 = fmaf(a0.y, b0.y, s0), s1 = fmaf(a1.y, b1.y, s1);
    s0 = fmaf(a0.z, b0.z, s0), s1 = fmaf(a1.z, b1.z, s1);
    s0 = fmaf(a0.w, b0.w, s0), s1 = fmaf(a1.w, b1.w, s1);
    float sacc = s0 + s1;
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) sacc += __shfl_xor_sync(0xffffffffu, sacc, off);
    return sacc;
  };
  float pmax = -INFINITY;
  for (int s = 0; s < nparts; ++s) pmax = fmaxf(pmax, __ldg(p.pm + (size_t)s * rows + r));
  const float thr = screen_threshold(__ldg(p.nd_a + r), __ldg(p.nh_a + r), __uint_as_float(__ldg(p.nd_b_max)), __uint_as_float(__ldg(p.nh_b_max)));
  float m = -INFINITY, cnt = 0.f, t0 = 0.f, t1 = 0.f, t2 = 0.f;
  int idx = 0x7fffffff;
  auto visit = [&](int col) {
    const float f = score(col);
    if (f >= m) {
      const float4 v = __ldg(Vg + col);
      if (f > m) m = f, cnt = 0.f, t0 = t1 = t2 = 0.f, idx = col;
      idx = min(idx, col), cnt += 1.f, t0 += v.x, t1 += v.y, t2 += v.z;
    }
  };
  const int ntiles_all = (p.NB + BN - 1) / BN;
  for (int s = 0; s < nparts; ++s) {
    const size_t o = (size_t)s * rows + r;
    if (__ldg(p.pm + o) < pmax - thr) continue;  // nothing in this part can be the maximum
    const int n = __ldg(p.pcnt + o);
    if (n >= 0) {
      for (int j = 0; j < n; ++j) visit(__ldg(p.pidx + o * SCREEN_K + j));
    } else {  // overflowed list: every column of the part's range (part = column-range split x column half of each tile)
      const int sp = s / SCREEN_HALVES, hf = s - sp * SCREEN_HALVES, hc = BN / SCREEN_HALVES;
      const int tl0 = sp * p.tiles_per_split, tl1 = min((sp + 1) * p.tiles_per_split, ntiles_all);
      for (int tl = tl0; tl < tl1; ++tl) {
        const int c0 = tl * BN + hf * hc, c1 = min(c0 + hc, p.NB);
        for (int col = c0; col < c1; ++col) visit(col);
      }
    }
  }
  if (lane == 0) {
    float4 v;
    if (cnt == 1.f) {
      v = __ldg(Vg + idx);
    } else {
      const float inv = 1.f / cnt;
      v = make_float4(t0 * inv, t1 * inv, t2 * inv, 0.f);
    }
    y[r] = make_float4(v.x, v.y, v.z, 0.f);
    sim[r] = m;
    if (argmax) argmax[r] = idx;
    for (int g = 0; g < peers.n; ++g) {
      reinterpret_cast<float4*>(peers.y4[g])[peers.row0 + r] = make_float4(v.x, v.y, v.z, 0.f);
      peers.sim[g][peers.row0 + r] = m;
    }
  }
}

// ---- merge the column-range splits ----------------------------------------------------------------------
template <bool SOFTMAX>
__global__ void __launch_bounds__(256) corr_merge_kernel(const SplitOut* __restrict__ part, int nsplit, int rows, int NA,
                                                         int NB, int Bphi, float sc_all, const float* __restrict__ row_sc,
                                                         const float4* __restrict__ V, float4* __restrict__ y, float* __restrict__ sim,
                                                         int* __restrict__ argmax, float* __restrict__ denom, const CorrPeers peers) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  const float sc = row_sc ? row_sc[r] : sc_all;
  const int b = r / NA;
  const float4* Vg = V + (size_t)((Bphi == 1) ? 0 : b) * NB;
  if (!SOFTMAX) {
    float m = -INFINITY;
    for (int s = 0; s < nsplit; ++s) m = fmaxf(m, part[(size_t)s * rows + r].m);
    // mean of the V rows of all bit-equal maxima (one row, exactly, when the maximum is unique)
    int idx = 0x7fffffff;
    float cnt = 0.f, a0 = 0.f, a1 = 0.f, a2 = 0.f;
    for (int s = 0; s < nsplit; ++s) {
      const SplitOut o = part[(size_t)s * rows + r];
      if (o.m == m && o.s > 0.f) cnt += o.s, a0 += o.a0, a1 += o.a1, a2 += o.a2, idx = min(idx, o.idx);
    }
    float4 v;
    if (cnt == 1.f) {
      v = __ldg(Vg + idx);
    } else {
      const float inv = 1.f / cnt;
      v = make_float4(a0 * inv, a1 * inv, a2 * inv, 0.f);
    }
    y[r] = make_float4(v.x, v.y, v.z, 0.f);
    sim[r] = m;
    if (argmax) argmax[r] = idx;
    for (int g = 0; g < peers.n; ++g) {  // fused all-gather: the row goes to every GPU's full-size result
      reinterpret_cast<float4*>(peers.y4[g])[peers.row0 + r] = make_float4(v.x, v.y, v.z, 0.f);
      peers.sim[g][peers.row0 + r] = m;
    }
  } else {
    float m = -INFINITY;
    for (int s = 0; s < nsplit; ++s) m = fmaxf(m, part[(size_t)s * rows + r].m);
    float ssum = 0.f, a0 = 0.f, a1 = 0.f, a2 = 0.f;
    for (int s = 0; s < nsplit; ++s) {
      const SplitOut o = part[(size_t)s * rows + r];
      if (o.m == -INFINITY) continue;
      const float w = exp2f((o.m - m) * sc);
      ssum += w * o.s, a0 += w * o.a0, a1 += w * o.a1, a2 += w * o.a2;
    }
    y[r] = make_float4(a0 / ssum, a1 / ssum, a2 / ssum, 0.f);
    sim[r] = m;
    if (denom) denom[r] = ssum;  // sum_j exp((f_ij - m_i) / T_i)
    if (argmax) argmax[r] = -1;
    for (int g = 0; g < peers.n; ++g) {
      reinterpret_cast<float4*>(peers.y4[g])[peers.row0 + r] = make_float4(a0 / ssum, a1 / ssum, a2 / ssum, 0.f);
      peers.sim[g][peers.row0 + r] = m;
    }
  }
}

int ws_get(CorrWorkspace* ws, int i, size_t bytes, void** out) {
  if (ws->cap[i] < bytes) {  // growth outside dvc_set_exemplar's reservation: a stand-alone call with a new shape
    if (ws->buf[i]) cudaFree(ws->buf[i]);
    ws->buf[i] = nullptr, ws->cap[i] = 0;
    if (i == 2 || i == 3 || i == 5) ws->phi_src = nullptr, ws->phi_version = -1;
    if (cudaMalloc(&ws->buf[i], bytes) != cudaSuccess) return -1;
    ws->cap[i] = bytes;
  }
  *out = ws->buf[i];
  return 0;
}

template <int FMT, bool SOFTMAX, int CL>
int launch_main_cl(const CUtensorMap& mAh, const CUtensorMap& mAl, const CUtensorMap& mBh, const CUtensorMap& mBl,
                   const TcParams& tp, dim3 grid, cudaStream_t s) {
  static unsigned long long attr_mask = 0;  // the attribute is per device
  if (first_use_on_device(&attr_mask)) {
    if (cudaFuncSetAttribute(corr_tc_kernel<FMT, SOFTMAX, CL>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<CL>::SMEM_BYTES) !=
        cudaSuccess)
      return -1;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid, cfg.blockDim = dim3(Epi<SOFTMAX>::NTHREADS), cfg.dynamicSmemBytes = Cfg<CL>::SMEM_BYTES, cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CL, at[0].val.clusterDim.y = 1, at[0].val.clusterDim.z = 1;
  cfg.attrs = at, cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, corr_tc_kernel<FMT, SOFTMAX, CL>, mAh, mAl, mBh, mBl, tp) == cudaSuccess ? 0 : -2;
}
template <int FMT, bool SOFTMAX>
int launch_main(const CUtensorMap& mAh, const CUtensorMap& mAl, const CUtensorMap& mBh, const CUtensorMap& mBl,
                const TcParams& tp, dim3 grid, int cl, cudaStream_t s) {
  return cl == 2 ? launch_main_cl<FMT, SOFTMAX, 2>(mAh, mAl, mBh, mBl, tp, grid, s)
                 : launch_main_cl<FMT, SOFTMAX, 1>(mAh, mAl, mBh, mBl, tp, grid, s);
}

}  // namespace

bool first_use_on_device(unsigned long long* mask) {
  static std::mutex mu;
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lk(mu);
  const unsigned long long bit = 1ull << (dev & 63);
  if (*mask & bit) return false;
  *mask |= bit;
  return true;
}

static size_t screen_norm_bytes(int B, int Bphi, int NA, int NB) { return ((size_t)2 * B * NA + (size_t)2 * Bphi * NB + 8) * 4; }
static size_t screen_cand_bytes(int nparts, int B, int NA) { return (size_t)nparts * B * NA * (8 + 4 * SCREEN_K); }

int corr_ws_reserve(CorrWorkspace* ws, int B, int Bphi, int NA, int NB) {
  void* d;
  const size_t ea = (size_t)B * NA * 256 * 4, ephi = (size_t)Bphi * NB * 256 * 4;  // tf32 words: the widest format
  const size_t part = (size_t)16 * 2 * B * NA * sizeof(SplitOut);                   // at most 16 column splits x 2 column halves
  if (ws_get(ws, 0, ea, &d) || ws_get(ws, 1, ea, &d) || ws_get(ws, 2, ephi, &d) || ws_get(ws, 3, ephi, &d) || ws_get(ws, 4, part, &d) ||
      ws_get(ws, 5, screen_norm_bytes(B, Bphi, NA, NB), &d) || ws_get(ws, 6, screen_cand_bytes(16 * SCREEN_HALVES, B, NA), &d))
    return -1;
  return 0;
}

void corr_ws_free(CorrWorkspace* ws) {
  for (int i = 0; i < CorrWorkspace::NBUF; ++i) {
    if (ws->buf[i]) cudaFree(ws->buf[i]);
    ws->buf[i] = nullptr, ws->cap[i] = 0;
  }
  ws->phi_src = nullptr, ws->phi_version = -1;
}

template <int CL>
static int launch_screen_cl(const CUtensorMap& mA, const CUtensorMap& mB, const ScreenParams& sp, dim3 grid, cudaStream_t s) {
  static unsigned long long attr_mask = 0;
  if (first_use_on_device(&attr_mask)) {
    if (cudaFuncSetAttribute(corr_screen_kernel<CL>, cudaFuncAttributeMaxDynamicSharedMemorySize, ScreenCfg<CL>::SMEM_BYTES) != cudaSuccess)
      return -1;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid, cfg.blockDim = dim3(SCREEN_THREADS), cfg.dynamicSmemBytes = ScreenCfg<CL>::SMEM_BYTES, cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CL, at[0].val.clusterDim.y = 1, at[0].val.clusterDim.z = 1;
  cfg.attrs = at, cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, corr_screen_kernel<CL>, mA, mB, sp) == cudaSuccess ? 0 : -2;
}

int launch_corr_tc(const CorrParams& p, int math, int cluster, int screen, CorrWorkspace* ws, long long phi_version,
                   cudaStream_t s, std::string* err) {
  auto fail = [&](const char* m) {
    if (err) *err = m;
    return -1;
  };
  const bool tf32 = (math == 1);
  const int fmt = math == 1 ? 0 : (math == 2 ? 1 : 2);  // DVC_MATH_TF32X3 / BF16X3 / FP16X3
  if (p.C < 64 || p.C % 64 || p.C > 4096) return fail("C must be a multiple of 64 (<= 4096)");
  const int eb = tf32 ? 4 : 2;
  const size_t ea = (size_t)p.B * p.NA * p.C, ephi = (size_t)p.Bphi * p.NB * p.C;
  void *Ah, *Al, *Bh, *Bl, *part;
  if (ws_get(ws, 0, ea * eb, &Ah) || ws_get(ws, 1, ea * eb, &Al) || ws_get(ws, 2, ephi * eb, &Bh) || ws_get(ws, 3, ephi * eb, &Bl))
    return fail("workspace allocation failed");

  // column-range splits so that (row blocks x batch x splits) fills the 148 SMs in whole waves
  const int cl = cluster == 2 ? 2 : 1;
  const int row_blocks = ((p.NA + BM - 1) / BM + cl - 1) / cl * cl;  // pairs: an even number of 128-row query tiles
  const int ntiles = (p.NB + BN - 1) / BN;
  int nsplit = 1;
  {
    const long ctas = (long)row_blocks * p.B;
    double best = -1.0;
    for (int sp = 1; sp <= 16 && sp <= ntiles; ++sp) {
      const int tps = (ntiles + sp - 1) / sp;
      const int eff_sp = (ntiles + tps - 1) / tps;
      const long total = ctas * eff_sp;
      const long waves = (total + 147) / 148;
      const double eff = (double)total / (double)(waves * 148) * ((double)ntiles / (double)(tps * eff_sp)) -
                         0.01 * sp;  // mild penalty: every split re-runs the prologue
      if (eff > best) best = eff, nsplit = eff_sp;
    }
  }
  const int tps = (ntiles + nsplit - 1) / nsplit;
  nsplit = (ntiles + tps - 1) / tps;
  const bool softmax = !(p.temperature <= 2e-10f) || p.row_scale != nullptr;
  if (screen && fmt == 2 && !softmax && p.C == 256) {  // (the resident query tile of the screening kernel is sized for C = 256)
    // ---- screened T -> 0 path: hi planes + error norms, one fp16 pass, exact re-scoring of the candidates ----
    const int rows = p.B * p.NA, rphi = p.Bphi * p.NB;
    void *norms, *cand;
    const int sparts = nsplit * SCREEN_HALVES;
    if (ws_get(ws, 5, screen_norm_bytes(p.B, p.Bphi, p.NA, p.NB), &norms) || ws_get(ws, 6, screen_cand_bytes(sparts, p.B, p.NA), &cand))
      return fail("workspace allocation failed");
    float* nd_a = (float*)norms;
    float* nh_a = nd_a + rows;
    float* nd_b = nh_a + rows;
    float* nh_b = nd_b + rphi;
    unsigned int* cells = (unsigned int*)(nh_b + rphi);  // [0,1]: query side (unused maxima), [2,3]: reference side
    const bool phi_cached = phi_version >= 0 && ws->phi_src == p.phi && ws->phi_version == phi_version && ws->phi_fmt == 3 &&
                            ws->phi_elems == ephi;
    if (cudaMemsetAsync(cells, 0, (phi_cached ? 2 : 4) * sizeof(unsigned int), s) != cudaSuccess) return fail("cudaMemsetAsync failed");
    screen_planes_kernel<<<(rows * 32 + 255) / 256, 256, 0, s>>>(p.theta, (__half*)Ah, nd_a, nh_a, cells + 0, cells + 1, rows);
    if (!phi_cached) screen_planes_kernel<<<(rphi * 32 + 255) / 256, 256, 0, s>>>(p.phi, (__half*)Bh, nd_b, nh_b, cells + 2, cells + 3, rphi);
    launch_counter_add(phi_cached ? 1 : 2);
    ws->phi_src = phi_version >= 0 ? p.phi : nullptr, ws->phi_version = phi_version, ws->phi_fmt = 3, ws->phi_elems = ephi;
    CUtensorMap mA, mB;
    if (encode_tmap_2d(&mA, Ah, (uint64_t)rows, p.C, BM, 64, 2) || encode_tmap_2d(&mB, Bh, (uint64_t)rphi, p.C, BN / cl, 64, 2))
      return fail("cuTensorMapEncodeTiled failed");
    ScreenParams sp;
    sp.NA = p.NA, sp.NB = p.NB, sp.B = p.B, sp.Bphi = p.Bphi, sp.C = p.C, sp.tiles_per_split = tps;
    sp.nd_a = nd_a, sp.nh_a = nh_a, sp.nd_b_max = cells + 2, sp.nh_b_max = cells + 3;
    sp.pm = (float*)cand;
    sp.pcnt = (int*)(sp.pm + (size_t)sparts * rows);
    sp.pidx = sp.pcnt + (size_t)sparts * rows;
    dim3 grid(row_blocks, p.B, nsplit);
    const int rc = cl == 2 ? launch_screen_cl<2>(mA, mB, sp, grid, s) : launch_screen_cl<1>(mA, mB, sp, grid, s);
    if (rc) return fail(rc == -1 ? "cudaFuncSetAttribute(max dynamic smem) failed" : "cudaLaunchKernelEx failed");
    corr_rescore_kernel<<<(rows * 32 + 255) / 256, 256, 0, s>>>(p.theta, p.phi, reinterpret_cast<const float4*>(p.V), sp, sparts,
                                                                reinterpret_cast<float4*>(p.y), p.sim, p.argmax, p.peers);
    launch_counter_add(2);
    return 0;
  }
  const int nparts = nsplit * Epi<true>::HALVES;  // partial rows the merge kernel combines
  if (ws_get(ws, 4, (size_t)nparts * p.B * p.NA * sizeof(SplitOut), &part)) return fail("workspace allocation failed");

  const int grid1 = 148 * 8;
  // the reference side's planes survive from launch to launch while (pointer, version, format, size) are unchanged
  const bool phi_cached = phi_version >= 0 && ws->phi_src == p.phi && ws->phi_version == phi_version && ws->phi_fmt == fmt &&
                          ws->phi_elems == ephi;
  if (fmt == 0) {
    split_planes_kernel<0><<<grid1, 256, 0, s>>>(p.theta, Ah, Al, ea / 4);
    if (!phi_cached) split_planes_kernel<0><<<grid1, 256, 0, s>>>(p.phi, Bh, Bl, ephi / 4);
  } else if (fmt == 1) {
    split_planes_kernel<1><<<grid1, 256, 0, s>>>(p.theta, Ah, Al, ea / 4);
    if (!phi_cached) split_planes_kernel<1><<<grid1, 256, 0, s>>>(p.phi, Bh, Bl, ephi / 4);
  } else {
    split_planes_kernel<2><<<grid1, 256, 0, s>>>(p.theta, Ah, Al, ea / 4);
    if (!phi_cached) split_planes_kernel<2><<<grid1, 256, 0, s>>>(p.phi, Bh, Bl, ephi / 4);
  }
  launch_counter_add(phi_cached ? 1 : 2);
  ws->phi_src = phi_version >= 0 ? p.phi : nullptr, ws->phi_version = phi_version, ws->phi_fmt = fmt, ws->phi_elems = ephi;

  CUtensorMap mAh, mAl, mBh, mBl;
  const uint32_t boxk = tf32 ? 32 : 64;
  if (encode_tmap_2d(&mAh, Ah, (uint64_t)p.B * p.NA, p.C, BM, boxk, eb) || encode_tmap_2d(&mAl, Al, (uint64_t)p.B * p.NA, p.C, BM, boxk, eb) ||
      encode_tmap_2d(&mBh, Bh, (uint64_t)p.Bphi * p.NB, p.C, BN / cl, boxk, eb) ||
      encode_tmap_2d(&mBl, Bl, (uint64_t)p.Bphi * p.NB, p.C, BN / cl, boxk, eb))
    return fail("cuTensorMapEncodeTiled failed");

  TcParams tp;
  tp.NA = p.NA, tp.NB = p.NB, tp.B = p.B, tp.Bphi = p.Bphi, tp.C = p.C, tp.tiles_per_split = tps;
  tp.sc = 1.4426950408889634f / p.temperature;
  tp.row_sc = p.row_scale;
  tp.out_scale = fmt == 2 ? 3.725290298461914e-09f /* 2^-28 */ : 1.0f;
  tp.V = reinterpret_cast<const float4*>(p.V);
  tp.part = reinterpret_cast<SplitOut*>(part);
  dim3 grid(row_blocks, p.B, nsplit);
  int rc;
  if (fmt == 0)
    rc = softmax ? launch_main<0, true>(mAh, mAl, mBh, mBl, tp, grid, cl, s) : launch_main<0, false>(mAh, mAl, mBh, mBl, tp, grid, cl, s);
  else if (fmt == 1)
    rc = softmax ? launch_main<1, true>(mAh, mAl, mBh, mBl, tp, grid, cl, s) : launch_main<1, false>(mAh, mAl, mBh, mBl, tp, grid, cl, s);
  else
    rc = softmax ? launch_main<2, true>(mAh, mAl, mBh, mBl, tp, grid, cl, s) : launch_main<2, false>(mAh, mAl, mBh, mBl, tp, grid, cl, s);
  if (rc) return fail(rc == -1 ? "cudaFuncSetAttribute(max dynamic smem) failed" : "cudaLaunchKernelEx failed");
  launch_counter_add(1);
  const int rows = p.B * p.NA;
  if (softmax)
    corr_merge_kernel<true><<<(rows + 255) / 256, 256, 0, s>>>(tp.part, nparts, rows, p.NA, p.NB, p.Bphi, tp.sc, p.row_scale, tp.V,
                                                                reinterpret_cast<float4*>(p.y), p.sim, p.argmax, p.denom, p.peers);
  else
    corr_merge_kernel<false><<<(rows + 255) / 256, 256, 0, s>>>(tp.part, nparts, rows, p.NA, p.NB, p.Bphi, tp.sc, nullptr, tp.V,
                                                                 reinterpret_cast<float4*>(p.y), p.sim, p.argmax, nullptr, p.peers);
  launch_counter_add(1);
  return 0;
}

}  // namespace dvc
