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
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*alignment slack*/ + 256 /*barriers*/;
};
constexpr int NTHREADS = 192;

struct SplitOut {  // per (split, row) partial statistics
  float m, s, a0, a1, a2;
  int idx;
  float pad0, pad1;
};

struct TcParams {
  int NA, NB, B, Bphi, C;
  int tiles_per_split;
  float sc;  // log2(e) / T
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

// ---- main kernel ---------------------------------------------------------------------------------------
template <int FMT, bool SOFTMAX, int CL>
__global__ void __launch_bounds__(NTHREADS, 1)
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
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

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
    for (int i = 0; i < 2; ++i) tc::mbar_init(&tfull[i], 1), tc::mbar_init(&tempty[i], 4 * CL);  // pair: both epilogues
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
    // ================= epilogue: one query row per thread =================
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    const int row_local = q * 32 + lane;
    const int row = m0 + row_local;
    const float sck = p.sc * p.out_scale;  // exponent scale in units of the (possibly pre-scaled) TMEM scores
    float run_m = -INFINITY, run_s = 0.f, a0 = 0.f, a1 = 0.f, a2 = 0.f;
    int run_i = 0;
    const float4* __restrict__ Vg = p.V + (size_t)bphi * p.NB;
    for (int t = 0; t < ntiles; ++t) {
      const int buf = t & 1;
      const uint32_t acc_phase = (t >> 1) & 1;
      tc::mbar_wait(&tfull[buf], acc_phase);
      tc::tc_fence_after();
      const int colbase = (t0 + t) * BN;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        const int cb = colbase + c * 32;
        if (cb >= p.NB) break;  // warp-uniform
        uint32_t r[32];
        __syncwarp();  // tcgen05.ld is .sync.aligned: the warp must be converged
        tc::tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + buf * BN + c * 32, r);
        tc::tmem_ld_wait();
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
          if (cm > run_m) {  // rare once the running maximum has settled
            run_m = cm;
#pragma unroll
            for (int i = 31; i >= 0; --i)
              if (i < nvalid && __uint_as_float(r[i]) == cm) run_i = cb + i;  // lowest index wins
          }
        } else {
          if (cm > run_m) {
            const float sc_old = (run_m == -INFINITY) ? 0.f : exp2f((run_m - cm) * sck);
            run_s *= sc_old, a0 *= sc_old, a1 *= sc_old, a2 *= sc_old;
            run_m = cm;
          }
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            if (i < nvalid) {
              const float e = exp2f((__uint_as_float(r[i]) - run_m) * sck);
              const float4 v = __ldg(Vg + cb + i);  // same address across the warp: one broadcast load
              run_s += e;
              a0 = fmaf(e, v.x, a0), a1 = fmaf(e, v.y, a1), a2 = fmaf(e, v.z, a2);
            }
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
      SplitOut o;
      o.m = run_m * p.out_scale, o.s = run_s, o.a0 = a0, o.a1 = a1, o.a2 = a2, o.idx = run_i, o.pad0 = o.pad1 = 0.f;
      p.part[((size_t)blockIdx.z * p.B + b) * p.NA + row] = o;
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

// ---- merge the column-range splits ----------------------------------------------------------------------
template <bool SOFTMAX>
__global__ void __launch_bounds__(256) corr_merge_kernel(const SplitOut* __restrict__ part, int nsplit, int rows, int NA,
                                                         int NB, int Bphi, float sc, const float4* __restrict__ V,
                                                         float4* __restrict__ y, float* __restrict__ sim,
                                                         int* __restrict__ argmax, const CorrPeers peers) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  const int b = r / NA;
  const float4* Vg = V + (size_t)((Bphi == 1) ? 0 : b) * NB;
  if (!SOFTMAX) {
    float m = -INFINITY;
    int idx = 0;
    for (int s = 0; s < nsplit; ++s) {
      const SplitOut o = part[(size_t)s * rows + r];
      if (o.m > m || (o.m == m && o.idx < idx)) m = o.m, idx = o.idx;
    }
    const float4 v = __ldg(Vg + idx);
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
    if (i == 2 || i == 3) ws->phi_src = nullptr, ws->phi_version = -1;
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
  cfg.gridDim = grid, cfg.blockDim = dim3(NTHREADS), cfg.dynamicSmemBytes = Cfg<CL>::SMEM_BYTES, cfg.stream = s;
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

int corr_ws_reserve(CorrWorkspace* ws, int B, int Bphi, int NA, int NB) {
  void* d;
  const size_t ea = (size_t)B * NA * 256 * 4, ephi = (size_t)Bphi * NB * 256 * 4;  // tf32 words: the widest format
  const size_t part = (size_t)16 * B * NA * sizeof(SplitOut);                       // at most 16 column splits
  if (ws_get(ws, 0, ea, &d) || ws_get(ws, 1, ea, &d) || ws_get(ws, 2, ephi, &d) || ws_get(ws, 3, ephi, &d) || ws_get(ws, 4, part, &d))
    return -1;
  return 0;
}

void corr_ws_free(CorrWorkspace* ws) {
  for (int i = 0; i < 5; ++i) {
    if (ws->buf[i]) cudaFree(ws->buf[i]);
    ws->buf[i] = nullptr, ws->cap[i] = 0;
  }
  ws->phi_src = nullptr, ws->phi_version = -1;
}

int launch_corr_tc(const CorrParams& p, int math, int cluster, CorrWorkspace* ws, long long phi_version, cudaStream_t s,
                   std::string* err) {
  auto fail = [&](const char* m) {
    if (err) *err = m;
    return -1;
  };
  const bool tf32 = (math == 1);
  const int fmt = math == 1 ? 0 : (math == 2 ? 1 : 2);  // DVC_MATH_TF32X3 / BF16X3 / FP16X3
  if (p.C != 256) return fail("C must be 256");
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
  if (ws_get(ws, 4, (size_t)nsplit * p.B * p.NA * sizeof(SplitOut), &part)) return fail("workspace allocation failed");

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
  tp.out_scale = fmt == 2 ? 3.725290298461914e-09f /* 2^-28 */ : 1.0f;
  tp.V = reinterpret_cast<const float4*>(p.V);
  tp.part = reinterpret_cast<SplitOut*>(part);
  dim3 grid(row_blocks, p.B, nsplit);
  const bool softmax = !(p.temperature <= 2e-10f);
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
    corr_merge_kernel<true><<<(rows + 255) / 256, 256, 0, s>>>(tp.part, nsplit, rows, p.NA, p.NB, p.Bphi, tp.sc, tp.V,
                                                                reinterpret_cast<float4*>(p.y), p.sim, p.argmax, p.peers);
  else
    corr_merge_kernel<false><<<(rows + 255) / 256, 256, 0, s>>>(tp.part, nsplit, rows, p.NA, p.NB, p.Bphi, tp.sc, tp.V,
                                                                 reinterpret_cast<float4*>(p.y), p.sim, p.argmax, p.peers);
  launch_counter_add(1);
  return 0;
}

}  // namespace dvc
