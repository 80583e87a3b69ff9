// tcgen05 GEMM of the MMDiT linears (sm_100a):  y = A W^T (+bias, fused epilogue), fp32 accumulation in TMEM.
//
//   operands   bf16 planes, K-major.  NSPLIT == 1: y = A_hi W_hi^T.  NSPLIT == 3 ("bf16x3", fp32-faithful to ~2^-17):
//              y = A_hi W_hi^T + A_hi W_lo^T + A_lo W_hi^T, all three products accumulated into the same TMEM tile.
//   tile       128 x 256 x 64 per pipeline stage, UMMA 128x256x16 (kind::f16, cta_group::1)
//   staging    TMA (cp.async.bulk.tensor.2d, SWIZZLE_128B) -> shared memory ring, mbarrier full/empty pairs
//   roles      warp 0: TMA producer (1 thread) | warp 1: MMA issuer (1 thread) | warps 2-9: epilogue (TMEM -> regs -> smem transpose -> HBM)
//   TMEM       512 columns = 2 accumulator tiles; the epilogue of tile i overlaps the MMAs of tile i+1
//   schedule   persistent CTAs (grid = #SMs), tiles rasterised in groups of 8 M-blocks for L2 reuse of W
//
// Replaces the cuBLAS SGEMMs behind nn.Linear in DismantledBlock (sd3/mmdit.py:266,269,293,301; other_impls.py:82-84)
// together with the elementwise kernels around them (bias, GELU-tanh, gate*y + residual; mmdit.py:485-496).
#include "common.cuh"
#include "kernels.h"

#include <cuda.h>   // CUtensorMap types only; the driver entry point is resolved at run time (no -lcuda)
#include <stdlib.h>

namespace stk {

namespace {

constexpr int BM = 128, BN = 256, BK = 64, UMMA_K = 16;
constexpr int A_TILE_BYTES = BM * BK * 2;      // 16 KiB
constexpr int B_TILE_BYTES = BN * BK * 2;      // 32 KiB
constexpr int NUM_THREADS = 64 + 8 * 32;     // TMA warp, MMA warp, 8 epilogue warps
constexpr int TMEM_COLS = 512;

template <int NSPLIT> struct Cfg {
  static constexpr int PLANES = NSPLIT == 3 ? 2 : 1;
  static constexpr int STAGE_BYTES = PLANES * (A_TILE_BYTES + B_TILE_BYTES);     // 48 KiB / 96 KiB
  static constexpr int STAGES = NSPLIT == 3 ? 2 : 4;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/ + 8 * 32 * 32 * 4 /*epilogue staging*/;
};

// ---------------------------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Suspend-time hint: a waiting warp sleeps in hardware until the phase flips (or the hint expires) instead of spinning.
// Measured on the 96 GEMMs of sampler step 0: 51.0 ms -> 49.1 ms (the eight epilogue warps and the producer no longer burn
// issue slots and power next to the MMA issuer).
constexpr uint32_t kSuspendHintNs = 20000;
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(bar), "r"(parity), "r"(kSuspendHintNs) : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug traps (-> CUDA error on the host) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 0x3ff) == 0 && clock64() - t0 > 8000000000LL) {
      printf("selftok gemm_tc: mbarrier timeout (block %d thread %d bar %u parity %u)\n", blockIdx.x, threadIdx.x, bar, parity);
      __trap();
    }
  }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major SWIZZLE_128B shared-memory operand descriptor (cute::UMMA::SmemDescriptor layout):
//   [0,14) start >> 4 | [16,30) LBO >> 4 (=1, unused for swizzled K-major) | [32,46) SBO >> 4 (8 rows * 128 B = 1024)
//   [46,48) version = 1 (sm_100) | [61,64) layout = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// kind::f16 instruction descriptor (cute::UMMA::InstrDescriptor): D fp32 (bits 4-5 = 1), A/B bf16 (bits 7-9, 10-12 = 1),
// both K-major (bits 15, 16 = 0), N >> 3 at bit 17, M >> 4 at bit 24.
__host__ __device__ constexpr uint32_t make_idesc(int m, int n, int fp16) {
  const uint32_t fmt = fp16 ? 0u : 1u;                   // F16F32Format: F16 = 0, BF16 = 1
  return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// Epilogue of one 32-row x 128-column accumulator block (TMEM lane quarter x column half), executed by one warp;
// 8 epilogue warps cover the 128 x 256 tile.  tcgen05.ld hands lane i row i of a 32 x 32 chunk; the chunk is transposed
// through an XOR-swizzled shared-memory tile (conflict-free 128-bit writes and reads) so that every global access of
// the warp covers four full 128-byte row segments (8 lanes x float4 per row).  For the gated-residual epilogue all
// residual / gate loads of a chunk are issued BEFORE the TMEM read: one memory latency per chunk, not one per row.
// (A lane-per-row epilogue touches 32 cache lines per instruction with nothing in flight and made the LSU, not the
// tensor pipe, the limiter of the first version of this kernel.)
constexpr int EPI_STAGE_FLOATS = 32 * 32;
constexpr int EPI_WARPS = 8;
__device__ __forceinline__ float4 ldg4(const float* p) { return *reinterpret_cast<const float4*>(p); }
// two fp32 -> packed 16-bit pair (IEEE half, saturating, or bf16): one cvt per pair
__device__ __forceinline__ uint32_t pack2(float lo, float hi, bool fp16) { return pack2_sat16(lo, hi, fp16); }
// MODE / GELU are compile-time so that the per-element path carries no mode branches; the arithmetic of all 8 row passes
// of a chunk is issued unconditionally (independent chains -> ILP) and only the global stores are predicated.
template <int MODE, bool GELU>
__device__ __forceinline__ void epilogue_block(const Epilogue& e, uint32_t tmem_addr, float* stage, int lane,
                                               int64_t m_base, int64_t M, int n_first, int N) {
  const int rsub = lane >> 3, cq = lane & 7;            // pass k: row 4k + rsub of the chunk; this lane's 4 columns
  const bool per_row_gate = MODE == EPI_RESID && e.gate && e.gate_period > 1;
  const bool per_row_add = MODE == EPI_STORE && e.addtab;
  const bool fp16 = e.fp16 != 0;
  int orow_[8];                                          // output row (32-bit; the 64-bit offset is formed at the access)
  int mrow[8];
  bool rvalid[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int64_t m = m_base + 4 * k + rsub;
    rvalid[k] = m < M;
    const int mi = (int)m;
    const int orow = e.rpb_in > 0 ? (mi / e.rpb_in) * e.rpb_out + e.row_off + (mi % e.rpb_in) : mi;
    orow_[k] = orow;
    mrow[k] = per_row_gate ? mi % e.gate_period : (per_row_add ? mi % e.add_period : 0);
  }
#pragma unroll 1
  for (int c = 0; c < (BN / 2) / 32; ++c) {
    const int n0 = n_first + c * 32;
    if (n0 >= N) break;                                 // warp-uniform
    const int n = n0 + cq * 4;
    const bool col_ok = n < N;                          // N % 4 == 0: a float4 group is all in or all out
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 bias = (e.bias && col_ok) ? ldg4(e.bias + n) : zero4;
    float4 res[8], gt[8];
    if (MODE == EPI_RESID) {
      const float4 g0 = (e.gate && !per_row_gate && col_ok) ? ldg4(e.gate + n) : make_float4(1.f, 1.f, 1.f, 1.f);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const bool ok = col_ok && rvalid[k];
        res[k] = ok ? ldg4(e.resid + (int64_t)orow_[k] * e.ldo + n) : zero4;
        gt[k] = (per_row_gate && ok) ? ldg4(e.gate + (int64_t)mrow[k] * e.gate_ld + n) : g0;
      }
    }
    uint32_t r[32];
    tmem_ld32(tmem_addr + (uint32_t)(c * 32), r);
    tmem_ld_wait();
    __syncwarp();                                       // previous chunk fully read back
#pragma unroll
    for (int q = 0; q < 8; ++q)
      *reinterpret_cast<uint4*>(&stage[lane * 32 + ((q ^ (lane & 7)) << 2)]) = make_uint4(r[4 * q], r[4 * q + 1], r[4 * q + 2], r[4 * q + 3]);
    __syncwarp();
    float4 y[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int row = 4 * k + rsub;
      // (a timing-only build that skipped this shared-memory transpose was 4.6 % faster: the staging is not the bottleneck)
      const float4 v = *reinterpret_cast<const float4*>(&stage[row * 32 + ((cq ^ (row & 7)) << 2)]);
      y[k] = make_float4(v.x + bias.x, v.y + bias.y, v.z + bias.z, v.w + bias.w);
      if (GELU) y[k] = gelu_tanh_fast4(y[k]);
      if (MODE == EPI_RESID) {
        y[k].x = fmaf(gt[k].x, y[k].x, res[k].x); y[k].y = fmaf(gt[k].y, y[k].y, res[k].y);
        y[k].z = fmaf(gt[k].z, y[k].z, res[k].z); y[k].w = fmaf(gt[k].w, y[k].w, res[k].w);
      }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if (!(col_ok && rvalid[k])) continue;
      const int64_t o = (int64_t)orow_[k] * e.ldo + n;
      if (MODE == EPI_SPLIT) {
        *reinterpret_cast<uint2*>(e.out_hi + o) = make_uint2(pack2(y[k].x, y[k].y, fp16), pack2(y[k].z, y[k].w, fp16));
        if (e.out_lo) {                                  // bf16x3: residual planes
          const float lx = y[k].x - __bfloat162float(__float2bfloat16_rn(y[k].x)), ly = y[k].y - __bfloat162float(__float2bfloat16_rn(y[k].y));
          const float lz = y[k].z - __bfloat162float(__float2bfloat16_rn(y[k].z)), lw = y[k].w - __bfloat162float(__float2bfloat16_rn(y[k].w));
          *reinterpret_cast<uint2*>(e.out_lo + o) = make_uint2(pack2(lx, ly, false), pack2(lz, lw, false));
        }
      } else {
        if (per_row_add) {
          const float4 a = ldg4(e.addtab + (int64_t)mrow[k] * e.add_ld + n);
          y[k].x += a.x; y[k].y += a.y; y[k].z += a.z; y[k].w += a.w;
        }
        *reinterpret_cast<float4*>(e.out + o) = y[k];
      }
    }
  }
}

// Issued by the epilogue warps BEFORE they wait for the accumulator: pulls this warp's 32 x 128 residual block (and its
// per-row gate rows) into L2 while the MMAs of the tile are still running, so the gated-residual loads of the epilogue hit L2
// instead of paying the HBM latency once per chunk (long-scoreboard stalls were 37 % of the epilogue's samples).
__device__ __forceinline__ void epilogue_prefetch(const Epilogue& e, int lane, int64_t m_base, int64_t M, int n_first, int N) {
  if (e.mode != EPI_RESID) return;
  const int64_t m = m_base + lane;
  if (m >= M || n_first >= N) return;
  const int mi = (int)m;
  const int orow = e.rpb_in > 0 ? (mi / e.rpb_in) * e.rpb_out + e.row_off + (mi % e.rpb_in) : mi;
  const float* rp = e.resid + (int64_t)orow * e.ldo + n_first;
#pragma unroll
  for (int j = 0; j < 4; ++j)
    if (n_first + j * 32 < N) asm volatile("prefetch.global.L2 [%0];" ::"l"(rp + j * 32));
  if (e.gate && e.gate_period > 1) {
    const float* gp = e.gate + (int64_t)(mi % e.gate_period) * e.gate_ld + n_first;
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (n_first + j * 32 < N) asm volatile("prefetch.global.L2 [%0];" ::"l"(gp + j * 32));
  }
}

// mode / activation dispatch (uniform across the grid)
__device__ __forceinline__ void epilogue_dispatch(const Epilogue& e, uint32_t tmem_addr, float* stage, int lane, int64_t m_base,
                                                  int64_t M, int n_first, int N) {
  if (e.mode == EPI_RESID) epilogue_block<EPI_RESID, false>(e, tmem_addr, stage, lane, m_base, M, n_first, N);
  else if (e.mode == EPI_SPLIT) {
    if (e.act == ACT_GELU) epilogue_block<EPI_SPLIT, true>(e, tmem_addr, stage, lane, m_base, M, n_first, N);
    else epilogue_block<EPI_SPLIT, false>(e, tmem_addr, stage, lane, m_base, M, n_first, N);
  } else {
    if (e.act == ACT_GELU) epilogue_block<EPI_STORE, true>(e, tmem_addr, stage, lane, m_base, M, n_first, N);
    else epilogue_block<EPI_STORE, false>(e, tmem_addr, stage, lane, m_base, M, n_first, N);
  }
}

struct TcMaps {
  CUtensorMap a_hi, a_lo, b_hi, b_lo;
};

struct GemmParams {
  int64_t M;
  int N, K;
  int fp16;
  Epilogue ep;
  // implicit-GEMM 3x3 convolution (stride 1, zero padding 1) over NHWC activations: A row m = output pixel (b, y, x), K index =
  // tap * C + c.  conv_C == 0: plain GEMM.  The A tile of k-block kb (tap = kb / (C / 64), channels chunk = kb % (C / 64)) is the
  // 4-D TMA box {64 channels, bw pixels, bh rows, bn images} (bw bh bn = 128) shifted by the tap offset; out-of-image elements
  // are zero-filled by the TMA unit -- exactly the padding.
  // conv_stride == 2 (Downsample, sd3_impls.py:287-298: zero pad right / bottom, 3x3 stride 2): the planes hold the four
  // polyphase components of the input, [image * 4 + (py * 2 + px)][H_out][W_out][C] with phase(py, px)[y][x] = in[2y + py][2x + px];
  // tap (dy, dx) reads phase (dy & 1, dx & 1) shifted by (dy >> 1, dx >> 1) -- unit-stride boxes again, the zero fill past the last
  // row / column is the one-sided padding.  conv_H / conv_W are the OUTPUT dims.
  int conv_C = 0, conv_H = 0, conv_W = 0, conv_stride = 1;
  int raster_gm = 4;     // pair-rows per raster group of the SM-pair kernel (pair_coords)
};

__device__ __forceinline__ void tile_coords(int t, int m_tiles, int n_tiles, int& m_blk, int& n_blk) {
  constexpr int GM = 8;
  const int per_group = GM * n_tiles;
  const int group = t / per_group;
  const int first_m = group * GM;
  const int gm = min(GM, m_tiles - first_m);
  const int local = t - group * per_group;
  m_blk = first_m + local % gm;
  n_blk = local / gm;
}

// ---------------------------------------------------------------------------------------------- kernel
template <int NSPLIT>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
               const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo,
               const GemmParams p) {
  using C = Cfg<NSPLIT>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;           // SWIZZLE_128B tiles need 1024 B alignment
  const uint32_t bar_base = smem_base + C::STAGES * C::STAGE_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (C::STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * C::STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * C::STAGES + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * C::STAGES + 4);
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_tiles = (int)((p.M + BM - 1) / BM), n_tiles = (p.N + BN - 1) / BN;
  const int num_tiles = m_tiles * n_tiles;
  const int nk = (p.K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a_hi);
    tma_prefetch_desc(&map_b_hi);
    if (NSPLIT == 3) { tma_prefetch_desc(&map_a_lo); tma_prefetch_desc(&map_b_lo); }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < C::STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), EPI_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"((uint32_t)TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // =========================================================== TMA producer
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        int m_blk, n_blk;
        tile_coords(t, m_tiles, n_tiles, m_blk, n_blk);
        for (int kb = 0; kb < nk; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          const uint32_t sa = smem_base + stage * C::STAGE_BYTES;
          const uint32_t fb = full_bar(stage);
          mbar_expect_tx(fb, C::STAGE_BYTES);
          tma_load_2d(sa, &map_a_hi, fb, kb * BK, m_blk * BM);
          tma_load_2d(sa + C::PLANES * A_TILE_BYTES, &map_b_hi, fb, kb * BK, n_blk * BN);
          if (NSPLIT == 3) {
            tma_load_2d(sa + A_TILE_BYTES, &map_a_lo, fb, kb * BK, m_blk * BM);
            tma_load_2d(sa + 2 * A_TILE_BYTES + B_TILE_BYTES, &map_b_lo, fb, kb * BK, n_blk * BN);
          }
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // =========================================================== MMA issuer
    if (lane == 0) {
      const uint32_t idesc = make_idesc(BM, BN, p.fp16);
      int stage = 0; uint32_t phase = 0;
      int it = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(tempty_bar(acc), acc_phase ^ 1);               // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)acc * BN;
        for (int kb = 0; kb < nk; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa_hi = smem_base + stage * C::STAGE_BYTES;
          const uint32_t sb_hi = sa_hi + C::PLANES * A_TILE_BYTES;
          const uint32_t sa_lo = sa_hi + A_TILE_BYTES;
          const uint32_t sb_lo = sb_hi + B_TILE_BYTES;
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            const uint32_t koff = k * UMMA_K * 2;                // bytes inside the 128 B swizzle row
            const uint64_t da_hi = make_smem_desc(sa_hi + koff), db_hi = make_smem_desc(sb_hi + koff);
            tc_mma_f16(d_tmem, da_hi, db_hi, idesc, (kb > 0 || k > 0) ? 1u : 0u);
            if (NSPLIT == 3) {
              const uint64_t da_lo = make_smem_desc(sa_lo + koff), db_lo = make_smem_desc(sb_lo + koff);
              tc_mma_f16(d_tmem, da_hi, db_lo, idesc, 1u);
              tc_mma_f16(d_tmem, da_lo, db_hi, idesc, 1u);
            }
          }
          tc_commit(empty_bar(stage));                           // smem slot reusable once these MMAs retire
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
        tc_commit(tfull_bar(acc));                               // accumulator complete -> epilogue
      }
    }
  } else {
    // =========================================================== epilogue warps 2..9 (TMEM lane quarter = warp % 4, column half = (warp-2)/4)
    const int quarter = warp & 3, half = (warp - 2) >> 2;        // TMEM lane quarter = warp % 4; column half
    const Epilogue& e = p.ep;
    float* stage = reinterpret_cast<float*>(smem_raw + (bar_base + 256 - smem_u32(smem_raw))) + (warp - 2) * EPI_STAGE_FLOATS;
    int it = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
      int m_blk, n_blk;
      tile_coords(t, m_tiles, n_tiles, m_blk, n_blk);
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      epilogue_prefetch(e, lane, (int64_t)m_blk * BM + quarter * 32, p.M, n_blk * BN + half * (BN / 2), p.N);
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      epilogue_dispatch(e, tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * BN + half * (BN / 2)), stage, lane,
                     (int64_t)m_blk * BM + quarter * 32, p.M, n_blk * BN + half * (BN / 2), p.N);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));
    }
  }
  // ---- teardown
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------- 2-CTA kernel
// cta_group::2: a cluster of two CTAs (one SM pair) computes a 256 x 256 output tile with UMMA 256x256x16.  Each CTA
// stages its own 128 A rows and HALF of the W tile (128 of the 256 N rows); the tensor cores of the pair exchange the
// halves, so per SM the shared-memory traffic (TMA fill + operand reads) drops from 156-192 B/clk to 104-128 B/clk and the
// L2 -> SM bytes per FLOP halve.  Only the leader CTA (cluster rank 0) issues MMAs.
//   full[s]    lives in the leader; both CTAs' TMA loads complete_tx on it (peer bit of the barrier address cleared)
//   empty[s]   one per CTA; the leader's tcgen05.commit multicasts the arrive to both
//   tfull[a]   one per CTA (multicast commit); tempty[a] lives in the leader and counts the 8 epilogue warps of the pair
template <int NSPLIT> struct Cfg2 {
  static constexpr int PLANES = NSPLIT == 3 ? 2 : 1;
  static constexpr int HALF_B_BYTES = (BN / 2) * BK * 2;                                 // 16 KiB
  static constexpr int STAGE_BYTES = PLANES * (A_TILE_BYTES + HALF_B_BYTES);             // 32 KiB / 64 KiB per CTA
  static constexpr int STAGES = NSPLIT == 3 ? 3 : 6;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256 + 8 * 32 * 32 * 4;
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const CUtensorMap* map, uint32_t leader_bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(leader_bar), "r"(c0), "r"(c1) : "memory");
}
// The weight tile is re-read by every group of M tiles; without a hint the activation / residual / output streams of the
// K = 6144 GEMM push it out of L2 between groups (fc2: 1.69 GB of DRAM reads for 0.93 GB algorithmic).  evict_last keeps it.
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void tma_load_4d_2sm(uint32_t dst, const CUtensorMap* map, uint32_t leader_bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(leader_bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm_hint(uint32_t dst, const CUtensorMap* map, uint32_t leader_bar, int c0, int c1,
                                                     uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(leader_bar), "r"(c0), "r"(c1), "l"(policy) : "memory");
}
__device__ __forceinline__ void tc_commit_mc2(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void tc_mma_f16_2sm(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// arrive on the barrier at the same offset in CTA `rank` of the cluster.  Default semantics (.release at CTA scope), as
// CUTLASS's ClusterBarrier::arrive: the barrier only hands the TMEM accumulator back to the MMA issuer, and the TMEM reads
// are already complete (tcgen05.wait::ld) and ordered (tcgen05.fence::before_thread_sync).  The explicit .release.cluster
// form used in round 1 compiled to MEMBAR.ALL.GPU + ERRBAR + CGAERRBAR in front of the arrive -- it drained the warp's
// outstanding global stores first and held 6-9 % of all warp samples in every GEMM (profiles/r1_layer_ncu_full_final.md).
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t local_bar, uint32_t rank) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}"
      ::"r"(local_bar), "r"(rank) : "memory");
}

__device__ __forceinline__ void pair_coords(int t, int pm_tiles, int n_tiles, int GM, int& pm, int& n_blk) {
  // GM pair-rows (default 4 = 1024 A rows) share each W tile in L2; SELFTOK_GEMM_GM is a measurement knob (any value is a
  // bijection of the tile list, results never change)
  const int per_group = GM * n_tiles;
  const int group = t / per_group;
  const int first = group * GM;
  const int gm = min(GM, pm_tiles - first);
  const int local = t - group * per_group;
  pm = first + local % gm;
  n_blk = local / gm;
}

template <int NSPLIT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
gemm_tc2_kernel(const __grid_constant__ TcMaps maps0, const __grid_constant__ TcMaps maps1, const GemmParams p0,
                const GemmParams p1) {
  // Up to two independent problems (same NSPLIT / operand type) share the launch: the context- and the image-stream GEMM of
  // a layer.  Tiles of problem 0 come first; problem 1 may be empty (M == 0).
  using C = Cfg2<NSPLIT>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + C::STAGES * C::STAGE_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (C::STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * C::STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * C::STAGES + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * C::STAGES + 4);
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int pm_tiles0 = (int)((p0.M + 2 * BM - 1) / (2 * BM)), n_tiles0 = (p0.N + BN - 1) / BN;
  const int pm_tiles1 = (int)((p1.M + 2 * BM - 1) / (2 * BM)), n_tiles1 = (p1.N + BN - 1) / BN;
  const int tiles0 = pm_tiles0 * n_tiles0;
  const int num_tiles = tiles0 + pm_tiles1 * n_tiles1;
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&maps0.a_hi);
    tma_prefetch_desc(&maps0.b_hi);
    if (NSPLIT == 3) { tma_prefetch_desc(&maps0.a_lo); tma_prefetch_desc(&maps0.b_lo); }
    if (p1.M > 0) {
      tma_prefetch_desc(&maps1.a_hi);
      tma_prefetch_desc(&maps1.b_hi);
      if (NSPLIT == 3) { tma_prefetch_desc(&maps1.a_lo); tma_prefetch_desc(&maps1.b_lo); }
    }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < C::STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 2 * EPI_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"((uint32_t)TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                     // peer barriers initialised before any remote arrive / multicast
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // =========================================================== TMA producer (both CTAs)
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      const uint64_t w_policy = l2_policy_evict_last();
      for (int t = cluster_id; t < num_tiles; t += num_clusters) {
        const bool second = t >= tiles0;
        const TcMaps& mp = second ? maps1 : maps0;
        int pm, n_blk;
        if (second) pair_coords(t - tiles0, pm_tiles1, n_tiles1, p1.raster_gm, pm, n_blk);
        else pair_coords(t, pm_tiles0, n_tiles0, p0.raster_gm, pm, n_blk);
        const GemmParams& pp = second ? p1 : p0;
        const int nk = (pp.K + BK - 1) / BK;
        const int m_row = pm * 2 * BM + (int)rank * BM;            // this CTA's 128 A rows
        const int n_row = n_blk * BN + (int)rank * (BN / 2);       // this CTA's half of the W tile
        // convolution: the 128 rows are 128 / bw image rows of bw pixels starting at (cb, cy, cx)
        const int chunks = pp.conv_C > 0 ? pp.conv_C / BK : 1;
        int cx = 0, cy = 0, cb = 0;
        if (pp.conv_C > 0) {
          cx = m_row % pp.conv_W;
          cy = (m_row / pp.conv_W) % pp.conv_H;
          cb = m_row / (pp.conv_W * pp.conv_H);
        }
        for (int kb = 0; kb < nk; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          const uint32_t sa = smem_base + stage * C::STAGE_BYTES;
          const uint32_t fb = full_bar(stage) & 0xFEFFFFFFu;       // leader's barrier (peer bit cleared)
          if (leader) mbar_expect_tx(full_bar(stage), 2 * C::STAGE_BYTES);
          if (pp.conv_C > 0) {
            const int tap = kb / chunks, c0 = (kb - tap * chunks) * BK;
            const int dy = tap / 3, dx = tap - dy * 3;
            int x0 = cx + dx - 1, y0 = cy + dy - 1, n0 = cb;
            if (pp.conv_stride == 2) { x0 = cx + (dx >> 1); y0 = cy + (dy >> 1); n0 = cb * 4 + (dy & 1) * 2 + (dx & 1); }
            tma_load_4d_2sm(sa, &mp.a_hi, fb, c0, x0, y0, n0);
            if (NSPLIT == 3) tma_load_4d_2sm(sa + A_TILE_BYTES, &mp.a_lo, fb, c0, x0, y0, n0);
          } else {
            tma_load_2d_2sm(sa, &mp.a_hi, fb, kb * BK, m_row);
            if (NSPLIT == 3) tma_load_2d_2sm(sa + A_TILE_BYTES, &mp.a_lo, fb, kb * BK, m_row);
          }
          tma_load_2d_2sm_hint(sa + C::PLANES * A_TILE_BYTES, &mp.b_hi, fb, kb * BK, n_row, w_policy);
          if (NSPLIT == 3) tma_load_2d_2sm_hint(sa + 2 * A_TILE_BYTES + C::HALF_B_BYTES, &mp.b_lo, fb, kb * BK, n_row, w_policy);
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // =========================================================== MMA issuer (leader CTA only)
    if (leader && lane == 0) {
      const uint32_t idesc = make_idesc(2 * BM, BN, p0.fp16);
      int stage = 0; uint32_t phase = 0;
      int it = 0;
      for (int t = cluster_id; t < num_tiles; t += num_clusters, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        const int nk = ((t >= tiles0 ? p1.K : p0.K) + BK - 1) / BK;
        mbar_wait(tempty_bar(acc), acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)acc * BN;
        for (int kb = 0; kb < nk; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa_hi = smem_base + stage * C::STAGE_BYTES;
          const uint32_t sb_hi = sa_hi + C::PLANES * A_TILE_BYTES;
          const uint32_t sa_lo = sa_hi + A_TILE_BYTES;
          const uint32_t sb_lo = sb_hi + C::HALF_B_BYTES;
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            const uint32_t koff = k * UMMA_K * 2;
            const uint64_t da_hi = make_smem_desc(sa_hi + koff), db_hi = make_smem_desc(sb_hi + koff);
            tc_mma_f16_2sm(d_tmem, da_hi, db_hi, idesc, (kb > 0 || k > 0) ? 1u : 0u);
            if (NSPLIT == 3) {
              const uint64_t da_lo = make_smem_desc(sa_lo + koff), db_lo = make_smem_desc(sb_lo + koff);
              tc_mma_f16_2sm(d_tmem, da_hi, db_lo, idesc, 1u);
              tc_mma_f16_2sm(d_tmem, da_lo, db_hi, idesc, 1u);
            }
          }
          tc_commit_mc2(empty_bar(stage));                         // frees the slot in BOTH CTAs
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
        tc_commit_mc2(tfull_bar(acc));                             // accumulator ready in BOTH CTAs
      }
    }
  } else {
    // =========================================================== epilogue warps 2..9 of both CTAs
    const int quarter = warp & 3, half = (warp - 2) >> 2;        // TMEM lane quarter = warp % 4; column half
    float* stage = reinterpret_cast<float*>(smem_raw + (bar_base + 256 - smem_u32(smem_raw))) + (warp - 2) * EPI_STAGE_FLOATS;
    int it = 0;
    for (int t = cluster_id; t < num_tiles; t += num_clusters, ++it) {
      const bool second = t >= tiles0;
      int pm, n_blk;
      if (second) pair_coords(t - tiles0, pm_tiles1, n_tiles1, p1.raster_gm, pm, n_blk);
      else pair_coords(t, pm_tiles0, n_tiles0, p0.raster_gm, pm, n_blk);
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int64_t m_base = (int64_t)pm * 2 * BM + (int64_t)rank * BM + quarter * 32;
      const int n_first = n_blk * BN + half * (BN / 2);
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * BN + half * (BN / 2));
      // the two problems are handled by separate (statically addressed) copies of the epilogue: selecting the parameter block
      // dynamically costs registers in the hottest loop of the kernel
      if (!second) {
        epilogue_prefetch(p0.ep, lane, m_base, p0.M, n_first, p0.N);
        mbar_wait(tfull_bar(acc), acc_phase);
        tc_fence_after();
        epilogue_dispatch(p0.ep, taddr, stage, lane, m_base, p0.M, n_first, p0.N);
      } else {
        epilogue_prefetch(p1.ep, lane, m_base, p1.M, n_first, p1.N);
        mbar_wait(tfull_bar(acc), acc_phase);
        tc_fence_after();
        epilogue_dispatch(p1.ep, taddr, stage, lane, m_base, p1.M, n_first, p1.N);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(tempty_bar(acc), 0);       // the leader's barrier counts all 8 epilogue warps
    }
  }
  // ---- teardown: nobody may exit (or free TMEM) while the peer can still touch its barriers / shared memory
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;
// per-device state: cudaFuncSetAttribute and the SM count belong to a device, and one process may hold a handle per GPU
constexpr int kMaxDev = 64;
int g_num_sms_dev[kMaxDev];
bool g_attr_dev[kMaxDev];
int g_raster_gm = 4;      // SELFTOK_GEMM_GM (measurement knob)
int g_gemm_ctas = 2;      // 2: cta_group::2 pair kernel (default); 1: single-CTA kernel (SELFTOK_GEMM_CTAS=1)

int make_map(CUtensorMap* map, const __nv_bfloat16* ptr, int64_t rows, int K, int box_rows, int fp16) {
  cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)K * 2};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(map, fp16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<__nv_bfloat16*>(ptr), gdim, gstride, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r));
    return -5;
  }
  return 0;
}

// 4-D map over NHWC 16-bit activations [B][H][W][C] with box {64 channels, bw pixels, bh rows, 128 / (bw bh) images}: 128 output
// pixels per CTA = part of an image row (W >= 128), whole rows of one image, or -- for images smaller than 128 pixels -- whole images
int make_map_nhwc(CUtensorMap* map, const __nv_bfloat16* ptr, int64_t B, int H, int W, int Cc, int fp16) {
  const int bw = W < BM ? W : BM;
  const int bh = (BM / bw) < H ? (BM / bw) : H;
  const int bn = BM / (bw * bh);
  cuuint64_t gdim[4] = {(cuuint64_t)Cc, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t gstride[3] = {(cuuint64_t)Cc * 2, (cuuint64_t)W * Cc * 2, (cuuint64_t)H * W * Cc * 2};
  cuuint32_t box[4] = {(cuuint32_t)BK, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bn};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = g_encode(map, fp16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<__nv_bfloat16*>(ptr), gdim, gstride,
                        box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (NHWC) failed with CUresult " + std::to_string((int)r));
    return -5;
  }
  return 0;
}

}  // namespace

// Generic 2-D SWIZZLE_128B tensor map over a row-major 16-bit matrix [rows, cols] (box_cols * 2 bytes must be 128).
int make_tensor_map_2d(CUtensorMap* map, const void* ptr, uint64_t rows, uint64_t cols, uint32_t box_rows, uint32_t box_cols,
                       int fp16) {
  STK_CHECK(g_encode, -3, "gemm_tc_init has not been called");
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)cols * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(map, fp16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim,
                        gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r));
    return -5;
  }
  return 0;
}

void gemm_tc_set_ctas(int n) { g_gemm_ctas = n == 1 ? 1 : 2; }

// Idempotent per device (the current one): resolves cuTensorMapEncodeTiled once per process, opts the kernels into their
// dynamic shared memory and records the SM count once per device.
int gemm_tc_init() {
  int dev = 0;
  STK_CUDA(cudaGetDevice(&dev));
  STK_CHECK(dev >= 0 && dev < kMaxDev, -1, "gemm_tc: device ordinal out of range");
  if (g_encode && g_attr_dev[dev]) return 0;
  if (!g_encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    STK_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    STK_CHECK(fn && qres == cudaDriverEntryPointSuccess, -5, "cuTensorMapEncodeTiled not available from the driver");
    const char* v = getenv("SELFTOK_GEMM_CTAS");
    if (v) g_gemm_ctas = atoi(v) == 1 ? 1 : 2;
    const char* gm = getenv("SELFTOK_GEMM_GM");
    if (gm && atoi(gm) >= 1 && atoi(gm) <= 1024) g_raster_gm = atoi(gm);
    g_encode = reinterpret_cast<EncodeTiledFn>(fn);
  }
  STK_CUDA(cudaDeviceGetAttribute(&g_num_sms_dev[dev], cudaDevAttrMultiProcessorCount, dev));
  STK_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<1>::SMEM_BYTES));
  STK_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<3>::SMEM_BYTES));
  STK_CUDA(cudaFuncSetAttribute(gemm_tc2_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg2<1>::SMEM_BYTES));
  STK_CUDA(cudaFuncSetAttribute(gemm_tc2_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg2<3>::SMEM_BYTES));
  g_attr_dev[dev] = true;
  return 0;
}

static int check_problem(const TcProblem& q, int nsplit, int fp16) {
  const Epilogue& ep = q.ep;
  STK_CHECK(q.A_hi && q.W_hi && q.M > 0 && q.N > 0 && q.K > 0, -1, "gemm_tc: bad arguments");
  STK_CHECK(nsplit == 1 || (nsplit == 3 && q.A_lo && q.W_lo), -1, "gemm_tc: nsplit must be 1, or 3 with lo planes");
  STK_CHECK(!fp16 || nsplit == 1, -1, "gemm_tc: the fp16 mode is single-pass");
  STK_CHECK(ep.act == ACT_NONE || (ep.act == ACT_GELU && ep.mode != EPI_RESID), -2,
            "gemm_tc: epilogue activation must be none, or GELU-tanh with the store / split modes");
  STK_CHECK(ep.mode != EPI_STORE || ep.addtab == nullptr || ep.add_ld % 4 == 0, -2, "gemm_tc: addtab pitch must be a multiple of 4");
  STK_CHECK(q.K % 8 == 0, -2, "gemm_tc: K must be a multiple of 8 (16-byte TMA row pitch)");
  STK_CHECK(ep.ldo % 4 == 0 && q.N % 4 == 0, -2, "gemm_tc: N and the output pitch must be multiples of 4");
  STK_CHECK(ep.mode != EPI_RESID || ep.gate == nullptr || ep.gate_ld % 4 == 0, -2, "gemm_tc: gate pitch must be a multiple of 4");
  if (q.conv_C > 0) {                                      // implicit 3x3 convolution over NHWC planes
    STK_CHECK(q.conv_C % BK == 0 && q.K == 9 * q.conv_C, -2, "gemm_tc conv: channels must be a multiple of 64 and K = 9 C");
    STK_CHECK(q.conv_W > 0 && q.conv_H > 0 && (q.conv_W % BM == 0 || BM % q.conv_W == 0), -2, "gemm_tc conv: image width must divide or be a multiple of 128");
    const int64_t hw = (int64_t)q.conv_H * q.conv_W;
    STK_CHECK(q.M % hw == 0 && (hw % BM == 0 || BM % hw == 0) && (q.conv_W >= BM || hw < BM || q.conv_H % (BM / q.conv_W) == 0), -2,
              "gemm_tc conv: 128-pixel tiles must be part of a row, whole rows of one image, or whole images");
    STK_CHECK(q.conv_stride == 1 || (q.conv_stride == 2 && hw % BM == 0), -2, "gemm_tc conv: stride 2 needs >= 128 output pixels per image");
  }
  return 0;
}

static int make_maps(TcMaps* m, const TcProblem& q, int nsplit, int fp16, int b_box) {
  const bool conv = q.conv_C > 0;
  const int64_t imgs = conv ? q.M / ((int64_t)q.conv_H * q.conv_W) * (q.conv_stride == 2 ? 4 : 1) : 0;   // stride 2: 4 phase planes per image
  if (conv) STK_TRY(make_map_nhwc(&m->a_hi, q.A_hi, imgs, q.conv_H, q.conv_W, q.conv_C, fp16));
  else STK_TRY(make_map(&m->a_hi, q.A_hi, q.M, q.K, BM, fp16));
  STK_TRY(make_map(&m->b_hi, q.W_hi, q.N, q.K, b_box, fp16));
  if (nsplit == 3) {
    if (conv) STK_TRY(make_map_nhwc(&m->a_lo, q.A_lo, imgs, q.conv_H, q.conv_W, q.conv_C, fp16));
    else STK_TRY(make_map(&m->a_lo, q.A_lo, q.M, q.K, BM, fp16));
    STK_TRY(make_map(&m->b_lo, q.W_lo, q.N, q.K, b_box, fp16));
  } else {
    m->a_lo = m->a_hi; m->b_lo = m->b_hi;
  }
  return 0;
}

// One launch for up to two independent problems (the context- and the image-stream GEMM of an MMDiT layer): their tiles share
// the persistent grid, so the small N = 1536 GEMMs no longer pay a partially filled last wave each.
int launch_gemm_tc_grouped(const TcProblem* probs, int n, int nsplit, cudaStream_t s, int fp16) {
  STK_TRY(gemm_tc_init());                                // no-op after the first call on this device
  STK_CHECK(probs && (n == 1 || n == 2), -1, "gemm_tc: one or two problems per launch");
  int dev = 0;
  STK_CUDA(cudaGetDevice(&dev));
  const int g_num_sms = g_num_sms_dev[dev];
  for (int i = 0; i < n; ++i) STK_TRY(check_problem(probs[i], nsplit, fp16));
  const bool pair = g_gemm_ctas == 2 && g_num_sms >= 2;
  for (int i = 0; i < n; ++i) STK_CHECK(pair || probs[i].conv_C == 0, -2, "gemm_tc conv: only the SM-pair kernel stages NHWC tiles");
  if (!pair) {                                            // single-CTA bisecting kernel: one launch per problem
    for (int i = 0; i < n; ++i) {
      const TcProblem& q = probs[i];
      TcMaps m;
      STK_TRY(make_maps(&m, q, nsplit, fp16, BN));
      GemmParams p{q.M, q.N, q.K, fp16, q.ep};
      const int tiles = (int)((q.M + BM - 1) / BM) * ((q.N + BN - 1) / BN);
      const int grid = tiles < g_num_sms ? tiles : g_num_sms;
      if (nsplit == 3) gemm_tc_kernel<3><<<grid, NUM_THREADS, Cfg<3>::SMEM_BYTES, s>>>(m.a_hi, m.a_lo, m.b_hi, m.b_lo, p);
      else gemm_tc_kernel<1><<<grid, NUM_THREADS, Cfg<1>::SMEM_BYTES, s>>>(m.a_hi, m.a_lo, m.b_hi, m.b_lo, p);
      count_launch();
      STK_CUDA(cudaGetLastError());
    }
    return 0;
  }
  TcMaps m0, m1;
  STK_TRY(make_maps(&m0, probs[0], nsplit, fp16, BN / 2));
  auto mk_params = [&](const TcProblem& q) {
    GemmParams g{q.M, q.N, q.K, fp16, q.ep};
    g.conv_C = q.conv_C; g.conv_H = q.conv_H; g.conv_W = q.conv_W; g.conv_stride = q.conv_stride;
    g.raster_gm = g_raster_gm;
    return g;
  };
  GemmParams p0 = mk_params(probs[0]);
  GemmParams p1 = p0;
  p1.M = 0;
  m1 = m0;
  int pairs = (int)((probs[0].M + 2 * BM - 1) / (2 * BM)) * ((probs[0].N + BN - 1) / BN);
  if (n == 2) {
    STK_TRY(make_maps(&m1, probs[1], nsplit, fp16, BN / 2));
    p1 = mk_params(probs[1]);
    pairs += (int)((probs[1].M + 2 * BM - 1) / (2 * BM)) * ((probs[1].N + BN - 1) / BN);
  }
  const int clusters = pairs < g_num_sms / 2 ? pairs : g_num_sms / 2;
  if (nsplit == 3) gemm_tc2_kernel<3><<<2 * clusters, NUM_THREADS, Cfg2<3>::SMEM_BYTES, s>>>(m0, m1, p0, p1);
  else gemm_tc2_kernel<1><<<2 * clusters, NUM_THREADS, Cfg2<1>::SMEM_BYTES, s>>>(m0, m1, p0, p1);
  count_launch();
  STK_CUDA(cudaGetLastError());
  return 0;
}

int launch_gemm_tc(const __nv_bfloat16* A_hi, const __nv_bfloat16* A_lo, const __nv_bfloat16* W_hi,
                   const __nv_bfloat16* W_lo, int64_t M, int N, int K, int nsplit, const Epilogue& ep,
                   cudaStream_t s, int fp16) {
  TcProblem q{A_hi, A_lo, W_hi, W_lo, M, N, K, ep};
  return launch_gemm_tc_grouped(&q, 1, nsplit, s, fp16);
}

}  // namespace stk
