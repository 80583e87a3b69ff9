// Joint attention of the MMDiT on tcgen05 tensor cores with TMEM accumulators (sm_100a), head_dim 64, single-pass 16-bit
// operands (IEEE half or bf16), fp32 softmax.  FOUR INDEPENDENT QUERY-TILE STREAMS PER SM.
//
// Why (measured on B200, profiles/r2_attention_trace.md): the previous kernel (attn_tc5: one 128-row query tile per CTA, S
// double-buffered, two CTAs per SM) was bound by a latency chain, not by a pipe: per tile softmax(g) -> p_ready -> P V (g) ->
// Q K^T (g+2) -> softmax(g+2) took ~1900 + ~1650 cycles for two tiles while MUFU sat at 56 % and the tensor pipe at 26 %
// (removing ALL exponentials only bought 23 %; a single-thread tcgen05.mma issue of a 128 x 64 x 16 instruction blocks for
// 46 - 92 cycles, profiles/mma_cadence.cu).  Look-ahead inside a stream cannot hide that chain; other streams can.
//
//   CTA            two query tiles X, Y (2 x 128 consecutive rows of one (image, head)) = two streams that share every K / V
//                  tile (loaded once: half the TMA / shared-memory traffic of two independent CTAs); 2 CTAs per SM -> 4 streams.
//   warp 0         TMA producer: Q_X, Q_Y of the NEXT item (double-buffered), K / V tiles (64 keys) through a 2-stage ring
//   warp 1 / 2     MMA issuer of stream X / Y (one thread each):  S = Q K^T -> TMEM [0,64) of the stream's 128 columns,
//                  O += P V -> TMEM [64,128); per tile: wait P, issue P V (g), then Q K^T (g+1) straight behind it
//   warps 3-6      softmax of stream X, ONE THREAD PER QUERY ROW (TMEM lane = row): no cross-thread exchange, no named barrier,
//   warps 7-10     softmax of stream Y.   Per tile a thread makes two passes over its 64 scores in TMEM: row maximum (lazy
//                  rescale: the reference maximum only moves on jumps > 2^8), then the exponentials one 32-column half at a
//                  time; P (16-bit pairs) overwrites S columns [0,32) in place as each half is done and feeds tcgen05.mma as
//                  a TMEM A operand.
//   registers      11 warps x 2 CTAs -> 88 registers per thread (a setmaxnreg split did not make ptxas use more in the softmax
//                  branch); the 64 scores of a row therefore pass through the registers in two 32-column halves
//   S is single-buffered per stream on purpose: TMEM (512 columns) and the register file (64 K) are spent on four streams with
//   a short chain each instead of two streams with look-ahead.
//
// Same contract as the kernel it replaces (sd3/mmdit.py:521-531, sd3/other_impls.py:37-45): dense non-causal attention over
// the joint [context prefix ; image] sequence; rows < ctx_rows only see keys < ctx_keys (renderer rule, mmdit.py:1581).
#include "common.cuh"
#include "kernels.h"

#include <cuda.h>

#include <algorithm>
#include <stdlib.h>

namespace stk {

// provided by gemm_tc.cu
int make_tensor_map_2d(CUtensorMap* map, const void* ptr, uint64_t rows, uint64_t cols, uint32_t box_rows, uint32_t box_cols,
                       int fp16);

namespace {

constexpr int HD = 64, BQ = 128, BKV = 64, KV_STAGES = 2;
constexpr int Q_BYTES = BQ * HD * 2, KV_TILE_BYTES = BKV * HD * 2;
constexpr int SMEM_TILES = 4 * Q_BYTES + KV_STAGES * 2 * KV_TILE_BYTES;            // Q: 2 streams x 2 buffers = 64 KiB; K/V 32 KiB
constexpr int SMEM_BYTES = SMEM_TILES + 1024 + 256;                                // + alignment slack + barriers
constexpr int TMEM_COLS = 256;      // stream s: S [128 s, 128 s + 64) (P in place of its upper half) | O [128 s + 64, 128 s + 128)
constexpr int NUM_THREADS = 11 * 32;            // warps 0-2: TMA, MMA_X, MMA_Y | 3-6 softmax X | 7-10 softmax Y (88 registers each)
constexpr float kRescaleThreshold = 8.0f;       // log2 units

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
      printf("selftok attn_tc6: mbarrier timeout (block %d thread %d bar %u parity %u)\n", blockIdx.x, threadIdx.x, bar, parity);
      __trap();
    }
  }
}
// MMA-issuer flavour: plain polling (no suspend hint) -- experiment switch SELFTOK_ATTN6_MMA_SPIN
__device__ __forceinline__ void mbar_wait_spin(uint32_t bar, uint32_t parity) {
#ifdef SELFTOK_ATTN6_MMA_SPIN
  uint32_t ok = 0;
  const long long t0 = clock64();
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (!ok && clock64() - t0 > 8000000000LL) __trap();
  }
#else
  mbar_wait(bar, parity);
#endif
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1) : "memory");
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
// A operand from tensor memory (P of the P V product), B from shared memory
__device__ __forceinline__ void tc_mma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
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
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
        "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
        "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// SWIZZLE_128B shared-memory descriptors (cute::UMMA::SmemDescriptor): 8 rows x 128 B atoms, SBO = 1024 B between atoms.
// The same encoding serves the K-major operands (Q, K: 64 K-elements per 128 B row) and the MN-major V tile (64 head dims
// contiguous per key row, 8 keys per atom); the major-ness is selected in the instruction descriptor.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// kind::f16 instruction descriptor: D fp32, A/B format (0 = F16, 1 = BF16), a_major bit 15, b_major bit 16 (1 = MN-major)
__device__ __forceinline__ uint32_t make_idesc(int m, int n, int fp16, int b_mn_major) {
  const uint32_t fmt = fp16 ? 0u : 1u;
  return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)b_mn_major << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// packed fp32 pipe: (x0, x1) = (s0, s1) * scale + nsub;  (a0, a1) += (e0, e1)
__device__ __forceinline__ void scale_sub2(uint32_t s0, uint32_t s1, float scale, float nsub, float& x0, float& x1) {
  asm("{\n\t.reg .b64 x, sc, sb;\n\t"
      "mov.b64 x, {%2, %3};\n\t"
      "mov.b64 sc, {%4, %4};\n\t"
      "mov.b64 sb, {%5, %5};\n\t"
      "fma.rn.f32x2 x, x, sc, sb;\n\t"
      "mov.b64 {%0, %1}, x;\n\t}"
      : "=f"(x0), "=f"(x1) : "r"(s0), "r"(s1), "f"(scale), "f"(nsub));
}
__device__ __forceinline__ void add2(float& a0, float& a1, float e0, float e1) {
  asm("{\n\t.reg .b64 a, b;\n\t"
      "mov.b64 a, {%0, %1};\n\t"
      "mov.b64 b, {%2, %3};\n\t"
      "add.f32x2 a, a, b;\n\t"
      "mov.b64 {%0, %1}, a;\n\t}"
      : "+f"(a0), "+f"(a1) : "f"(e0), "f"(e1));
}
// two fp32 -> packed 16-bit pair (IEEE half or bf16), one cvt instruction
__device__ __forceinline__ uint32_t pack2_16(float lo, float hi, bool fp16) {
  uint32_t r;
  if (fp16) asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  else asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

#ifdef SELFTOK_ATTN_TRACE
__device__ unsigned long long g_trace6[16][4096];
#define TRACE_DECL int trace_i = 0
#define TRACE(tag, g)                                                                                                   \
  do {                                                                                                                  \
    if (lane == 0 && blockIdx.x == 0 && trace_i < 4096)                                                                 \
      g_trace6[warp][trace_i++] = ((unsigned long long)(tag) << 56) | ((unsigned long long)(warp & 0xff) << 48) |       \
                                  ((unsigned long long)((g) & 0xffff) << 32) | (unsigned long long)(clock64() & 0xffffffffu); \
  } while (0)
#else
#define TRACE_DECL
#define TRACE(tag, g) do { } while (0)
#endif

struct Attn6Params {
  AttnOut out;
  int B, S, H, ctx_rows, ctx_keys, fp16;
  float scale_log2e;
};

template <bool FP16>
__global__ void __maxnreg__(80)       // 2 CTAs x 11 (allocated as 12) warps x 80 registers: 88 leaves room for one CTA only (measured)
attention_tc6_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_kv, const Attn6Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t q_s = base;                                  // Q of stream s, buffer qb at q_s + (2 s + qb) * Q_BYTES
  const uint32_t kv_s = base + 4 * Q_BYTES;                   // stage st: K at kv_s + st*2*KV_TILE_BYTES, V right after
  const uint32_t bars = kv_s + KV_STAGES * 2 * KV_TILE_BYTES;
  auto q_full = [&](int s, int qb) { return bars + 8u * (2 * s + qb); };            // TMA -> MMA_s
  auto q_empty = [&](int s, int qb) { return bars + 32 + 8u * (2 * s + qb); };      // MMA_s (commit) -> TMA
  auto s_full = [&](int s) { return bars + 64 + 8u * s; };                          // MMA_s (commit) -> softmax_s
  auto p_ready = [&](int s) { return bars + 80 + 8u * s; };                         // softmax_s (4 warps) -> MMA_s
  auto pv_done = [&](int s) { return bars + 96 + 8u * s; };                         // MMA_s (commit) -> softmax_s (item end, rescale)
  auto kv_full = [&](int st) { return bars + 112 + 8u * st; };                      // TMA -> both MMA warps
  auto kv_empty = [&](int st) { return bars + 112 + 8u * KV_STAGES + 8u * st; };    // both MMA warps -> TMA (count 2)
  const uint32_t tmem_slot = bars + 112 + 16u * KV_STAGES;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  TRACE_DECL;
  const int S = p.S;
  // Work item = (image b, head h, PAIR of consecutive query tiles): stream X takes tile 2 pr, stream Y tile 2 pr + 1 (absent
  // for the last pair of an odd tile count).  Pair index fastest, so that the CTAs running side by side share K / V in L2.
  const int nq = (S + BQ - 1) / BQ, npair = (nq + 1) / 2;
  const int n_items = npair * p.H * p.B;
  const int G = (int)gridDim.x;
  // key tiles a query tile needs: every row of a pure context tile only sees the context keys (renderer rule)
  auto tile_keys = [&](int qt) { return ((qt + 1) * BQ <= p.ctx_rows) ? p.ctx_keys : S; };
  auto tiles_of = [&](int qt) { return qt < nq ? (tile_keys(qt) + BKV - 1) / BKV : 0; };
  // the pair walks max(tiles_X, tiles_Y) K / V tiles; a stream that needs fewer only hands the surplus stages back
  auto pair_tiles = [&](int pr) { return max(tiles_of(2 * pr), tiles_of(2 * pr + 1)); };

  if (warp == 1 && lane == 0) {
    for (int s = 0; s < 2; ++s) {
      for (int qb = 0; qb < 2; ++qb) { mbar_init(q_full(s, qb), 1); mbar_init(q_empty(s, qb), 1); }
      mbar_init(s_full(s), 1); mbar_init(p_ready(s), 4); mbar_init(pv_done(s), 1);
    }
    for (int st = 0; st < KV_STAGES; ++st) { mbar_init(kv_full(st), 1); mbar_init(kv_empty(st), 2); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"((uint32_t)TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp < 3) {
    if (warp == 0) {
      // =========================================================== TMA producer
      if (lane == 0) {
        // Q buffers are double-buffered PER STREAM by the stream's own use counter (stream Y is absent in the last pair of an
        // odd tile count, so the two streams do not use their buffers in lock step)
        int uq[2] = {0, 0};
        auto load_q = [&](int item) {
          const int pr = item % npair, h = (item / npair) % p.H, b = item / (npair * p.H);
          for (int s = 0; s < 2; ++s) {
            const int qt = 2 * pr + s;
            if (qt >= nq) continue;
            const int qb = uq[s] & 1;
            mbar_wait(q_empty(s, qb), ((uq[s] >> 1) & 1) ^ 1);
            mbar_expect_tx(q_full(s, qb), Q_BYTES);
            tma_load_2d(q_s + (2 * s + qb) * Q_BYTES, &map_q, q_full(s, qb), h * HD, b * S + qt * BQ);
            ++uq[s];
          }
        };
        int g = 0;
        if ((int)blockIdx.x < n_items) load_q(blockIdx.x);
        for (int item = blockIdx.x; item < n_items; item += G) {
          const int pr = item % npair, h = (item / npair) % p.H, b = item / (npair * p.H);
          const int row0 = b * S;                                       // first row of this image in the [B*S, 3*H*64] matrix
          const int n_tiles = pair_tiles(pr);
          if (item + G < n_items) load_q(item + G);                     // next item's Q tiles, one item ahead
          for (int j = 0; j < n_tiles; ++j, ++g) {
            const int st = g % KV_STAGES;
            mbar_wait(kv_empty(st), ((g / KV_STAGES) & 1) ^ 1);
            const uint32_t ks = kv_s + st * 2 * KV_TILE_BYTES;
            mbar_expect_tx(kv_full(st), 2 * KV_TILE_BYTES);
            tma_load_2d(ks, &map_kv, kv_full(st), (p.H + h) * HD, row0 + j * BKV);
            tma_load_2d(ks + KV_TILE_BYTES, &map_kv, kv_full(st), (2 * p.H + h) * HD, row0 + j * BKV);
          }
        }
      }
    } else {
      // =========================================================== MMA issuer of stream s = warp - 1
      if (lane == 0) {
        const int s = warp - 1;
        const uint32_t s_tmem = tmem_base + 128 * s, o_tmem = s_tmem + 64;
        const uint32_t idesc_qk = make_idesc(BQ, BKV, FP16 ? 1 : 0, 0);        // S[128 x 64 keys]: B = K tile, K-major (d contiguous)
        const uint32_t idesc_pv = make_idesc(BQ, HD, FP16 ? 1 : 0, 1);         // O[128 x 64 dims]: B = V tile, MN-major (d contiguous)
        int g = 0;                   // CTA-global K / V tile counter (ring stage / parity), the same walk as the producer's
        int t = 0;                   // stream-local tile counter (parities of s_full / p_ready / pv_done)
        int uq = 0;                  // stream-local item counter (Q buffer / parities of q_full / q_empty)
        for (int item = blockIdx.x; item < n_items; item += G) {
          const int pr = item % npair;
          const int n_pair = pair_tiles(pr), n_mine = tiles_of(2 * pr + s);
          const int qb = uq & 1;
          const uint32_t qs = q_s + (2 * s + qb) * Q_BYTES;
          auto issue_qk = [&](int gg, bool last) {                            // S = Q K_gg^T; `last` also releases the Q buffer
            const int st = gg % KV_STAGES;
            mbar_wait_spin(kv_full(st), (gg / KV_STAGES) & 1);
            tc_fence_after();
            const uint32_t ks = kv_s + st * 2 * KV_TILE_BYTES;
#pragma unroll
            for (int k = 0; k < HD / 16; ++k)                                  // K dimension = head dim: 32 B per k-step inside the row
              tc_mma_f16(s_tmem, make_smem_desc(qs + k * 32), make_smem_desc(ks + k * 32), idesc_qk, k > 0 ? 1u : 0u);
            tc_commit(s_full(s));
            if (last) tc_commit(q_empty(s, qb));
            TRACE(12, gg);
          };
          if (n_mine > 0) {
            mbar_wait(q_full(s, qb), (uq >> 1) & 1);
            issue_qk(g, n_mine == 1);
            ++uq;
          }
          for (int j = 0; j < n_pair; ++j, ++g) {
            const int st = g % KV_STAGES;
            if (j < n_mine) {
              const uint32_t vs = kv_s + st * 2 * KV_TILE_BYTES + KV_TILE_BYTES;
              mbar_wait_spin(p_ready(s), t & 1);                               // P_t in TMEM, O rescaled (or read out)
              TRACE(10, g);
              tc_fence_after();
#pragma unroll
              for (int k = 0; k < BKV / 16; ++k)                               // K dimension = keys: 16 keys = 8 packed TMEM columns
                tc_mma_f16_ts(o_tmem, s_tmem + 8 * k, make_smem_desc(vs + k * 2048), idesc_pv, (j > 0 || k > 0) ? 1u : 0u);
              tc_commit(kv_empty(st));                                         // my half of the stage release (Q K^T_g and P V_g retired)
              tc_commit(pv_done(s));
              TRACE(11, g);
              ++t;
              if (j + 1 < n_mine) issue_qk(g + 1, j + 2 == n_mine);            // next S straight behind P V (same accumulator columns)
            } else {
              // this stream needs fewer key tiles than its partner (or is absent): hand the stage back once it has landed
              mbar_wait(kv_full(st), (g / KV_STAGES) & 1);
              mbar_arrive(kv_empty(st));
            }
          }
        }
      }
    }
  } else {
    // =========================================================== softmax / correction / epilogue: one thread per query row
    const int s = (warp - 3) >> 2, quarter = warp & 3;                  // TMEM lane quarter = warp % 4 (hardware rule)
    const int rl = quarter * 32 + lane;                                  // row inside the tile = TMEM lane
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    const uint32_t s_tmem = tmem_base + 128 * s + lane_addr, o_tmem = s_tmem + 64;
    int t = 0;                                                           // stream-local tile counter
    for (int item = blockIdx.x; item < n_items; item += G) {
      const int pr = item % npair, h = (item / npair) % p.H, b = item / (npair * p.H);
      const int qt = 2 * pr + s;
      const int n_tiles = tiles_of(qt);
      if (n_tiles == 0) continue;                                        // absent stream (odd tile count)
      const int row = qt * BQ + rl;
      const int kmax = (row < p.ctx_rows) ? p.ctx_keys : S;
      // a quarter that lies entirely past the end of the sequence does no softmax work: it only keeps pace with the barriers
      const bool warp_valid = qt * BQ + quarter * 32 < S;
      const int kmax_w = (qt * BQ + quarter * 32 < p.ctx_rows) ? p.ctx_keys : S;     // smallest key limit of the warp's rows
      float m_run = -INFINITY, l_run = 0.f;
      for (int j = 0; j < n_tiles; ++j, ++t) {
        mbar_wait(s_full(s), t & 1);
        TRACE(1, t);
        if (warp_valid) {
          tc_fence_after();
          const int k0 = j * BKV;
          const bool masked_tile = k0 + BKV > kmax_w;                    // warp-uniform: some row of the warp loses keys here
          // ---- pass 1: row maximum over the 64 scores, one 32-column half at a time (80 registers per thread)
          float mx = -INFINITY;
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            uint32_t a[32];
            tmem_ld32(s_tmem + 32 * hh, a);
            tmem_ld_wait();
            if (masked_tile) {
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (k0 + 32 * hh + i >= kmax) a[i] = 0xff800000u;        // -inf
            }
            float mp[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) mp[i] = __uint_as_float(a[i]);
#pragma unroll
            for (int i = 4; i < 32; ++i) mp[i & 3] = fmaxf(mp[i & 3], __uint_as_float(a[i]));
            mx = fmaxf(mx, fmaxf(fmaxf(mp[0], mp[1]), fmaxf(mp[2], mp[3])));
          }
          // lazy rescale: the reference maximum only moves when the running maximum grew by more than 2^8 (P <= 256 stays exact
          // enough in 16 bits; the final O / l normalisation cancels the stale offset).  (-inf - -inf = NaN keeps m_run.)
          float m_new = fmaxf(m_run, mx * p.scale_log2e);
          if (m_new - m_run <= kRescaleThreshold) m_new = m_run;
          const float sub = (m_new == -INFINITY) ? 0.f : m_new;
          const float corr = (m_new == m_run || m_new == -INFINITY) ? 1.f : ex2_approx(m_run - m_new);
          const float nsub = -sub;
          // ---- pass 2: P = 2^(s * scale - m), one 32-key half at a time: scale / subtract and the row sum on the packed fp32
          // pipe (FFMA2 / FADD2), ex2.approx on the MUFU pipe, one cvt per pair; the 16 packed columns of a half overwrite S
          // columns this thread has already consumed (P = columns [0,32) of the stream's S region)
          float rs = 0.f;
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            uint32_t r0[32], w[16];
            tmem_ld32(s_tmem + 32 * hh, r0);
            tmem_ld_wait();
            if (masked_tile) {
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (k0 + 32 * hh + i >= kmax) r0[i] = 0xff800000u;
            }
            float rsp[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int q = 0; q < 16; ++q) {
              float x0, x1;
              scale_sub2(r0[2 * q], r0[2 * q + 1], p.scale_log2e, nsub, x0, x1);
              const float e0 = ex2_approx(x0), e1 = ex2_approx(x1);
              add2(rsp[2 * (q & 1)], rsp[2 * (q & 1) + 1], e0, e1);
              w[q] = pack2_16(e0, e1, FP16);
            }
            rs += (rsp[0] + rsp[1]) + (rsp[2] + rsp[3]);
            tmem_st16(s_tmem + 16 * hh, w);
          }
          l_run = l_run * corr + rs;
          m_run = m_new;
          // rescale this row's 64 output dims only when some row of the warp moved its maximum (never on the first tile)
          if (j > 0 && !__all_sync(0xffffffffu, corr == 1.0f)) {
            mbar_wait(pv_done(s), (t - 1) & 1);                             // P V of the previous tile retired: O is stable
            tc_fence_after();
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
              uint32_t r0[32];
              tmem_ld32(o_tmem + 32 * hh, r0);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 32; ++i) r0[i] = __float_as_uint(__uint_as_float(r0[i]) * corr);
              tmem_st32(o_tmem + 32 * hh, r0);
            }
          }
          tmem_st_wait();                                                   // P (and the rescaled O) are in TMEM
          TRACE(4, t);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_ready(s));
      }
      // ---- item epilogue: O / l -> the 16-bit planes of the proj GEMM (and / or fp32).  The other three streams of the SM
      // keep the pipes busy while this one waits for its last P V.
      mbar_wait(pv_done(s), (t - 1) & 1);
      if (warp_valid) {                 // warp-uniform: tcgen05.ld is .sync.aligned -- every lane runs it, only the stores are per row
        tc_fence_after();
        const bool row_ok = row < S;
        const float inv = 1.0f / l_run;
        const AttnOut& o = p.out;
        const bool inA = row < o.split;
        const int64_t orow = inA ? ((int64_t)b * o.split + row) : ((int64_t)b * (S - o.split) + (row - o.split));
        float* of = inA ? o.f32_a : o.f32_b;
        uint16_t* oh = reinterpret_cast<uint16_t*>(inA ? o.hi_a : o.hi_b);
        uint16_t* ol = reinterpret_cast<uint16_t*>(inA ? o.lo_a : o.lo_b);
        const int64_t off = orow * o.ld + (int64_t)h * HD;
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          uint32_t r0[32];
          tmem_ld32(o_tmem + 32 * hh, r0);
          tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            float y[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) y[q] = __uint_as_float(r0[c * 8 + q]) * inv;
            const int64_t oo = off + 32 * hh + c * 8;
            if (of && row_ok) {
              *reinterpret_cast<float4*>(of + oo) = make_float4(y[0], y[1], y[2], y[3]);
              *reinterpret_cast<float4*>(of + oo + 4) = make_float4(y[4], y[5], y[6], y[7]);
            }
            if (oh && row_ok) {
              uint32_t hp[4];
#pragma unroll
              for (int q = 0; q < 4; ++q) hp[q] = pack2_sat16(y[2 * q], y[2 * q + 1], FP16);
              *reinterpret_cast<uint4*>(oh + oo) = make_uint4(hp[0], hp[1], hp[2], hp[3]);
              if (!FP16 && ol) {                                             // bf16 residual planes (split-bf16 consumers)
                uint32_t lp[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) lp[q] = pack2_resid_bf16(y[2 * q], y[2 * q + 1], hp[q]);
                *reinterpret_cast<uint4*>(ol + oo) = make_uint4(lp[0], lp[1], lp[2], lp[3]);
              }
            }
          }
        }
      }
      // the O columns are overwritten by the first P V of the next item, which waits for a p_ready that this warp only
      // arrives at after the reads above (tcgen05.wait::ld inside the loop) -- no extra barrier needed
      tc_fence_before();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS) : "memory");
  }
}

int g_num_sms_dev[64];
bool g_attr_dev[64];       // cudaFuncSetAttribute is per device: one handle per GPU may live in the same process

}  // namespace

int launch_attention_tc6(const __nv_bfloat16* qkv16, int B, int S, int H, int ctx_rows, int ctx_keys, const AttnOut& out,
                         cudaStream_t s, int fp16) {
  STK_CHECK(qkv16 && B > 0 && S > 0 && H > 0, -1, "attention_tc6: bad arguments");
  STK_CHECK(out.ld % 8 == 0, -1, "attention_tc6: output pitch must be a multiple of 8");
  STK_CHECK(ctx_keys <= S && ctx_rows <= S && ctx_keys >= 0 && ctx_rows >= 0, -1, "attention_tc6: context limits exceed the sequence");
  STK_CHECK(ctx_rows == 0 || ctx_keys > 0, -1, "attention_tc6: context rows need at least one visible key");
  STK_TRY(gemm_tc_init());
  int dev = 0;
  STK_CUDA(cudaGetDevice(&dev));
  STK_CHECK(dev >= 0 && dev < 64, -1, "attention_tc6: device ordinal out of range");
  if (!g_attr_dev[dev]) {
    STK_CUDA(cudaDeviceGetAttribute(&g_num_sms_dev[dev], cudaDevAttrMultiProcessorCount, dev));
    STK_CUDA(cudaFuncSetAttribute(attention_tc6_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    STK_CUDA(cudaFuncSetAttribute(attention_tc6_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    g_attr_dev[dev] = true;
    if (getenv("SELFTOK_DEBUG")) {
      int occ = 0;
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, attention_tc6_kernel<true>, NUM_THREADS, SMEM_BYTES);
      cudaFuncAttributes fa;
      cudaFuncGetAttributes(&fa, attention_tc6_kernel<true>);
      fprintf(stderr, "attention_tc6: %d resident CTAs per SM (%d registers, %d B dynamic smem, %zu B local)\n", occ, fa.numRegs, SMEM_BYTES,
              (size_t)fa.localSizeBytes);
    }
  }
  CUtensorMap mq, mkv;
  const uint64_t rows = (uint64_t)B * S, cols = (uint64_t)3 * H * HD;
  STK_TRY(make_tensor_map_2d(&mq, qkv16, rows, cols, BQ, HD, fp16));
  STK_TRY(make_tensor_map_2d(&mkv, qkv16, rows, cols, BKV, HD, fp16));
  const int nq = (S + BQ - 1) / BQ, npair = (nq + 1) / 2;
  const int n_items = npair * H * B;
  const int grid = std::min(n_items, 2 * g_num_sms_dev[dev]);      // persistent: two CTAs (four streams) per SM
  Attn6Params p{out, B, S, H, ctx_rows, ctx_keys, fp16, 0.125f * 1.4426950408889634f};
  if (fp16) attention_tc6_kernel<true><<<grid, NUM_THREADS, SMEM_BYTES, s>>>(mq, mkv, p);
  else attention_tc6_kernel<false><<<grid, NUM_THREADS, SMEM_BYTES, s>>>(mq, mkv, p);
  count_launch();
  STK_CUDA(cudaGetLastError());
  return 0;
}

#ifdef SELFTOK_ATTN_TRACE
extern "C" __attribute__((visibility("default"))) int selftok_dbg_attn_trace6(unsigned long long* out_host, int max_n) {
  cudaDeviceSynchronize();
  if (max_n < 16 * 4096) return -1;
  cudaMemcpyFromSymbol(out_host, g_trace6, sizeof(unsigned long long) * 16 * 4096);
  static unsigned long long zeros[16 * 4096];
  cudaMemcpyToSymbol(g_trace6, zeros, sizeof(zeros));
  return 16 * 4096;
}
#endif

}  // namespace stk
