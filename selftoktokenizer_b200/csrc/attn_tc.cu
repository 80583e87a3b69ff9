// Joint attention of the MMDiT on tensor cores (sm_100a), head_dim 64, non-causal, optional two-range rule.
//
// softmax(Q K^T / 8) V over the joint sequence [context prefix ; image tokens] (sd3/mmdit.py:521-531,
// sd3/other_impls.py:37-45).  After the masked context rows are dropped the decode attention is plain dense
// attention; the renderer keeps one rule: query rows < ctx_rows only see keys < ctx_keys (mmdit.py:1581).
//
// One CTA = 64 queries of one (image, head); 4 warps x 16 query rows; keys/values streamed in tiles of 64.
// Products run on mma.sync.m16n8k16 bf16 with fp32 accumulation; in the NSPLIT == 3 mode every operand is split
// into bf16 hi + lo and each product is three MMAs (hi*hi + hi*lo + lo*hi), i.e. ~2^-17 relative error — the same
// fp32-faithful recipe as gemm_tc.cu.  Softmax statistics, the running rescale and the final normalisation are fp32.
// Round-1 note: this is the warp-level (legacy) tensor path; the tcgen05/TMEM version is the round-2 item.
#include "common.cuh"
#include "kernels.h"

namespace stk {
namespace {

constexpr int HD = 64, BQ = 64, BKV = 64, PITCH = 72;   // bf16 elements per smem row (64 + 8 pad -> conflict-free fragments)

template <int FP16>
__device__ __forceinline__ void mma_16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  if (FP16)
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  else
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], const void* smem_row) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(smem_row);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
template <int FP16>
__device__ __forceinline__ void split2(float x, float y, uint32_t& hi, uint32_t& lo) {
  uint16_t hx, lx, hy, ly;
  split16(x, FP16 != 0, hx, lx);
  split16(y, FP16 != 0, hy, ly);
  hi = hx | ((uint32_t)hy << 16);
  lo = lx | ((uint32_t)ly << 16);
}

struct AttnTcParams {
  const uint16_t* qkv_hi;   // [B, S, 3, H, 64] 16-bit planes written by the QKV GEMM epilogue (bf16 hi / IEEE half)
  const uint16_t* qkv_lo;   // bf16 lo plane (NSPLIT == 3) or NULL
  AttnOut out;
  int S, H, ctx_rows, ctx_keys;
  float scale_log2e;
};

template <int NSPLIT, int FP16>
__global__ void __launch_bounds__(128) attention_tc_kernel(const AttnTcParams p) {
  __shared__ __align__(16) __nv_bfloat16 Ks[2][BKV][PITCH];     // [hi/lo][key][d]
  __shared__ __align__(16) __nv_bfloat16 Vs[2][BKV][PITCH];     // [hi/lo][key][d]  (B fragments via ldmatrix.trans)
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, tig = lane & 3;
  const int q0 = blockIdx.x * BQ, h = blockIdx.y, b = blockIdx.z;
  const int S = p.S;
  const int64_t row_stride = (int64_t)3 * p.H * HD;
  const int64_t base_off = (int64_t)b * S * row_stride + (int64_t)h * HD;
  const uint16_t* base_hi = p.qkv_hi + base_off;
  const uint16_t* base_lo = NSPLIT == 3 ? p.qkv_lo + base_off : nullptr;
  // ---- Q fragments (rows r0 = q0 + warp*16 + g, r1 = r0 + 8), kept for the whole kernel
  const int r0 = q0 + warp * 16 + g, r1 = r0 + 8;
  uint32_t qh[4][4], ql[4][4];
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    const int d0 = kk * 16 + tig * 2;
    const int64_t o0 = (int64_t)r0 * row_stride + d0, o1 = (int64_t)r1 * row_stride + d0;
    qh[kk][0] = r0 < S ? *reinterpret_cast<const uint32_t*>(base_hi + o0) : 0u;
    qh[kk][1] = r1 < S ? *reinterpret_cast<const uint32_t*>(base_hi + o1) : 0u;
    qh[kk][2] = r0 < S ? *reinterpret_cast<const uint32_t*>(base_hi + o0 + 8) : 0u;
    qh[kk][3] = r1 < S ? *reinterpret_cast<const uint32_t*>(base_hi + o1 + 8) : 0u;
    if (NSPLIT == 3) {
      ql[kk][0] = r0 < S ? *reinterpret_cast<const uint32_t*>(base_lo + o0) : 0u;
      ql[kk][1] = r1 < S ? *reinterpret_cast<const uint32_t*>(base_lo + o1) : 0u;
      ql[kk][2] = r0 < S ? *reinterpret_cast<const uint32_t*>(base_lo + o0 + 8) : 0u;
      ql[kk][3] = r1 < S ? *reinterpret_cast<const uint32_t*>(base_lo + o1 + 8) : 0u;
    }
  }
  float o[8][4];
#pragma unroll
  for (int n = 0; n < 8; ++n) { o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f; }
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
  const int kmax0 = (r0 < p.ctx_rows) ? p.ctx_keys : S;
  const int kmax1 = (r1 < p.ctx_rows) ? p.ctx_keys : S;
  const int kmax_cta = (q0 + BQ <= p.ctx_rows) ? p.ctx_keys : S;

  for (int k0 = 0; k0 < kmax_cta; k0 += BKV) {
    __syncthreads();
    // ---- stage the K and V tiles (already 16-bit in HBM): 64 keys x 128 B per plane, 16 B per thread-copy
    for (int f = tid; f < BKV * 8; f += 128) {
      const int key = f >> 3, ch = (f & 7) * 8;
      uint4 kh = make_uint4(0, 0, 0, 0), vh = kh, kl = kh, vl = kh;
      if (k0 + key < S) {
        const int64_t ro = (int64_t)(k0 + key) * row_stride + ch;
        kh = *reinterpret_cast<const uint4*>(base_hi + ro + (int64_t)p.H * HD);
        vh = *reinterpret_cast<const uint4*>(base_hi + ro + (int64_t)2 * p.H * HD);
        if (NSPLIT == 3) {
          kl = *reinterpret_cast<const uint4*>(base_lo + ro + (int64_t)p.H * HD);
          vl = *reinterpret_cast<const uint4*>(base_lo + ro + (int64_t)2 * p.H * HD);
        }
      }
      *reinterpret_cast<uint4*>(&Ks[0][key][ch]) = kh;
      *reinterpret_cast<uint4*>(&Vs[0][key][ch]) = vh;
      if (NSPLIT == 3) {
        *reinterpret_cast<uint4*>(&Ks[1][key][ch]) = kl;
        *reinterpret_cast<uint4*>(&Vs[1][key][ch]) = vl;
      }
    }
    __syncthreads();
    // ---- S = Q K^T  (16 x 64 per warp)
    float sc[8][4];
#pragma unroll
    for (int n = 0; n < 8; ++n) { sc[n][0] = sc[n][1] = sc[n][2] = sc[n][3] = 0.f; }
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
      for (int n = 0; n < 8; ++n) {
        const __nv_bfloat16* kr = &Ks[0][n * 8 + g][kk * 16 + tig * 2];
        const uint32_t b0 = *reinterpret_cast<const uint32_t*>(kr), b1 = *reinterpret_cast<const uint32_t*>(kr + 8);
        mma_16<FP16>(sc[n], qh[kk], b0, b1);
        if (NSPLIT == 3) {
          const __nv_bfloat16* kl = &Ks[1][n * 8 + g][kk * 16 + tig * 2];
          const uint32_t c0 = *reinterpret_cast<const uint32_t*>(kl), c1 = *reinterpret_cast<const uint32_t*>(kl + 8);
          mma_16<FP16>(sc[n], qh[kk], c0, c1);
          mma_16<FP16>(sc[n], ql[kk], b0, b1);
        }
      }
    }
    // ---- online softmax (fp32); columns of sc[n]: key = k0 + n*8 + tig*2 (+1)
    const bool need_mask = (k0 + BKV > kmax0) || (k0 + BKV > kmax1);
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int n = 0; n < 8; ++n) {
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int key = k0 + n * 8 + tig * 2 + j;
        float s0 = sc[n][j] * p.scale_log2e, s1 = sc[n][2 + j] * p.scale_log2e;
        if (need_mask) {
          if (key >= kmax0) s0 = -INFINITY;
          if (key >= kmax1) s1 = -INFINITY;
        }
        sc[n][j] = s0; sc[n][2 + j] = s1;
        mx0 = fmaxf(mx0, s0); mx1 = fmaxf(mx1, s1);
      }
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);
    const float cr0 = (mn0 == -INFINITY) ? 1.f : exp2f(m0 - mn0), cr1 = (mn1 == -INFINITY) ? 1.f : exp2f(m1 - mn1);
    const float sub0 = (mn0 == -INFINITY) ? 0.f : mn0, sub1 = (mn1 == -INFINITY) ? 0.f : mn1;
    float rs0 = 0.f, rs1 = 0.f;
#pragma unroll
    for (int n = 0; n < 8; ++n) {
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const float p0 = exp2f(sc[n][j] - sub0), p1 = exp2f(sc[n][2 + j] - sub1);
        sc[n][j] = p0; sc[n][2 + j] = p1;
        rs0 += p0; rs1 += p1;
      }
    }
    rs0 += __shfl_xor_sync(0xffffffffu, rs0, 1); rs0 += __shfl_xor_sync(0xffffffffu, rs0, 2);
    rs1 += __shfl_xor_sync(0xffffffffu, rs1, 1); rs1 += __shfl_xor_sync(0xffffffffu, rs1, 2);
    l0 = l0 * cr0 + rs0; l1 = l1 * cr1 + rs1;
    m0 = mn0; m1 = mn1;
#pragma unroll
    for (int n = 0; n < 8; ++n) { o[n][0] *= cr0; o[n][1] *= cr0; o[n][2] *= cr1; o[n][3] *= cr1; }
    // ---- O += P V   (P fragments straight from the S accumulators)
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      uint32_t ph[4], pl[4];
      split2<FP16>(sc[2 * kk][0], sc[2 * kk][1], ph[0], pl[0]);
      split2<FP16>(sc[2 * kk][2], sc[2 * kk][3], ph[1], pl[1]);
      split2<FP16>(sc[2 * kk + 1][0], sc[2 * kk + 1][1], ph[2], pl[2]);
      split2<FP16>(sc[2 * kk + 1][2], sc[2 * kk + 1][3], ph[3], pl[3]);
      // V is row-major [key][d]; ldmatrix.trans hands each thread (k = key pair, n = d) fragments for two d-tiles at once:
      // lanes 0-7 / 8-15 address keys kk*16 + 0..7 / 8..15 of d-tile n, lanes 16-23 / 24-31 the same keys of d-tile n+1.
      const int lrow = kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
#pragma unroll
      for (int n = 0; n < 8; n += 2) {
        const int lcol = (n + (lane >> 4)) * 8;
        uint32_t b[4];
        ldmatrix_x4_trans(b, &Vs[0][lrow][lcol]);
        mma_16<FP16>(o[n], ph, b[0], b[1]);
        mma_16<FP16>(o[n + 1], ph, b[2], b[3]);
        if (NSPLIT == 3) {
          uint32_t c[4];
          ldmatrix_x4_trans(c, &Vs[1][lrow][lcol]);
          mma_16<FP16>(o[n], ph, c[0], c[1]);
          mma_16<FP16>(o[n + 1], ph, c[2], c[3]);
          mma_16<FP16>(o[n], pl, b[0], b[1]);
          mma_16<FP16>(o[n + 1], pl, b[2], b[3]);
        }
      }
    }
  }
  // ---- normalise and store: thread holds rows r0 (o[n][0..1]) and r1 (o[n][2..3]), columns n*8 + tig*2 (+1)
  const AttnOut& t = p.out;
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const int row = half ? r1 : r0;
    if (row >= S) continue;
    const float inv = 1.0f / (half ? l1 : l0);
    const bool inA = row < t.split;
    const int64_t orow = inA ? ((int64_t)b * t.split + row) : ((int64_t)b * (S - t.split) + (row - t.split));
    float* of = inA ? t.f32_a : t.f32_b;
    __nv_bfloat16* oh = inA ? t.hi_a : t.hi_b;
    __nv_bfloat16* ol = inA ? t.lo_a : t.lo_b;
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      const float y0 = o[n][half * 2] * inv, y1 = o[n][half * 2 + 1] * inv;
      const int64_t idx = orow * t.ld + (int64_t)h * HD + n * 8 + tig * 2;
      if (of) *reinterpret_cast<float2*>(of + idx) = make_float2(y0, y1);
      if (oh) {
        uint32_t hi, lo;
        split2<FP16>(y0, y1, hi, lo);
        *reinterpret_cast<uint32_t*>(oh + idx) = hi;
        if (ol) *reinterpret_cast<uint32_t*>(ol + idx) = lo;
      }
    }
  }
}

}  // namespace

int launch_attention_tc(const __nv_bfloat16* qkv_hi, const __nv_bfloat16* qkv_lo, int B, int S, int H, int nsplit,
                        int ctx_rows, int ctx_keys, const AttnOut& out, cudaStream_t s, int fp16) {
  STK_CHECK(qkv_hi && B > 0 && S > 0 && H > 0, -1, "attention_tc: bad arguments");
  STK_CHECK(nsplit != 3 || qkv_lo, -1, "attention_tc: the split mode needs the lo plane");
  STK_CHECK(nsplit == 1 || nsplit == 3, -1, "attention_tc: nsplit must be 1 or 3");
  STK_CHECK(out.ld % 2 == 0, -1, "attention_tc: output pitch must be even");
  AttnTcParams p{reinterpret_cast<const uint16_t*>(qkv_hi), reinterpret_cast<const uint16_t*>(qkv_lo), out, S, H, ctx_rows, ctx_keys,
                 0.125f * 1.4426950408889634f};
  dim3 grid((S + BQ - 1) / BQ, H, B);
  STK_CHECK(!fp16 || nsplit == 1, -1, "attention_tc: the fp16 mode is single-pass");
  if (nsplit == 3) attention_tc_kernel<3, 0><<<grid, 128, 0, s>>>(p);
  else if (fp16) attention_tc_kernel<1, 1><<<grid, 128, 0, s>>>(p);
  else attention_tc_kernel<1, 0><<<grid, 128, 0, s>>>(p);
  count_launch();
  STK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace stk
