"""Applies the half-chunk software pipeline to the epilogue of a copy of gemm_tc.cu (measured slower at the 168-register budget;
kept as a record of the variant):  python experiments/round2/epilogue_pipeline_patch.py <in.cu> <out.cu> [maxnreg]"""
import sys

src, dst = sys.argv[1], sys.argv[2]
maxnreg = int(sys.argv[3]) if len(sys.argv) > 3 else 0
s = open(src).read()
a = s.index('#pragma unroll 1\n  for (int c = 0; c < (BN / 2) / 32; ++c) {')
b = s.index('// Issued by the epilogue warps BEFORE they wait for the accumulator')
new = '''  constexpr int CHUNKS = (BN / 2) / 32;
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 res[4], gt[4], bias = zero4, bias_next = zero4;
  auto issue_loads = [&](int u) {                       // u = 2 * chunk + half
    const int c = u >> 1, h = u & 1;
    const int n = n_first + c * 32 + cq * 4;
    const bool col_ok = n < N;
    if (h == 0) bias_next = (e.bias && col_ok) ? ldg4(e.bias + n) : zero4;
    if (MODE == EPI_RESID) {
      const float4 g0 = (e.gate && !per_row_gate && col_ok) ? ldg4(e.gate + n) : make_float4(1.f, 1.f, 1.f, 1.f);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int kk = 4 * h + k;
        const bool ok = col_ok && rvalid[kk];
        res[k] = ok ? ldg4(e.resid + (int64_t)orow_[kk] * e.ldo + n) : zero4;
        gt[k] = (per_row_gate && ok) ? ldg4(e.gate + (int64_t)mrow[kk] * e.gate_ld + n) : g0;
      }
    }
  };
  if (n_first < N) issue_loads(0);
#pragma unroll 1
  for (int c = 0; c < CHUNKS; ++c) {
    const int n0 = n_first + c * 32;
    if (n0 >= N) break;                                 // warp-uniform
    const int n = n0 + cq * 4;
    const bool col_ok = n < N;
    {
      uint32_t r[32];
      tmem_ld32(tmem_addr + (uint32_t)(c * 32), r);
      tmem_ld_wait();
      __syncwarp();
#pragma unroll
      for (int q = 0; q < 8; ++q)
        *reinterpret_cast<uint4*>(&stage[lane * 32 + ((q ^ (lane & 7)) << 2)]) = make_uint4(r[4 * q], r[4 * q + 1], r[4 * q + 2], r[4 * q + 3]);
      __syncwarp();
    }
    bias = bias_next;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      float4 y[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int row = 4 * (4 * h + k) + rsub;
        const float4 v = *reinterpret_cast<const float4*>(&stage[row * 32 + ((cq ^ (row & 7)) << 2)]);
        y[k] = make_float4(v.x + bias.x, v.y + bias.y, v.z + bias.z, v.w + bias.w);
        if (GELU) y[k] = gelu_tanh_fast4(y[k]);
        if (MODE == EPI_RESID) {
          y[k].x = fmaf(gt[k].x, y[k].x, res[k].x); y[k].y = fmaf(gt[k].y, y[k].y, res[k].y);
          y[k].z = fmaf(gt[k].z, y[k].z, res[k].z); y[k].w = fmaf(gt[k].w, y[k].w, res[k].w);
        }
      }
      const int u_next = 2 * c + h + 1;
      if (u_next < 2 * CHUNKS && n_first + (u_next >> 1) * 32 < N) issue_loads(u_next);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int kk = 4 * h + k;
        if (!(col_ok && rvalid[kk])) continue;
        const int64_t o = (int64_t)orow_[kk] * e.ldo + n;
        if (MODE == EPI_SPLIT) {
          *reinterpret_cast<uint2*>(e.out_hi + o) = make_uint2(pack2(y[k].x, y[k].y, fp16), pack2(y[k].z, y[k].w, fp16));
          if (e.out_lo) {
            const float lx = y[k].x - __bfloat162float(__float2bfloat16_rn(y[k].x)), ly = y[k].y - __bfloat162float(__float2bfloat16_rn(y[k].y));
            const float lz = y[k].z - __bfloat162float(__float2bfloat16_rn(y[k].z)), lw = y[k].w - __bfloat162float(__float2bfloat16_rn(y[k].w));
            *reinterpret_cast<uint2*>(e.out_lo + o) = make_uint2(pack2(lx, ly, false), pack2(lz, lw, false));
          }
        } else {
          if (per_row_add) {
            const float4 a = ldg4(e.addtab + (int64_t)mrow[kk] * e.add_ld + n);
            y[k].x += a.x; y[k].y += a.y; y[k].z += a.z; y[k].w += a.w;
          }
          *reinterpret_cast<float4*>(e.out + o) = y[k];
        }
      }
    }
  }
}

'''
s = s[:a] + new + s[b:]
if maxnreg:
    s = s.replace('__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)', '__global__ void __cluster_dims__(2, 1, 1) __maxnreg__(%d)' % maxnreg)
open(dst, 'w').write(s)
