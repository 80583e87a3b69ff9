/* TEST INFRASTRUCTURE ONLY -- plain-C restatement of the index-producing tail of the encode path.
 *
 * Only tests/ (and, if ever needed, __graft_entry__.smoke() / bench.py's cpu legs) may load this; the product path never does.
 * It is an independent (non-torch) pin of the integer work of the path: which code id each token gets.
 *
 * Follows, line by line:
 *   VectorQuantize.forward, eval branch      mimogpt/.../vector_quantize_pytorch.py:844-876   project_in (nn.Linear 512 -> 16)
 *   l2norm                                   vector_quantize_pytorch.py:51-52                 F.normalize(x, p=2, dim=-1), eps 1e-12
 *   CosineSimCodebook.forward, eval branch   vector_quantize_pytorch.py:525-563,580           dist = x_hat . embed^T ; argmax (first maximum)
 *   batched_embedding / get_output           vector_quantize_pytorch.py:310-314, 787-809      codes = embed[ids]
 *   final_layer_norm3                        models_ours.py:88,241-242                        affine LayerNorm(16), eps 1e-6, biased variance
 *
 * Arithmetic: fp32, one sequential fused multiply-add chain per dot product (the order the CUDA kernels use as well); torch's CPU
 * GEMM blocks its sums differently, so ids may legitimately differ where the reference's own top-1 / top-2 margin is at rounding
 * level -- the tests compare against the reference-generated fixtures and use the fixture's margin for exactly that.
 */
#include <math.h>
#include <stdint.h>
#include <stddef.h>

/* z [n, qdim] -> ids [n], optional x_hat [n, cdim] (normalised projection) and margin [n] (best - second best similarity) */
int selftok_oracle_vq(const float* z, int64_t n, int qdim, const float* proj_w /* [cdim, qdim] */, const float* proj_b /* [cdim] or NULL */,
                      const float* embed /* [ncode, cdim], unit-norm rows */, int ncode, int cdim, int64_t* ids, float* x_hat,
                      float* margin) {
  if (!z || !proj_w || !embed || !ids || n < 0 || qdim <= 0 || cdim <= 0 || cdim > 64 || ncode <= 0) return -1;
  for (int64_t r = 0; r < n; ++r) {
    float x[64];
    const float* zr = z + (size_t)r * qdim;
    float ss = 0.0f;
    for (int c = 0; c < cdim; ++c) {
      const float* w = proj_w + (size_t)c * qdim;
      float acc = 0.0f;
      for (int k = 0; k < qdim; ++k) acc = fmaf(zr[k], w[k], acc);
      x[c] = acc + (proj_b ? proj_b[c] : 0.0f);
      ss = fmaf(x[c], x[c], ss);
    }
    const float denom = fmaxf(sqrtf(ss), 1e-12f);                 /* F.normalize: x / max(||x||_2, eps) */
    for (int c = 0; c < cdim; ++c) x[c] = x[c] / denom;
    float best = -INFINITY, second = -INFINITY;
    int64_t best_id = 0;
    for (int code = 0; code < ncode; ++code) {
      const float* e = embed + (size_t)code * cdim;
      float s = 0.0f;
      for (int c = 0; c < cdim; ++c) s = fmaf(x[c], e[c], s);
      if (s > best) { second = best; best = s; best_id = code; }  /* strict '>' : the FIRST maximum wins, as torch.argmax */
      else if (s > second) second = s;
    }
    ids[r] = best_id;
    if (x_hat) for (int c = 0; c < cdim; ++c) x_hat[(size_t)r * cdim + c] = x[c];
    if (margin) margin[r] = best - second;
  }
  return 0;
}

/* codes = embed[ids]; out = LayerNorm(codes) * w + b   (get_output_from_indices + final_layer_norm3) */
int selftok_oracle_lookup_ln(const int64_t* ids, int64_t n, const float* embed, int ncode, int cdim, const float* ln_w, const float* ln_b,
                             float eps, float* out) {
  if (!ids || !embed || !ln_w || !ln_b || !out || cdim <= 0) return -1;
  for (int64_t r = 0; r < n; ++r) {
    if (ids[r] < 0 || ids[r] >= ncode) return -2;
    const float* e = embed + (size_t)ids[r] * cdim;
    float mean = 0.0f;
    for (int c = 0; c < cdim; ++c) mean += e[c];
    mean /= (float)cdim;
    float var = 0.0f;
    for (int c = 0; c < cdim; ++c) { const float d = e[c] - mean; var = fmaf(d, d, var); }
    var /= (float)cdim;
    const float rstd = 1.0f / sqrtf(var + eps);
    for (int c = 0; c < cdim; ++c) out[(size_t)r * cdim + c] = (e[c] - mean) * rstd * ln_w[c] + ln_b[c];
  }
  return 0;
}

/* DiTi_cont.to_indices on the sampler's integer timesteps (diti_utils.py:84-110, Segment.process :79-82; the caller passes
 * timestep_map.long(), rectified_flow.py:202): for every segment in order, wherever t - low >= 0 the index is OVERWRITTEN by
 * trunc(fp32(slope) * fp32(t - low)) + base, slope = k_seg / (high - low); finally clamped to [0, K-1].
 * stages [n_seg] are the segment upper ends (the first segment starts at 0), k_per_stage [n_seg] the tokens added per segment. */
int selftok_oracle_diti_k(const int64_t* t_mapped, int n, const int* stages, const int* k_per_stage, int n_seg, int K, int64_t* k_out) {
  if (!t_mapped || !stages || !k_per_stage || !k_out || n_seg <= 0) return -1;
  for (int i = 0; i < n; ++i) {
    int64_t ind = 0;
    int64_t base = 0;
    for (int s = 0; s < n_seg; ++s) {
      const int low = (s == 0) ? 0 : stages[s - 1];
      const float slope = (float)((double)k_per_stage[s] / (double)(stages[s] - low));
      const int64_t xp = t_mapped[i] - low;
      if (xp >= 0) ind = (int64_t)(slope * (float)xp) + base;     /* C conversion truncates toward zero, as .to(torch.long) */
      base += k_per_stage[s];
    }
    if (ind < 0) ind = 0;
    if (ind > K - 1) ind = K - 1;
    k_out[i] = ind;
  }
  return 0;
}
