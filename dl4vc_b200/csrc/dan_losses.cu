// dan_losses.cu — the trainer's per-step losses, their gradient with respect to the model outputs and the "close example" flags in ONE
// kernel on the device: replaces dl4vc/trainer.py:252-255,309-313,426-427 + dl4vc/objectives.py:75-112 and the three host
// round trips per step around them (trainer.py:258,263,267).
//
//   binary / genotype heads: SoftBCEWithLogitsFocalLoss (objectives.py:49-112): smoothed one-hot target t, ce_k = w_ex * BCEWithLogits(z_k, t_k),
//     p = softmax(z), pt_k = t_k p_k + (1 - t_k)(1 - p_k), focal weight (1 - pt_k)^gamma * pos_weight_k / sum(pos_weight), loss = mean_ex alpha sum_k fw_k ce_k
//     (the focal weight is part of the autograd graph in the reference, so its derivative is included); close flag = sum_k |p_k - t_k| / 2 <= smoothing * window
//   allele frequency: F.binary_cross_entropy(sigmoid output, target, weight = w_ex), coverage: F.mse_loss, variant / reference base: F.cross_entropy with the
//     class weights of trainer.py:312-313
//   total = bin * binary_weight + (vt + af * aux_allele + cov + (vb + vr) * aux_bases) * aux     (trainer.py:426-427)
#include "dan_internal.h"

namespace {

__device__ __forceinline__ double block_sum(double v, double* sh) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sh[w];
  return t;
}

// focal soft-BCE of one example over n classes: returns the example's loss, writes d(loss)/dz (unscaled by 1/B) and the close flag
template <int NC>
__device__ float focal_soft_bce(const float* z, int target, float w_ex, const dan_loss_config& c, const float (&pw)[NC], float* dz, bool* close) {
  const float smooth = c.label_smoothing, conf = 1.f - smooth, off = smooth / (float)(NC - 1);
  float mx = z[0];
#pragma unroll
  for (int k = 1; k < NC; ++k) mx = fmaxf(mx, z[k]);
  float p[NC], t[NC], e = 0.f, pws = 0.f;
#pragma unroll
  for (int k = 0; k < NC; ++k) { p[k] = expf(z[k] - mx); e += p[k]; pws += pw[k]; t[k] = k == target ? conf : off; }
  float loss = 0.f, dist = 0.f, ce[NC], fw[NC], dfw_dpt[NC];
#pragma unroll
  for (int k = 0; k < NC; ++k) {
    p[k] = fminf(fmaxf(p[k] / e, 0.f), 1.f);
    ce[k] = w_ex * (fmaxf(z[k], 0.f) - z[k] * t[k] + log1pf(expf(-fabsf(z[k]))));      // F.binary_cross_entropy_with_logits, reduction none
    const float pt = t[k] * p[k] + (1.f - t[k]) * (1.f - p[k]);
    const float om = 1.f - pt;
    const float cw = pw[k] / pws;
    if (c.focal_gamma > 0.f) {
      fw[k] = powf(om, c.focal_gamma) * cw;
      dfw_dpt[k] = om > 0.f ? -c.focal_gamma * powf(om, c.focal_gamma - 1.f) * cw : 0.f;
    } else { fw[k] = cw; dfw_dpt[k] = 0.f; }
    loss += c.focal_alpha * fw[k] * ce[k];
    dist += fabsf(p[k] - t[k]);
  }
  *close = dist * 0.5f <= smooth * c.close_match_window;
  // d/dz_j: alpha * [ fw_j * w_ex * (sigmoid(z_j) - t_j) + sum_k ce_k * dfw_k/dpt_k * (2 t_k - 1) * p_k (delta_kj - p_j) ]
  float s = 0.f;                       // sum_k ce_k dfw_dpt_k (2 t_k - 1) p_k
#pragma unroll
  for (int k = 0; k < NC; ++k) s += ce[k] * dfw_dpt[k] * (2.f * t[k] - 1.f) * p[k];
#pragma unroll
  for (int j = 0; j < NC; ++j) {
    const float sig = 1.f / (1.f + expf(-z[j]));
    dz[j] = c.focal_alpha * (fw[j] * w_ex * (sig - t[j]) + ce[j] * dfw_dpt[j] * (2.f * t[j] - 1.f) * p[j] - p[j] * s);
  }
  return loss;
}

__global__ void __launch_bounds__(256) dan_losses_kernel(const float* __restrict__ heads, int B, const int32_t* __restrict__ tbin, const int32_t* __restrict__ tvt,
                                                         const float* __restrict__ taf, const float* __restrict__ tcov, const int32_t* __restrict__ tvb,
                                                         const int32_t* __restrict__ tvr, const float* __restrict__ wex, dan_loss_config c,
                                                         float* __restrict__ losses, float* __restrict__ dheads, uint8_t* __restrict__ close_vt, uint8_t* __restrict__ close_bin) {
  __shared__ double sh[8];
  const float base_w[DAN_VOCAB] = {0.001f, 1.f, 1.f, 1.f, 1.f, 1.f, 0.001f, 0.001f, 1.f, 0.001f};      // trainer.py:312-313
  // pass 1: normalisers of the two weighted cross entropies (sum of the class weight of every example's target)
  double swb = 0.0, swr = 0.0;
  for (int b = threadIdx.x; b < B; b += blockDim.x) { swb += base_w[min(max(tvb[b], 0), DAN_VOCAB - 1)]; swr += base_w[min(max(tvr[b], 0), DAN_VOCAB - 1)]; }
  swb = block_sum(swb, sh); swr = block_sum(swr, sh);
  const float invB = 1.f / (float)B;
  const float g_bin = c.binary_weight, g_vt = c.aux_weight, g_af = c.aux_weight * c.aux_allele_weight, g_cov = c.aux_weight, g_base = c.aux_weight * c.aux_bases_weight;
  double l_bin = 0.0, l_vt = 0.0, l_af = 0.0, l_cov = 0.0, l_vb = 0.0, l_vr = 0.0, n_close = 0.0;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    const float* h = heads + (long)b * DAN_NUM_HEAD_OUTPUTS;
    float* d = dheads + (long)b * DAN_NUM_HEAD_OUTPUTS;
    const float w = wex ? wex[b] : 1.f;
    bool cb, cv;
    {
      const float pw2[2] = {c.fp_train_weight, 1.f};
      float dz[2];
      l_bin += focal_soft_bce<2>(h, tbin[b], w, c, pw2, dz, &cb);
      d[0] = dz[0] * invB * g_bin; d[1] = dz[1] * invB * g_bin;
    }
    {
      const float pw3[3] = {c.fp_train_weight, 1.f, 1.f};
      float dz[3];
      l_vt += focal_soft_bce<3>(h + 2, tvt[b], w, c, pw3, dz, &cv);
      d[2] = dz[0] * invB * g_vt; d[3] = dz[1] * invB * g_vt; d[4] = dz[2] * invB * g_vt;
    }
    close_vt[b] = cv; close_bin[b] = cb;
    n_close += cv ? 1.0 : 0.0;
    {   // allele frequency: binary_cross_entropy on the sigmoid output (logs clamped at -100, gradient denominator at 1e-12, like ATen)
      const float a = h[5], y = taf[b];
      l_af += w * -(y * fmaxf(logf(a), -100.f) + (1.f - y) * fmaxf(logf(1.f - a), -100.f));
      d[5] = w * (a - y) / fmaxf((1.f - a) * a, 1e-12f) * invB * g_af;
    }
    {   // coverage: mse
      const float df = h[6] - tcov[b];
      l_cov += df * df;
      d[6] = 2.f * df * invB * g_cov;
    }
#pragma unroll
    for (int which = 0; which < 2; ++which) {      // variant / reference base: weighted cross entropy over the 10 tokens
      const float* v = h + 7 + 10 * which;
      const int y = min(max(which ? tvr[b] : tvb[b], 0), DAN_VOCAB - 1);
      const float cw = base_w[y];
      const double sw = which ? swr : swb;
      float mx = v[0];
      for (int k = 1; k < DAN_VOCAB; ++k) mx = fmaxf(mx, v[k]);
      float e = 0.f;
      for (int k = 0; k < DAN_VOCAB; ++k) e += expf(v[k] - mx);
      const float lse = mx + logf(e);
      (which ? l_vr : l_vb) += cw * (lse - v[y]);
      for (int k = 0; k < DAN_VOCAB; ++k) d[7 + 10 * which + k] = cw * (expf(v[k] - lse) - (k == y ? 1.f : 0.f)) / (float)sw * g_base;
    }
  }
  l_bin = block_sum(l_bin, sh) * invB; l_vt = block_sum(l_vt, sh) * invB; l_af = block_sum(l_af, sh) * invB; l_cov = block_sum(l_cov, sh) * invB;
  l_vb = block_sum(l_vb, sh) / swb; l_vr = block_sum(l_vr, sh) / swr; n_close = block_sum(n_close, sh);
  if (threadIdx.x == 0) {
    losses[0] = (float)l_bin; losses[1] = (float)l_vt; losses[2] = (float)l_af; losses[3] = (float)l_cov; losses[4] = (float)l_vb; losses[5] = (float)l_vr;
    losses[6] = (float)(l_bin * g_bin + l_vt * g_vt + l_af * g_af + l_cov * g_cov + (l_vb + l_vr) * g_base);
    losses[7] = (float)n_close;
  }
}

// close_table[idx[b]] = flag[b]   (trainer.py:263 / dataset.py:480: the per-example "easy" table the next epoch's sampler reads), on the device
__global__ void close_table_update_kernel(uint8_t* __restrict__ table, long table_len, const int64_t* __restrict__ idx, const uint8_t* __restrict__ flag, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < B && idx[b] >= 0 && idx[b] < table_len) table[idx[b]] = flag[b];
}

}  // namespace

extern "C" {

int dan_losses(const float* heads, int batch, const int32_t* target_binary, const int32_t* target_var_type, const float* target_allele_freq,
               const float* target_coverage, const int32_t* target_var_base, const int32_t* target_ref_base, const float* example_weight,
               const dan_loss_config* cfg, float* losses_out, float* dheads_out, uint8_t* close_vt, uint8_t* close_bin, void* stream) {
  if (batch < 1) { dan_set_error("loss of an empty batch"); return DAN_E_INVALID; }
  if (!heads || !target_binary || !target_var_type || !target_allele_freq || !target_coverage || !target_var_base || !target_ref_base || !cfg || !losses_out ||
      !dheads_out || !close_vt || !close_bin) { dan_set_error("null pointer"); return DAN_E_INVALID; }
  if (!(cfg->label_smoothing >= 0.f && cfg->label_smoothing <= 1.f) || cfg->focal_gamma < 0.f) { dan_set_error("label smoothing outside [0, 1] or negative focal gamma (objectives.py:73, trainer.py:100)"); return DAN_E_INVALID; }
  dan_losses_kernel<<<1, 256, 0, static_cast<cudaStream_t>(stream)>>>(heads, batch, target_binary, target_var_type, target_allele_freq, target_coverage, target_var_base,
                                                                      target_ref_base, example_weight, *cfg, losses_out, dheads_out, close_vt, close_bin);
  dan_count_launch(1);
  DAN_CUDA_TRY(cudaGetLastError());
  return DAN_OK;
}

int dan_close_table_update(uint8_t* table, int64_t table_len, const int64_t* idx, const uint8_t* flags, int batch, void* stream) {
  if (batch < 0 || !table || (batch > 0 && (!idx || !flags))) { dan_set_error("bad argument"); return DAN_E_INVALID; }
  if (batch == 0) return DAN_OK;
  close_table_update_kernel<<<(batch + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(table, table_len, idx, flags, batch);
  dan_count_launch(1);
  DAN_CUDA_TRY(cudaGetLastError());
  return DAN_OK;
}

}  // extern "C"
