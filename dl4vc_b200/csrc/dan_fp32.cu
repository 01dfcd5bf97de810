// dan_fp32.cu — fp32 (CUDA-core FFMA) implementation of the DAN forward: the accuracy path
// (logits within 1e-4 relative of the reference) and the on-device yardstick for the tcgen05 path.
//
// Layer-by-layer over "row matrices": one row per (candidate, read, position), reads laid end to end with `gap`
// zero rows between them (RowGeom). Every contraction of the network — the dilated (1x3) convolutions
// (dl4vc/model.py:749), the residual / bottleneck 1x1 convolutions (:760,:774), the (1x201) highway compression
// (:776), the FC trunk and the heads (:917-958) — is one call of a tiled SGEMM whose A operand is gathered with a
// per-tap row offset.
#include "dan_fp32_kernels.cuh"

namespace {

// workspace carve-up --------------------------------------------------------------------------------------
struct Fp32Plan {
  int S;            // candidates per conv pass
  int Bc;           // candidates per FC chunk
  long rows, rowsPad, rowsAlloc;   // data rows per pass, padded to 128, with leading/trailing gap
  long readsPad;
  size_t off_x0, off_h[4], off_t, off_pool, off_hw, off_pooled, off_fcin, off_fc[DAN_MAX_FC + 1], off_part, off_heads, total;
  long hw_layer_stride;
  int maxN;
  int splits;
};

// The (1 x 201) compression is a K = 6432 product with one 32-column tile per 128 reads: a fixed number of K ranges (independent of the
// batch shape, so results stay bitwise batch-invariant) brings a 148-candidate pass from 116 to 464 CTAs.
constexpr int kCompSplitsFp32 = 4;

Fp32Plan make_plan(const dan_model* m, int batch) {
  Fp32Plan pl{};
  pl.S = m->pass_candidates < batch ? m->pass_candidates : (batch > 0 ? batch : 1);
  pl.Bc = batch < 256 ? (batch > 0 ? batch : 1) : 256;
  pl.rows = m->geom.rows_of(pl.S);
  pl.rowsPad = (pl.rows + 127) / 128 * 128;
  pl.rowsAlloc = pl.rowsPad + 2L * m->geom.gap + 128;
  pl.readsPad = ((long)pl.S * m->R + 127) / 128 * 128;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += round_up_z(bytes, 256); return o; };
  pl.off_x0 = take((size_t)pl.rowsAlloc * m->CinPad * 4);
  for (int i = 0; i < 4; ++i) pl.off_h[i] = take((size_t)pl.rowsAlloc * m->C * 4);
  pl.off_t = take(((size_t)pl.rowsAlloc + (size_t)128 * m->geom.pitch) * (m->bott > 0 ? m->bott : 16) * 4);
  pl.off_pool = take((size_t)pl.S * m->P * m->C * 4);
  pl.hw_layer_stride = pl.readsPad * (m->bott > 0 ? m->bott : 16);
  pl.off_hw = take((size_t)m->L * pl.hw_layer_stride * 4);
  const int BcPad = round_up_i(pl.Bc, 128);
  pl.off_pooled = take((size_t)BcPad * m->pooledPad * 4);
  pl.off_fcin = take((size_t)BcPad * m->fcInPad * 4);
  pl.maxN = DAN_HEAD_PAD;
  for (int i = 0; i < m->cfg.num_fc; ++i) {
    pl.off_fc[i] = take((size_t)BcPad * m->cfg.fc_sizes[i] * 4);
    if (m->cfg.fc_sizes[i] > pl.maxN) pl.maxN = m->cfg.fc_sizes[i];
  }
  if (m->cfg.pool_combine_dimension > pl.maxN) pl.maxN = m->cfg.pool_combine_dimension;
  pl.splits = 16;
  size_t part_bytes = (size_t)pl.splits * BcPad * pl.maxN * 4;
  const size_t comp_part = (size_t)kCompSplitsFp32 * pl.readsPad * (m->bott > 0 ? m->bott : 16) * 4;      // split-K partials of the (1 x 201) compression
  if (comp_part > part_bytes) part_bytes = comp_part;
  pl.off_part = take(part_bytes);
  pl.off_heads = take((size_t)BcPad * DAN_HEAD_PAD * 4);
  pl.total = off;
  return pl;
}

// y = act(x W + b) with split-K when the contraction is long and the grid would be tiny
int run_linear(const float* A, int lda, int M, int K, const float* W, int N, int ldw, const float* bias, int relu,
               int head_act, float* out, int ldo, float* part, int max_splits, cudaStream_t st) {
  GemmParams g{};
  g.A = A; g.lda = lda; g.a_rows = M; g.M = M; g.ntaps = 1; g.tap_off[0] = 0; g.Kc = K;
  g.W = W; g.N = N; g.ldw = ldw; g.out = out; g.ldo = ldo; g.splits = 1;
  const int chunks = K / 16;
  int splits = 1;
  if (chunks >= 256) splits = max_splits;
  if (splits > 1) {
    g.out = part; g.ldo = N; g.splits = splits; g.split_stride = (long)M * N;
    int rc = launch_gemm(g, st);
    if (rc) return rc;
    splitk_finish_kernel<<<grid_for((long)M * N), 256, 0, st>>>(part, splits, g.split_stride, M, N, N, bias, relu, head_act, out, ldo);
    dan_count_launch();
    DAN_CUDA_TRY(cudaGetLastError());
    return DAN_OK;
  }
  g.bias = bias; g.relu = relu; g.head_act = head_act;
  return launch_gemm(g, st);
}

}  // namespace

// ----------------------------------------------------------------------------------------------------------
size_t dan_fp32_workspace_bytes(const dan_model* m, int batch) { return make_plan(m, batch).total; }

static EncodeParams make_encode_params(const dan_model* m, const DevInputs& in) {
  EncodeParams e{};
  e.in = in; e.emb = m->emb; e.pe = m->pe; e.D = m->cfg.embed_dim; e.Cin = m->Cin; e.CinPad = m->CinPad;
  e.use_q = m->cfg.use_q_scores; e.use_s = m->cfg.use_strands; e.use_m = m->cfg.use_reads_ref_var_mask; e.g = m->geom;
  return e;
}

static int launch_encode_fp32(const dan_model* m, const DevInputs& in, long cand0, int cands, float* rows, cudaStream_t st) {
  const size_t smem = encode_smem_bytes(m->P, m->R, m->cfg.embed_dim);
  static DanSmemAttr attr;          // models with different read counts share the process
  DAN_CUDA_TRY(attr.ensure(encode_rows_fp32_kernel, smem));
  encode_rows_fp32_kernel<<<cands, 256, smem, st>>>(make_encode_params(m, in), cand0, rows);
  dan_count_launch();
  DAN_CUDA_TRY(cudaGetLastError());
  return DAN_OK;
}

int dan_fp32_encode_reference_order(dan_model* m, const DevInputs& in, int batch, float* x0_out, cudaStream_t st) {
  // stage through a temporary row matrix, a few candidates at a time
  const int S = batch < 8 ? batch : 8;
  float* rows = nullptr;
  const long per = m->geom.rows_of(1) * m->CinPad;
  DAN_CUDA_TRY(cudaMalloc(&rows, (size_t)S * per * 4));
  int rc = DAN_OK;
  for (int b0 = 0; b0 < batch && rc == DAN_OK; b0 += S) {
    const int n = batch - b0 < S ? batch - b0 : S;
    rc = launch_encode_fp32(m, in, b0, n, rows, st);
    if (rc) break;
    rows_to_reference_order_kernel<<<grid_for((long)n * m->Cin * m->R * m->P), 256, 0, st>>>(
        rows, x0_out + (long)b0 * m->Cin * m->R * m->P, n, m->Cin, m->CinPad, m->geom);
    dan_count_launch();
    if (cudaGetLastError() != cudaSuccess) { dan_set_error("rows_to_reference_order launch failed"); rc = DAN_E_CUDA; }
  }
  cudaStreamSynchronize(st);
  cudaFree(rows);
  return rc;
}

int dan_fp32_forward(dan_model* m, const DevInputs& in, int batch, float* heads_out, void* ws, size_t ws_bytes, cudaStream_t st) {
  const Fp32Plan pl = make_plan(m, batch);
  if (ws_bytes < pl.total) { dan_set_error("workspace too small: %zu < %zu", ws_bytes, pl.total); return DAN_E_WORKSPACE; }
  char* base = static_cast<char*>(ws);
  const RowGeom g = m->geom;
  const int C = m->C, L = m->L, bott = m->bott;
  const long lead = (long)g.gap;   // leading zero rows
  auto rows_ptr = [&](size_t off, int ld) { return reinterpret_cast<float*>(base + off) + lead * ld; };
  float* X0 = rows_ptr(pl.off_x0, m->CinPad);
  float* H[4]; for (int i = 0; i < 4; ++i) H[i] = rows_ptr(pl.off_h[i], C);
  float* T = reinterpret_cast<float*>(base + pl.off_t);
  float* POOL = reinterpret_cast<float*>(base + pl.off_pool);
  float* HW = reinterpret_cast<float*>(base + pl.off_hw);
  float* POOLED = reinterpret_cast<float*>(base + pl.off_pooled);
  float* FCIN = reinterpret_cast<float*>(base + pl.off_fcin);
  float* PART = reinterpret_cast<float*>(base + pl.off_part);
  float* HEADS = reinterpret_cast<float*>(base + pl.off_heads);
  const int D = m->cfg.pool_combine_dimension;
  int rc;

  // zero the halo rows around every row matrix once per call (kernels rewrite all data rows, masked)
  DAN_CUDA_TRY(cudaMemsetAsync(base + pl.off_x0, 0, pl.off_pool - pl.off_x0, st));

  for (int c0 = 0; c0 < batch; c0 += pl.Bc) {
    const int nb = batch - c0 < pl.Bc ? batch - c0 : pl.Bc;
    float* pooled_dst = D > 0 ? POOLED : FCIN;
    const int pooled_ld = D > 0 ? m->pooledPad : m->fcInPad;
    if (m->fcInPad != m->fcIn || m->pooledPad != m->pooled)
      DAN_CUDA_TRY(cudaMemsetAsync(base + pl.off_pooled, 0, pl.off_fc[0] - pl.off_pooled, st));
    for (int s0 = 0; s0 < nb; s0 += pl.S) {
      const int ns = nb - s0 < pl.S ? nb - s0 : pl.S;
      const long rows = g.rows_of(ns);
      rc = launch_encode_fp32(m, in, c0 + s0, ns, X0, st);
      if (rc) return rc;
      const float* cur = X0; int ld_cur = m->CinPad;
      int hsel = 0;
      for (int l = 0; l < L; ++l) {
        const int d = m->cfg.dilation[l];
        const float* conv_in = cur;
        if (l > 0 && m->cfg.pool_after[l - 1]) {
          float* hp = H[(hsel + 2) & 3];
          add_pool_kernel<<<grid_for(rows * (C / 4)), 256, 0, st>>>(cur, POOL, hp, C, g, rows);
          dan_count_launch();
          DAN_CUDA_TRY(cudaGetLastError());
          conv_in = hp;
        }
        GemmParams p{};
        p.A = conv_in; p.lda = ld_cur; p.a_rows = pl.rowsPad + g.gap; p.M = (int)rows;
        p.ntaps = 3; p.tap_off[0] = -d; p.tap_off[1] = 0; p.tap_off[2] = d; p.Kc = (l == 0 ? m->CinPad : C);
        p.W = m->convW[l]; p.N = C; p.ldw = C; p.bias = m->convB[l]; p.relu = 1;
        if (m->cfg.use_batchnorm) { p.scale = m->bnScale[l]; p.shift = m->bnShift[l]; }
        p.mask_pitch = g.pitch; p.mask_valid = g.P; p.mask_total = rows; p.splits = 1;
        float* next;
        if (m->cfg.is_residual[l]) {
          float* y = H[(hsel + 3) & 3];
          p.out = y; p.ldo = C;
          if ((rc = launch_gemm(p, st))) return rc;
          GemmParams q{};
          next = H[(hsel + 1) & 3];
          q.A = y; q.lda = C; q.a_rows = pl.rowsPad; q.M = (int)rows; q.ntaps = 1; q.tap_off[0] = 0; q.Kc = C;
          q.W = m->resW[l]; q.N = C; q.ldw = C; q.bias = m->resB[l];
          q.resid = cur; q.ldr = ld_cur;     // residual = layer input BEFORE the pool add (model.py:732)
          q.mask_pitch = g.pitch; q.mask_valid = g.P; q.mask_total = rows; q.out = next; q.ldo = C; q.splits = 1;
          if ((rc = launch_gemm(q, st))) return rc;
        } else {
          next = H[(hsel + 1) & 3];
          p.out = next; p.ldo = C;
          if ((rc = launch_gemm(p, st))) return rc;
        }
        if (m->cfg.pool_after[l]) {
          pool_mean_kernel<<<dim3(g.P, ns), 128, 0, st>>>(next, C, POOL, C, g);
          dan_count_launch();
          DAN_CUDA_TRY(cudaGetLastError());
        }
        if (m->cfg.highway) {
          GemmParams b{};
          b.A = next; b.lda = C; b.a_rows = pl.rowsPad; b.M = (int)rows; b.ntaps = 1; b.tap_off[0] = 0; b.Kc = C;
          b.W = m->bottW[l]; b.N = bott; b.ldw = bott; b.bias = m->bottB[l]; b.relu = 1;
          b.mask_pitch = g.pitch; b.mask_valid = g.P; b.mask_total = rows; b.out = T; b.ldo = bott; b.splits = 1;
          if ((rc = launch_gemm(b, st))) return rc;
          GemmParams c{};   // (1x201) compression = one long dot product per read (model.py:776)
          c.A = T; c.lda = g.pitch * bott; c.a_rows = (long)ns * m->R; c.M = ns * m->R; c.ntaps = 1; c.tap_off[0] = 0;
          c.Kc = g.P * bott; c.W = m->compW[l]; c.N = bott; c.ldw = bott; c.bias = m->compB[l];
          float* hw_out = HW + (long)l * pl.hw_layer_stride;
          c.bias = nullptr; c.out = PART; c.ldo = bott; c.splits = kCompSplitsFp32; c.split_stride = (long)c.M * bott;
          if ((rc = launch_gemm(c, st))) return rc;
          splitk_finish_kernel<<<grid_for((long)c.M * bott), 256, 0, st>>>(PART, kCompSplitsFp32, c.split_stride, c.M, bott, bott, m->compB[l], 0, 0, hw_out, bott);
          dan_count_launch();
        }
        cur = next; ld_cur = C; hsel = (hsel + 1) & 3;
      }
      pool_final_kernel<<<dim3(g.P, ns), 128, 0, st>>>(cur, C, pooled_dst + (long)s0 * pooled_ld, pooled_ld, C, g,
                                                        m->cfg.skip_final_maxpool);
      dan_count_launch();
      DAN_CUDA_TRY(cudaGetLastError());
      if (m->cfg.highway) {
        const int base_feat = D > 0 ? D : m->pooled;
        highway_assemble_kernel<<<grid_for((long)ns * m->hwFeat), 256, 0, st>>>(
            HW, pl.hw_layer_stride, L, bott, m->R, m->cfg.concat_hw_reads, FCIN + (long)s0 * m->fcInPad + base_feat,
            m->fcInPad, ns);
        dan_count_launch();
        DAN_CUDA_TRY(cudaGetLastError());
      }
    }
    // ---- FC trunk + heads on the chunk (model.py:841-843, 917-958) ----
    if (D > 0) {
      rc = run_linear(POOLED, m->pooledPad, nb, m->pooledPad, m->postW, D, D, m->postB, 1, 0, FCIN, m->fcInPad, PART, pl.splits, st);
      if (rc) return rc;
    }
    const float* x = FCIN; int ldx = m->fcInPad; int K = m->fcInPad;
    for (int i = 0; i < m->cfg.num_fc; ++i) {
      float* y = reinterpret_cast<float*>(base + pl.off_fc[i]);
      const int N = m->cfg.fc_sizes[i];
      rc = run_linear(x, ldx, nb, K, m->fcW[i], N, N, m->fcB[i], 1, 0, y, N, PART, pl.splits, st);
      if (rc) return rc;
      x = y; ldx = N; K = N;
    }
    rc = run_linear(x, ldx, nb, K, m->headW, DAN_HEAD_PAD, DAN_HEAD_PAD, m->headB, 0, 1, HEADS, DAN_HEAD_PAD, PART, pl.splits, st);
    if (rc) return rc;
    DAN_CUDA_TRY(cudaMemcpy2DAsync(heads_out + (long)c0 * DAN_NUM_HEAD_OUTPUTS, DAN_NUM_HEAD_OUTPUTS * 4, HEADS,
                                   DAN_HEAD_PAD * 4, DAN_NUM_HEAD_OUTPUTS * 4, nb, cudaMemcpyDeviceToDevice, st));
  }
  return DAN_OK;
}

int dan_fp32_debug_fc_input(dan_model* m, int batch, const void* ws, float* out, cudaStream_t st) {
  const Fp32Plan pl = make_plan(m, batch);
  const int nb = batch % pl.Bc == 0 ? pl.Bc : batch % pl.Bc;
  const float* FCIN = reinterpret_cast<const float*>(static_cast<const char*>(ws) + pl.off_fcin);
  DAN_CUDA_TRY(cudaMemcpy2DAsync(out, (size_t)m->fcIn * 4, FCIN, (size_t)m->fcInPad * 4, (size_t)m->fcIn * 4, nb,
                                 cudaMemcpyDeviceToDevice, st));
  return nb;
}

// ----------------------------------------------------------------------------------------------------------
static int alloc_f(float** p, size_t n) {
  DAN_CUDA_TRY(cudaMalloc(p, n * sizeof(float)));
  return DAN_OK;
}

int dan_fp32_pack(dan_model* m, const dan_weights* w, cudaStream_t st) {
  const int C = m->C, L = m->L, bott = m->bott, P = m->P;
  int rc;
#define NEED(ptr, what) if (!(ptr)) { dan_set_error("missing weight tensor: %s", what); return DAN_E_INVALID; }
  NEED(w->embeddings, "embeddings"); NEED(w->pe, "pe"); NEED(w->head_w, "head_w"); NEED(w->head_b, "head_b");
  if (!m->emb) { if ((rc = alloc_f(&m->emb, DAN_VOCAB * m->cfg.embed_dim))) return rc; }
  if (!m->pe) { if ((rc = alloc_f(&m->pe, (size_t)P * m->cfg.embed_dim))) return rc; }
  DAN_CUDA_TRY(cudaMemcpyAsync(m->emb, w->embeddings, DAN_VOCAB * m->cfg.embed_dim * 4, cudaMemcpyDeviceToDevice, st));
  DAN_CUDA_TRY(cudaMemcpyAsync(m->pe, w->pe, (size_t)P * m->cfg.embed_dim * 4, cudaMemcpyDeviceToDevice, st));
  for (int l = 0; l < L; ++l) {
    const int cin = l == 0 ? m->Cin : C, cinPad = l == 0 ? m->CinPad : C;
    NEED(w->conv_w[l], "conv_w"); NEED(w->conv_b[l], "conv_b");
    if (!m->convW[l]) { if ((rc = alloc_f(&m->convW[l], (size_t)3 * cinPad * C))) return rc; if ((rc = alloc_f(&m->convB[l], C))) return rc; }
    pack_conv_w_kernel<<<grid_for((long)3 * cinPad * C), 256, 0, st>>>(w->conv_w[l], m->convW[l], C, cin, cinPad, 3);
    DAN_CUDA_TRY(cudaMemcpyAsync(m->convB[l], w->conv_b[l], C * 4, cudaMemcpyDeviceToDevice, st));
    if (m->cfg.use_batchnorm) {
      NEED(w->bn_w[l], "bn_w"); NEED(w->bn_b[l], "bn_b"); NEED(w->bn_mean[l], "bn_mean"); NEED(w->bn_var[l], "bn_var");
      if (!m->bnScale[l]) { if ((rc = alloc_f(&m->bnScale[l], C))) return rc; if ((rc = alloc_f(&m->bnShift[l], C))) return rc; }
      pack_bn_kernel<<<(C + 127) / 128, 128, 0, st>>>(w->bn_w[l], w->bn_b[l], w->bn_mean[l], w->bn_var[l], m->bnScale[l], m->bnShift[l], C);
    }
    if (m->cfg.is_residual[l]) {
      NEED(w->res_w[l], "res_w"); NEED(w->res_b[l], "res_b");
      if (!m->resW[l]) { if ((rc = alloc_f(&m->resW[l], (size_t)C * C))) return rc; if ((rc = alloc_f(&m->resB[l], C))) return rc; }
      pack_linear_w_kernel<<<grid_for((long)C * C), 256, 0, st>>>(w->res_w[l], m->resW[l], C, C, C, C);
      DAN_CUDA_TRY(cudaMemcpyAsync(m->resB[l], w->res_b[l], C * 4, cudaMemcpyDeviceToDevice, st));
    }
    if (m->cfg.highway) {
      NEED(w->bott_w[l], "bott_w"); NEED(w->bott_b[l], "bott_b"); NEED(w->comp_w[l], "comp_w"); NEED(w->comp_b[l], "comp_b");
      if (!m->bottW[l]) {
        if ((rc = alloc_f(&m->bottW[l], (size_t)C * bott))) return rc; if ((rc = alloc_f(&m->bottB[l], bott))) return rc;
        if ((rc = alloc_f(&m->compW[l], (size_t)P * bott * bott))) return rc; if ((rc = alloc_f(&m->compB[l], bott))) return rc;
      }
      pack_linear_w_kernel<<<grid_for((long)C * bott), 256, 0, st>>>(w->bott_w[l], m->bottW[l], bott, C, C, bott);
      DAN_CUDA_TRY(cudaMemcpyAsync(m->bottB[l], w->bott_b[l], bott * 4, cudaMemcpyDeviceToDevice, st));
      pack_comp_w_kernel<<<grid_for((long)P * bott * bott), 256, 0, st>>>(w->comp_w[l], m->compW[l], bott, P);
      DAN_CUDA_TRY(cudaMemcpyAsync(m->compB[l], w->comp_b[l], bott * 4, cudaMemcpyDeviceToDevice, st));
    }
  }
  const int D = m->cfg.pool_combine_dimension;
  if (D > 0) {
    NEED(w->post_pool_w, "post_pool_w"); NEED(w->post_pool_b, "post_pool_b");
    if (!m->postW) { if ((rc = alloc_f(&m->postW, (size_t)m->pooledPad * D))) return rc; if ((rc = alloc_f(&m->postB, D))) return rc; }
    pack_linear_w_kernel<<<grid_for((long)m->pooledPad * D), 256, 0, st>>>(w->post_pool_w, m->postW, D, m->pooled, m->pooledPad, D);
    DAN_CUDA_TRY(cudaMemcpyAsync(m->postB, w->post_pool_b, D * 4, cudaMemcpyDeviceToDevice, st));
  }
  int K = m->fcIn, KPad = m->fcInPad;
  for (int i = 0; i < m->cfg.num_fc; ++i) {
    const int N = m->cfg.fc_sizes[i];
    NEED(w->fc_w[i], "fc_w"); NEED(w->fc_b[i], "fc_b");
    if (!m->fcW[i]) { if ((rc = alloc_f(&m->fcW[i], (size_t)KPad * N))) return rc; if ((rc = alloc_f(&m->fcB[i], N))) return rc; }
    pack_linear_w_kernel<<<grid_for((long)KPad * N), 256, 0, st>>>(w->fc_w[i], m->fcW[i], N, K, KPad, N);
    DAN_CUDA_TRY(cudaMemcpyAsync(m->fcB[i], w->fc_b[i], N * 4, cudaMemcpyDeviceToDevice, st));
    K = N; KPad = N;
  }
  if (!m->headW) { if ((rc = alloc_f(&m->headW, (size_t)m->hidden * DAN_HEAD_PAD))) return rc; if ((rc = alloc_f(&m->headB, DAN_HEAD_PAD))) return rc; }
  pack_linear_w_kernel<<<grid_for((long)m->hidden * DAN_HEAD_PAD), 256, 0, st>>>(w->head_w, m->headW, DAN_NUM_HEAD_OUTPUTS, m->hidden, m->hidden, DAN_HEAD_PAD);
  pad_copy_kernel<<<1, 32, 0, st>>>(w->head_b, m->headB, DAN_NUM_HEAD_OUTPUTS, DAN_HEAD_PAD);
  DAN_CUDA_TRY(cudaGetLastError());
#undef NEED
  return DAN_OK;
}

void dan_fp32_free(dan_model* m) {
  auto fr = [](float*& p) { if (p) { cudaFree(p); p = nullptr; } };
  fr(m->emb); fr(m->pe); fr(m->postW); fr(m->postB); fr(m->headW); fr(m->headB);
  for (int l = 0; l < DAN_MAX_LAYERS; ++l) {
    fr(m->convW[l]); fr(m->convB[l]); fr(m->bnScale[l]); fr(m->bnShift[l]); fr(m->resW[l]); fr(m->resB[l]);
    fr(m->bottW[l]); fr(m->bottB[l]); fr(m->compW[l]); fr(m->compB[l]);
  }
  for (int i = 0; i < DAN_MAX_FC; ++i) { fr(m->fcW[i]); fr(m->fcB[i]); }
}
