// Kernels for the "thin-input" layers: sources with <= 4 channels (the 3-channel mesh offsets entering the encoder
// and the discriminator, the 3-channel output gradient entering the decoder's backward pass, the 1-channel logit
// gradient).  These layers have almost no arithmetic (K*Fin <= 32 multiply-adds per output), so they are pure
// HBM problems: read a few MB, write/read the wide [N, rows, 32..128] tensor once.  The generic tiled kernels
// spend their time on padded reductions and scalar gathers here; these two do the Chebyshev basis on the thin
// side in registers/shared memory and stream the wide side with fully coalesced accesses.
//   thin_fwd_kernel : cape_cheb_fwd when every term has F <= 4   (lib/models.py:69-109 for enc/disc conv1, and the
//                     data-gradient of the decoder's output conv / the discriminator's prediction map)
//   thin_dw_kernel  : cape_cheb_dw  when F <= 4                  (weight gradient of the same layers)
#include "common.cuh"
#include "ellconv_params.cuh"

namespace cape {

namespace {

constexpr int TH_ROWS = 128;      // output rows per block (forward)
constexpr int TH_MAXKF = 32;      // total thin channels over all terms
constexpr int TH_MAXCOLS = 128;
constexpr int TH_QS = 1024;       // condition vectors: samples-in-tile x slots x ncols

// thin gather: out[f] = sum_j w[r,j] * src[n, idx[r,j], f]  for f < F (F <= 4), four taps per table fetch
__device__ __forceinline__ void thin_gather(const OpView& op, int r, const float* base, size_t stride, int F,
                                            float (&v)[4]) {
  v[0] = v[1] = v[2] = v[3] = 0.f;
  if (op.idx == nullptr) {
    const float* s = base + (size_t)r * stride;
#pragma unroll
    for (int f = 0; f < 4; ++f)
      if (f < F) v[f] = __ldg(s + f);
    return;
  }
  const int4* ip = reinterpret_cast<const int4*>(op.idx + (size_t)r * op.width);
  const float4* wp = reinterpret_cast<const float4*>(op.w + (size_t)r * op.width);
  const int nb = op.width >> 2;
  for (int b = 0; b < nb; ++b) {
    const int4 id = __ldg(ip + b);
    if (id.x < 0) break;
    const float4 ww = __ldg(wp + b);
    const int ids[4] = {id.x, max(id.y, 0), max(id.z, 0), max(id.w, 0)};
    const float ws[4] = {ww.x, ww.y, ww.z, ww.w};
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float* s = base + (size_t)ids[u] * stride;
#pragma unroll
      for (int f = 0; f < 4; ++f)
        if (f < F) v[f] = fmaf(ws[u], __ldg(s + f), v[f]);
    }
  }
}

__global__ void __launch_bounds__(256) thin_fwd_kernel(const __grid_constant__ ConvParams p, int KF) {
  __shared__ float Bs[TH_ROWS][TH_MAXKF + 1];
  __shared__ __align__(16) float Ws[TH_MAXKF * TH_MAXCOLS];
  __shared__ float qs[TH_QS];
  __shared__ int s_n[TH_ROWS], s_r[TH_ROWS];
  __shared__ int s_off[CAPE_MAX_TERMS];

  const int tid = threadIdx.x;
  const long long row0 = (long long)blockIdx.x * TH_ROWS;
  const int ncols = p.ncols;
  if (tid < TH_ROWS) {
    const long long R = row0 + tid;
    if (R < p.total_rows) { s_n[tid] = (int)(R / p.rows_out); s_r[tid] = (int)(R % p.rows_out); }
    else { s_n[tid] = -1; s_r[tid] = 0; }
  }
  if (tid == 0) {
    int o = 0;
    for (int t = 0; t < p.nterms; ++t) { s_off[t] = o; o += p.terms[t].F; }
  }
  // weights: Ws[q][c] with q running over (term, f)
  for (int e = tid; e < KF * ncols; e += 256) {
    const int q = e / ncols, c = e % ncols;
    int t = 0, o = 0;
    while (t + 1 < p.nterms && q >= o + p.terms[t].F) { o += p.terms[t].F; ++t; }
    Ws[e] = __ldg(p.terms[t].w + (size_t)(q - o) * p.terms[t].w_stride + c);
  }
  __syncthreads();

  // ---- phase 1: Chebyshev basis of the thin source, one (row, term) item per thread iteration
  for (int item = tid; item < TH_ROWS * p.nterms; item += 256) {
    const int row = item % TH_ROWS, t = item / TH_ROWS;
    const TermDev& tm = p.terms[t];
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    const int n = s_n[row];
    if (n >= 0) thin_gather(tm.op, s_r[row], tm.src + (size_t)n * tm.src_rows * tm.src_stride, (size_t)tm.src_stride, tm.F, v);
    const int o = s_off[t];
#pragma unroll
    for (int f = 0; f < 4; ++f)
      if (f < tm.F) Bs[row][o + f] = v[f];
  }
  // ---- condition broadcast vectors
  const int n_first = s_n[0];
  if (p.nslots > 0) {
    int n_last = n_first;
    for (int i = TH_ROWS - 1; i > 0; --i)
      if (s_n[i] >= 0) { n_last = s_n[i]; break; }
    const int S = n_last - n_first + 1;
    for (int o = tid; o < S * p.nslots * ncols; o += 256) {
      const int c = o % ncols, slot = (o / ncols) % p.nslots, s = o / (ncols * p.nslots);
      const float* y = p.cond + (size_t)(n_first + s) * p.C;
      const float* wc = p.slot_w[slot] + c;
      const int ws = p.terms[p.slot_term[slot]].w_stride;
      float q = 0.f;
      for (int j = 0; j < p.C; ++j) q = fmaf(__ldg(y + j), __ldg(wc + (size_t)j * ws), q);
      qs[o] = q;
    }
  }
  __syncthreads();

  // ---- phase 2: outputs, coalesced along the columns
  for (int o = tid; o < TH_ROWS * ncols; o += 256) {
    const int row = o / ncols, c = o % ncols;
    const int n = s_n[row];
    if (n < 0) continue;
    const int r = s_r[row];
    float acc = 0.f;
    for (int q = 0; q < KF; ++q) acc = fmaf(Bs[row][q], Ws[q * ncols + c], acc);
    for (int slot = 0; slot < p.nslots; ++slot) {
      const TermDev& tm = p.terms[p.slot_term[slot]];
      const float coef = tm.op.rowsum ? __ldg(tm.op.rowsum + r) : 1.f;
      acc = fmaf(coef, qs[((n - n_first) * p.nslots + slot) * ncols + c], acc);
    }
    const size_t oi = (size_t)(row0 + row) * ncols + c;
    if (p.epilogue == CAPE_EPI_LINEAR) {
      if (p.bias != nullptr) acc += __ldg(p.bias + (p.bias_per_row ? (size_t)r * ncols : 0) + c);
      if (p.act == CAPE_ACT_LEAKY) acc = acc > 0.f ? acc : p.alpha * acc;
      else if (p.act == CAPE_ACT_RELU) acc = fmaxf(acc, 0.f);
      p.out[oi] = acc;
    } else if (p.epilogue == CAPE_EPI_SLOPE) {
      p.out[oi] = acc * (__ldg(p.aux + oi) > 0.f ? 1.f : p.alpha);
    } else {  // DUALMASK
      p.out[oi] = acc;
      if (p.out2 != nullptr) p.out2[oi] = __ldg(p.aux + oi) > 0.f ? acc : 0.f;
    }
  }
}

struct ThinDwParams {
  int rows_out, ncols, F, src_rows, src_stride;
  long long total_rows, rows_per_block;
  const float* src;
  OpView op;
  const float* g;
  float* out;      // partial sums [gridDim.x, F, ncols]
};

constexpr int TD_CHUNK = 256;

__global__ void __launch_bounds__(256) thin_dw_kernel(const __grid_constant__ ThinDwParams p) {
  __shared__ float Bs[TD_CHUNK][4];
  __shared__ float red[4][256];
  const int tid = threadIdx.x;
  const int ncols = p.ncols;
  const int RG = 256 / ncols;                 // row groups (ncols in {32, 64, 128, 256})
  const int c = tid % ncols, rg = tid / ncols;
  const long long rbeg = (long long)blockIdx.x * p.rows_per_block;
  const long long rend = min(p.total_rows, rbeg + p.rows_per_block);
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (long long rb = rbeg; rb < rend; rb += TD_CHUNK) {
    // basis of 256 rows, one row per thread
    {
      const long long R = rb + tid;
      float v[4] = {0.f, 0.f, 0.f, 0.f};
      if (R < rend) {
        const int n = (int)(R / p.rows_out), r = (int)(R % p.rows_out);
        thin_gather(p.op, r, p.src + (size_t)n * p.src_rows * p.src_stride, (size_t)p.src_stride, p.F, v);
      }
      Bs[tid][0] = v[0]; Bs[tid][1] = v[1]; Bs[tid][2] = v[2]; Bs[tid][3] = v[3];
    }
    __syncthreads();
    const int lim = (int)min((long long)TD_CHUNK, rend - rb);
    for (int r = rg; r < lim; r += RG) {
      const float gv = __ldg(p.g + (size_t)(rb + r) * ncols + c);
      acc[0] = fmaf(Bs[r][0], gv, acc[0]); acc[1] = fmaf(Bs[r][1], gv, acc[1]);
      acc[2] = fmaf(Bs[r][2], gv, acc[2]); acc[3] = fmaf(Bs[r][3], gv, acc[3]);
    }
    __syncthreads();
  }
  // reduce the row groups (fixed order) and write this block's partial sums
#pragma unroll
  for (int f = 0; f < 4; ++f) red[f][tid] = acc[f];
  __syncthreads();
  if (rg == 0) {
    for (int f = 0; f < p.F; ++f) {
      float s = 0.f;
      for (int k = 0; k < RG; ++k) s += red[f][k * ncols + c];
      p.out[((size_t)blockIdx.x * p.F + f) * ncols + c] = s;
    }
  }
}

}  // namespace

// returns 1 if launched, 0 if not eligible
int launch_thin_fwd(const cape_topology* t, const ConvParams& p, bool dual, cudaStream_t st) {
  (void)t;
  if (dual || p.epilogue == CAPE_EPI_AFFINE) return 0;
  if (p.ncols > TH_MAXCOLS || p.ncols < 16) return 0;
  int KF = 0;
  for (int i = 0; i < p.nterms; ++i) {
    if (p.terms[i].F > 4 || p.terms[i].stash != nullptr) return 0;
    KF += p.terms[i].F;
  }
  if (KF > TH_MAXKF) return 0;
  if (p.nslots > 0) {
    const long long max_samples = (TH_ROWS - 1) / p.rows_out + 2;
    if (max_samples * p.nslots * p.ncols > TH_QS) return 0;
    for (int s = 0; s < p.nslots; ++s)
      if (p.slot_acc[s] != 0) return 0;
  }
  const unsigned grid = (unsigned)((p.total_rows + TH_ROWS - 1) / TH_ROWS);
  thin_fwd_kernel<<<grid, 256, 0, st>>>(p, KF);
  CAPE_CHECK_CUDA(cudaGetLastError());
  count_launches(1);
  return 1;
}

// returns 1 if launched (partials [*nsplit_out, F, ncols] in the workspace), 0 if not eligible
int launch_thin_dw(const cape_topology* t, const cape_dw_args* a, const OpView& op, int* nsplit_out, cudaStream_t st) {
  if (a->F > 4) return 0;
  if (!(a->ncols == 32 || a->ncols == 64 || a->ncols == 128 || a->ncols == 256)) return 0;
  ThinDwParams p{};
  p.rows_out = a->rows_out; p.ncols = a->ncols; p.F = a->F; p.src_rows = a->src_rows; p.src_stride = a->src_stride;
  p.total_rows = (long long)a->N * a->rows_out;
  p.src = a->src; p.op = op; p.g = a->g;
  long long nblk = 4LL * t->sm_count;
  const long long max_by_rows = (p.total_rows + TD_CHUNK - 1) / TD_CHUNK;
  if (nblk > max_by_rows) nblk = max_by_rows;
  const long long per = (long long)a->F * a->ncols * (long long)sizeof(float);
  if (nblk * per > t->workspace_bytes) nblk = t->workspace_bytes / per;
  if (nblk < 1) return 0;
  long long rpb = (p.total_rows + nblk - 1) / nblk;
  rpb = (rpb + TD_CHUNK - 1) / TD_CHUNK * TD_CHUNK;
  nblk = (p.total_rows + rpb - 1) / rpb;
  p.rows_per_block = rpb;
  p.out = (float*)t->workspace;
  thin_dw_kernel<<<(unsigned)nblk, 256, 0, st>>>(p);
  CAPE_CHECK_CUDA(cudaGetLastError());
  count_launches(1);
  *nsplit_out = (int)nblk;
  return 1;
}

}  // namespace cape
