// Group normalisation + ReLU for the non-affine decoder residual blocks
// (CAPE.gn at lib/models.py:681-712 followed by tf.nn.relu at :752,756,760).
// Layout [N, rows, C]: the statistics of group g of sample n span (C/G contiguous channels) x (all rows),
// biased variance (tf.nn.moments), eps inside the sqrt.  Group sums are accumulated in fp64 so the
// E[x^2]-E[x]^2 form stays within fp32 rounding of the two-pass reference.
#include "common.cuh"

namespace cape {

constexpr int GN_ROWS = 32;      // rows per CTA in the reduction passes
constexpr int GN_MAXG = 32;

__global__ void __launch_bounds__(256) gn_stats_kernel(const float* __restrict__ x, int rows, int C, int G,
                                                       double* __restrict__ acc) {
  __shared__ double gs[GN_MAXG][2];
  const int n = blockIdx.y;
  const int r0 = blockIdx.x * GN_ROWS, r1 = min(rows, r0 + GN_ROWS);
  const int cpg = C / G;
  if (threadIdx.x < GN_MAXG) { gs[threadIdx.x][0] = 0.0; gs[threadIdx.x][1] = 0.0; }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 0.f, q = 0.f;
    const float* xp = x + ((size_t)n * rows) * C + c;
    for (int r = r0; r < r1; ++r) {
      const float v = __ldg(xp + (size_t)r * C);
      s += v; q = fmaf(v, v, q);
    }
    atomicAdd(&gs[c / cpg][0], (double)s);
    atomicAdd(&gs[c / cpg][1], (double)q);
  }
  __syncthreads();
  if (threadIdx.x < G) {
    atomicAdd(acc + ((size_t)n * G + threadIdx.x) * 2 + 0, gs[threadIdx.x][0]);
    atomicAdd(acc + ((size_t)n * G + threadIdx.x) * 2 + 1, gs[threadIdx.x][1]);
  }
}

__global__ void gn_finalize_kernel(const double* __restrict__ acc, int total, double inv_cnt, float eps,
                                   float* __restrict__ stats) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const double mean = acc[2 * i] * inv_cnt;
  double var = acc[2 * i + 1] * inv_cnt - mean * mean;
  if (var < 0.0) var = 0.0;
  stats[2 * i] = (float)mean;
  stats[2 * i + 1] = (float)(1.0 / sqrt(var + (double)eps));
}

__global__ void gn_apply_kernel(const float* __restrict__ x, long long total, int rows, int C, int G,
                                const float* __restrict__ gamma, const float* __restrict__ beta,
                                const float* __restrict__ stats, float* __restrict__ y) {
  const int cpg = C / G;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const int n = (int)(i / ((long long)rows * C));
    const float* st = stats + ((size_t)n * G + c / cpg) * 2;
    const float v = (x[i] - st[0]) * st[1] * __ldg(gamma + c) + __ldg(beta + c);
    y[i] = fmaxf(v, 0.f);
  }
}

// pass 1 of backward: per-channel sums (dgamma, dbeta) and per-(n,g) sums of dxhat and dxhat*xhat
__global__ void __launch_bounds__(256) gn_bwd_stats_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                                           const float* __restrict__ dy, int rows, int C, int G,
                                                           const float* __restrict__ gamma,
                                                           const float* __restrict__ stats, float* __restrict__ dgamma,
                                                           float* __restrict__ dbeta, double* __restrict__ acc) {
  __shared__ double gs[GN_MAXG][2];
  const int n = blockIdx.y;
  const int r0 = blockIdx.x * GN_ROWS, r1 = min(rows, r0 + GN_ROWS);
  const int cpg = C / G;
  if (threadIdx.x < GN_MAXG) { gs[threadIdx.x][0] = 0.0; gs[threadIdx.x][1] = 0.0; }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const int g = c / cpg;
    const float mean = stats[((size_t)n * G + g) * 2], rstd = stats[((size_t)n * G + g) * 2 + 1];
    float sb = 0.f, sg = 0.f;
    const size_t off = ((size_t)n * rows) * C + c;
    for (int r = r0; r < r1; ++r) {
      const size_t e = off + (size_t)r * C;
      const float gy = (y[e] > 0.f) ? dy[e] : 0.f;
      const float xh = (x[e] - mean) * rstd;
      sb += gy; sg = fmaf(gy, xh, sg);
    }
    atomicAdd(dbeta + c, sb);
    atomicAdd(dgamma + c, sg);
    const float gm = __ldg(gamma + c);
    atomicAdd(&gs[g][0], (double)(gm * sb));
    atomicAdd(&gs[g][1], (double)(gm * sg));
  }
  __syncthreads();
  if (threadIdx.x < G) {
    atomicAdd(acc + ((size_t)n * G + threadIdx.x) * 2 + 0, gs[threadIdx.x][0]);
    atomicAdd(acc + ((size_t)n * G + threadIdx.x) * 2 + 1, gs[threadIdx.x][1]);
  }
}

__global__ void gn_bwd_apply_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                    const float* __restrict__ dy, long long total, int rows, int C, int G,
                                    const float* __restrict__ gamma, const float* __restrict__ stats,
                                    const double* __restrict__ acc, double inv_cnt, float* __restrict__ dx,
                                    int accumulate) {
  const int cpg = C / G;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const int n = (int)(i / ((long long)rows * C));
    const size_t sg = (size_t)n * G + c / cpg;
    const float mean = stats[sg * 2], rstd = stats[sg * 2 + 1];
    const float m1 = (float)(acc[sg * 2] * inv_cnt), m2 = (float)(acc[sg * 2 + 1] * inv_cnt);
    const float gy = (y[i] > 0.f) ? dy[i] : 0.f;
    const float xh = (x[i] - mean) * rstd;
    const float d = rstd * (__ldg(gamma + c) * gy - m1 - xh * m2);
    dx[i] = accumulate ? dx[i] + d : d;
  }
}

}  // namespace cape

using namespace cape;

static int gn_check(cape_topology* t, int N, int rows, int C, int G) {
  CAPE_REQUIRE(t, "null handle");
  CAPE_REQUIRE(N > 0 && rows > 0 && C > 0 && G > 0 && G <= GN_MAXG && C % G == 0, "bad group-norm shape");
  CAPE_REQUIRE((int64_t)N * G * 2 * (int64_t)sizeof(double) <= t->workspace_bytes, "workspace too small for group norm");
  CAPE_REQUIRE(N <= 65535, "batch too large");
  return 0;
}

extern "C" int cape_gn_relu_fwd(cape_topology* t, const float* x, int N, int rows, int C, int G, float eps,
                                const float* gamma, const float* beta, float* y, float* stats, void* stream) {
  if (gn_check(t, N, rows, C, G) != 0) return -1;
  CAPE_REQUIRE(x && gamma && beta && y && stats, "null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  double* acc = (double*)t->workspace;
  CAPE_CHECK_CUDA(cudaMemsetAsync(acc, 0, (size_t)N * G * 2 * sizeof(double), st));
  dim3 grid((rows + GN_ROWS - 1) / GN_ROWS, N);
  gn_stats_kernel<<<grid, 256, 0, st>>>(x, rows, C, G, acc);
  CAPE_CHECK_CUDA(cudaGetLastError());
  cape::count_launches(1);
  const double inv_cnt = 1.0 / ((double)rows * (C / G));
  gn_finalize_kernel<<<(N * G + 127) / 128, 128, 0, st>>>(acc, N * G, inv_cnt, eps, stats);
  CAPE_CHECK_CUDA(cudaGetLastError());
  cape::count_launches(1);
  const long long total = (long long)N * rows * C;
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  gn_apply_kernel<<<(unsigned)blocks, 256, 0, st>>>(x, total, rows, C, G, gamma, beta, stats, y);
  CAPE_CHECK_CUDA(cudaGetLastError());
  cape::count_launches(1);
  return 0;
}

extern "C" int cape_gn_relu_bwd(cape_topology* t, const float* x, const float* y, const float* dy, int N, int rows,
                                int C, int G, const float* gamma, const float* stats, float* dx, int accumulate_dx,
                                float* dgamma, float* dbeta, void* stream) {
  if (gn_check(t, N, rows, C, G) != 0) return -1;
  CAPE_REQUIRE(x && y && dy && gamma && stats && dx && dgamma && dbeta, "null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  double* acc = (double*)t->workspace;
  CAPE_CHECK_CUDA(cudaMemsetAsync(acc, 0, (size_t)N * G * 2 * sizeof(double), st));
  dim3 grid((rows + GN_ROWS - 1) / GN_ROWS, N);
  gn_bwd_stats_kernel<<<grid, 256, 0, st>>>(x, y, dy, rows, C, G, gamma, stats, dgamma, dbeta, acc);
  CAPE_CHECK_CUDA(cudaGetLastError());
  cape::count_launches(1);
  const double inv_cnt = 1.0 / ((double)rows * (C / G));
  const long long total = (long long)N * rows * C;
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  gn_bwd_apply_kernel<<<(unsigned)blocks, 256, 0, st>>>(x, y, dy, total, rows, C, G, gamma, stats, acc, inv_cnt, dx, accumulate_dx);
  CAPE_CHECK_CUDA(cudaGetLastError());
  cape::count_launches(1);
  return 0;
}
