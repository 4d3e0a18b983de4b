// GPU-resident preprocessing prologue of a fold: what the reference does per fold and modality on the host
// before the first minibatch --
//   RobustScaler().fit_transform(train) / .transform(test)      multimodal_kfold_train_cvae_supervised.py:101-102,
//                                                               ..._test_cvae_supervised.py:83-90
//   AGE.rank(method='first') -> pd.qcut(q=27) -> np.eye(27)[.]  ..._train_cvae_supervised.py:105-114 (q=2 for PTGENDER)
//   the bootstrap / merge row selection                         utils.py:73-93, 112-168 (row positions come from the host:
//                                                               they are integer work on the legacy numpy RNG stream)
//   torch.cat((x, c), dim=1)                                    cVAE.py:163
// -- as three kernels over the raw float64 feature table resident in HBM: per-column median / IQR of the selected rows
// (shared-memory bitonic sort, numpy's 'linear' percentile arithmetic in float64), stable rank -> quantile bin, and one
// fused gather + scale + one-hot + pack pass that writes the packed fp32 rows the training kernels consume.  float64
// in, float64 arithmetic, one rounding to fp32 at the end: bit-identical to sklearn + astype(float32) on the host.
#include "nmb_internal.h"

namespace nmb {

constexpr int kProMaxRows = 8192;       // rows per fold the shared-memory sort holds (64 KB of doubles)

// numpy.percentile(method='linear') on a sorted array: virtual index (n - 1) q, numpy's two-sided lerp
__device__ inline double np_percentile_sorted(const double* a, int n, double q) {
  const double vi = (double)n * q + (1.0 + q * (1.0 - 1.0 - 1.0)) - 1.0;       // _compute_virtual_index, alpha = beta = 1
  int lo = (int)floor(vi);
  if (lo < 0) lo = 0;
  if (lo > n - 1) lo = n - 1;
  const int hi = lo + 1 > n - 1 ? n - 1 : lo + 1;
  const double t = vi - floor(vi);
  const double x = a[lo], y = a[hi], d = y - x;
  // _lerp: a + (b - a) t, and from the other end for t >= 0.5.  Separate multiply and add like numpy (the explicit
  // round-to-nearest intrinsics are never contracted into an FMA, which would round differently).
  double r = __dadd_rn(x, __dmul_rn(d, t));
  if (t >= 0.5) r = __dsub_rn(y, __dmul_rn(d, __dsub_rn(1.0, t)));
  if (t == 1.0) r = y;                                        // (numpy's special cases; unreachable for t in [0, 1))
  if (d == 0.0) r = x;
  return r;
}

// CTA = one column.  x[idx[i] * ld + col] for i < n -> sorted in shared memory -> center = median, scale = q75 - q25.
__global__ void __launch_bounds__(256) robust_fit_kernel(const double* __restrict__ x, long long ld, int d,
                                                         const int* __restrict__ idx, int n, double* center, double* scale) {
  extern __shared__ double col[];
  const int c = blockIdx.x;
  int p2 = 1;
  while (p2 < n) p2 <<= 1;
  for (int i = threadIdx.x; i < p2; i += blockDim.x)
    col[i] = i < n ? x[(long long)(idx ? idx[i] : i) * ld + c] : __longlong_as_double(0x7ff0000000000000LL);   // +inf pad
  __syncthreads();
  for (int k = 2; k <= p2; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < p2; i += blockDim.x) {
        const int p = i ^ j;
        if (p > i) {
          const double a = col[i], b = col[p];
          const bool up = (i & k) == 0;
          if ((a > b) == up) { col[i] = b; col[p] = a; }
        }
      }
      __syncthreads();
    }
  if (threadIdx.x == 0) {
    // np.nanmedian: mean of the two middle elements for even n (np.mean -> (a + b) / 2 in float64)
    const double med = (n & 1) ? col[n / 2] : (col[n / 2 - 1] + col[n / 2]) / 2.0;
    const double q25 = np_percentile_sorted(col, n, 0.25), q75 = np_percentile_sorted(col, n, 0.75);
    double s = q75 - q25;
    if (s < 10.0 * 2.220446049250313e-16) s = 1.0;            // sklearn _handle_zeros_in_scale
    center[c] = med; scale[c] = s;
  }
}

// rank(method='first') (1-based, ties by order of appearance) -> pandas qcut bin against host-computed edges
// (right-closed intervals, lowest edge included): thread per row, O(n^2) comparisons -- n is a fold's row count.
__global__ void __launch_bounds__(256) rank_bins_kernel(const double* __restrict__ v, const int* __restrict__ idx, int n,
                                                        const double* __restrict__ edges, int q, int* bins) {
  extern __shared__ double vals[];
  for (int i = threadIdx.x; i < n; i += blockDim.x) vals[i] = v[idx ? idx[i] : i];
  __syncthreads();
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const double mine = vals[i];
    int rank = 1;
    for (int j = 0; j < n; ++j) { const double o = vals[j]; rank += (o < mine) || (o == mine && j < i); }
    const double r = (double)rank;
    int lo = 0, hi = q + 1;                                    // searchsorted(edges, r, side='left')
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (edges[mid] < r) lo = mid + 1; else hi = mid; }
    int b = lo - 1;
    if (r == edges[0]) b = 0;
    bins[i] = b < 0 ? 0 : (b > q - 1 ? q - 1 : b);
  }
}

// out[i] = [ fp32((x[idx[i]][j] - center[j]) / scale[j]) | one-hot(age_bin[i], n_age) | one-hot(sex_bin[i], n_sex) | 1 | 0 ]
__global__ void pack_scaled_kernel(const double* __restrict__ x, long long ld, int d, const int* __restrict__ idx,
                                   long long n, const double* __restrict__ center, const double* __restrict__ scale,
                                   const int* __restrict__ age_bin, int n_age, const int* __restrict__ sex_bin, int n_sex,
                                   int ldx, float* __restrict__ out) {
  const long long total = n * ldx;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long i = e / ldx;
    const int j = (int)(e - i * ldx);
    float val = 0.f;
    if (j < d) val = (float)((x[(long long)(idx ? idx[i] : i) * ld + j] - center[j]) / scale[j]);
    else if (j < d + n_age) val = (j - d) == age_bin[i] ? 1.f : 0.f;
    else if (j < d + n_age + n_sex) val = (j - d - n_age) == sex_bin[i] ? 1.f : 0.f;
    else if (j == d + n_age + n_sex) val = 1.f;
    out[e] = val;
  }
}

int prologue_max_rows() { return kProMaxRows; }

cudaError_t launch_robust_fit(const double* x, long long ld, int d, const int* idx, int n, double* center, double* scale,
                              cudaStream_t st) {
  int p2 = 1;
  while (p2 < n) p2 <<= 1;
  const int smem = p2 * (int)sizeof(double);
  cudaError_t e = cudaFuncSetAttribute(robust_fit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  if (e != cudaSuccess) return e;
  robust_fit_kernel<<<d, 256, smem, st>>>(x, ld, d, idx, n, center, scale);
  return cudaGetLastError();
}

cudaError_t launch_rank_bins(const double* v, const int* idx, int n, const double* edges, int q, int* bins, cudaStream_t st) {
  cudaError_t e = cudaFuncSetAttribute(rank_bins_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  if (e != cudaSuccess) return e;
  const int blocks = (n + 255) / 256;
  rank_bins_kernel<<<blocks, 256, n * (int)sizeof(double), st>>>(v, idx, n, edges, q, bins);
  return cudaGetLastError();
}

cudaError_t launch_pack_scaled(const double* x, long long ld, int d, const int* idx, long long n, const double* center,
                               const double* scale, const int* age_bin, int n_age, const int* sex_bin, int n_sex, int ldx,
                               float* out, cudaStream_t st) {
  const long long total = n * ldx;
  if (total == 0) return cudaSuccess;
  const int blocks = (int)min((total + 255) / 256, (long long)sm_count() * 16);
  pack_scaled_kernel<<<blocks, 256, 0, st>>>(x, ld, d, idx, n, center, scale, age_bin, n_age, sex_bin, n_sex, ldx, out);
  return cudaGetLastError();
}

}  // namespace nmb
