// Streaming (HBM-bound) kernels of the scoring half of the hot path: row packing,
// per-ROI normative statistics, deviation / z-score tables, ROC-AUC by exact pair counting.
//
// Reference behaviour replaced:
//   (x - xhat)^2 per ROI                 multimodal_kfold_test_cvae_supervised.py:141
//   sum_d (x - xhat)^2 / D per subject   cVAE.py:1210-1211, utils_vae.py:147-148
//   HC-referenced z-score                utils_vae.py:155-161 (latent space) -> ROI space,
//                                        definition in oracle/deviation.py
//   roc_curve + auc                      multimodal_kfold_cvae_group_analysis_1x1.py:123-124
#include "nmb_internal.h"

namespace nmb {

// SMs of the current device (queried once per device; grids are sized in multiples of it)
int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n < 1) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

// ---- pack rows: [x | c | 1 | 0] ------------------------------------------------------------
__global__ void pack_rows_kernel(const float* __restrict__ x, const float* __restrict__ c, long long n_rows,
                                 int d, int c_dim, int ldx, float* __restrict__ out) {
  const long long total = n_rows * ldx;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / ldx;
    const int j = (int)(i - r * ldx);
    float v = 0.f;
    if (j < d) v = x[r * d + j];
    else if (j < d + c_dim) v = c[r * c_dim + (j - d)];
    else if (j == d + c_dim) v = 1.f;
    out[i] = v;
  }
}

void launch_pack_rows(const float* x, const float* c, long long n_rows, int d, int c_dim, int ldx, float* out,
                      cudaStream_t st) {
  const long long total = n_rows * ldx;
  if (total == 0) return;
  const int blocks = (int)min((total + 255) / 256, (long long)sm_count() * 16);
  pack_rows_kernel<<<blocks, 256, 0, st>>>(x, c, n_rows, d, c_dim, ldx, out);
}

// ---- normative statistics -------------------------------------------------------------------
// grid = (ceil(max_d/32), n_seg); block = 32 columns x 8 row groups.  Consecutive lanes read
// consecutive ROIs (coalesced 128 B rows); sums are accumulated in fp64 and combined across the
// 8 row groups in a fixed order, so the result does not depend on the launch geometry.
__global__ void __launch_bounds__(256) stats_kernel(SegTable t) {
  const int s = blockIdx.y;
  const int d = t.d[s];
  const int col = blockIdx.x * 32 + (threadIdx.x & 31);
  const int rg = threadIdx.x >> 5;
  const float* x = t.x[s];
  const float* xh = t.xhat[s];
  const uint8_t* mask = t.mask ? t.mask[s] : nullptr;
  const int n = t.n_rows[s], ldx = t.ldx[s];
  double sum = 0.0, sq = 0.0;
  int cnt = 0;
  if (col < d) {
    // four rows of this thread's row group per pass: the mask bytes and the eight loads are issued before the first use
    // (one row per pass kept ~2 KB of reads in flight per CTA); accumulation stays in row order, so the sums are unchanged
    for (int i0 = rg; i0 < n; i0 += 32) {
      float xv[4], hv[4];
      bool ok[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + 8 * u;
        ok[u] = i < n && (!mask || mask[i]);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + 8 * u;
        xv[u] = ok[u] ? x[(long long)i * ldx + col] : 0.f;
        hv[u] = ok[u] ? xh[(long long)i * d + col] : 0.f;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (!ok[u]) continue;
        const float r = xv[u] - hv[u];
        const double r2 = (double)(r * r);
        sum += r2; sq += r2 * r2; ++cnt;
      }
    }
  }
  __shared__ double s_sum[8][32], s_sq[8][32];
  __shared__ int s_cnt[8][32];
  s_sum[rg][threadIdx.x & 31] = sum; s_sq[rg][threadIdx.x & 31] = sq; s_cnt[rg][threadIdx.x & 31] = cnt;
  __syncthreads();
  if (rg == 0 && col < d) {
    double a = 0.0, b = 0.0; int c = 0;
    for (int g = 0; g < 8; ++g) { a += s_sum[g][threadIdx.x]; b += s_sq[g][threadIdx.x]; c += s_cnt[g][threadIdx.x]; }
    float* out = t.stats_out[s];
    if (c > 0) {
      const double mean = a / c;
      double var = b / c - mean * mean;
      if (var < 0.0) var = 0.0;
      out[col] = (float)mean; out[d + col] = (float)sqrt(var);
    } else {
      out[col] = nanf(""); out[d + col] = nanf("");
    }
  }
}

void launch_stats(const SegTable& t, int n_seg, int max_d, cudaStream_t st) {
  if (n_seg == 0 || max_d == 0) return;
  dim3 grid((max_d + 31) / 32, n_seg);
  stats_kernel<<<grid, 256, 0, st>>>(t);
}

// ---- deviation tables ------------------------------------------------------------------------
// One warp per subject row; lanes stride over ROIs (float4 when the row is 16 B aligned),
// warp-shuffle reduction for the per-subject mean.  grid = (row blocks, n_seg).
__global__ void __launch_bounds__(256) deviation_kernel(SegTable t) {
  const int s = blockIdx.y;
  const int d = t.d[s], n = t.n_rows[s], ldx = t.ldx[s];
  const float* x = t.x[s];
  const float* xh = t.xhat[s];
  const float* stats = t.stats ? t.stats[s] : nullptr;
  float* roi = t.dev_roi ? t.dev_roi[s] : nullptr;
  float* zs = (t.z && stats) ? t.z[s] : nullptr;
  float* subj = t.dev_subj ? t.dev_subj[s] : nullptr;
  const int lane = threadIdx.x & 31;
  const int warps = (blockDim.x >> 5) * gridDim.x;
  const bool vec = (d & 3) == 0;
  int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (vec && d <= 128) {
    // the usual widths (<= 128 ROIs: one float4 per lane and row): two rows per warp pass, all four loads of a lane in
    // flight before the first use, so that a warp keeps ~2 KB of reads outstanding instead of ~1 KB
    const int j = lane * 4;
    const bool on = j < d;
    float4 mu = make_float4(0.f, 0.f, 0.f, 0.f), sd = make_float4(1.f, 1.f, 1.f, 1.f);
    if (zs && on) { mu = *reinterpret_cast<const float4*>(stats + j); sd = *reinterpret_cast<const float4*>(stats + d + j); }
    for (; i < n; i += 2 * warps) {
      const int i2 = i + warps;
      const bool two = i2 < n;
      float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), b0 = a0, a1 = a0, b1 = a0;
      if (on) {
        a0 = *reinterpret_cast<const float4*>(x + (long long)i * ldx + j);
        b0 = *reinterpret_cast<const float4*>(xh + (long long)i * d + j);
        if (two) {
          a1 = *reinterpret_cast<const float4*>(x + (long long)i2 * ldx + j);
          b1 = *reinterpret_cast<const float4*>(xh + (long long)i2 * d + j);
        }
      }
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        if (k == 1 && !two) break;
        const float4 a = k ? a1 : a0, b = k ? b1 : b0;
        const long long row = k ? i2 : i;
        float4 r = make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w);
        r.x *= r.x; r.y *= r.y; r.z *= r.z; r.w *= r.w;
        float acc = (r.x + r.y) + (r.z + r.w);
        if (on) {
          if (roi) *reinterpret_cast<float4*>(roi + row * d + j) = r;
          if (zs) *reinterpret_cast<float4*>(zs + row * d + j) =
              make_float4((r.x - mu.x) / sd.x, (r.y - mu.y) / sd.y, (r.z - mu.z) / sd.z, (r.w - mu.w) / sd.w);
        }
        acc = warp_sum(acc);
        if (subj && lane == 0) subj[row] = acc / d;
      }
    }
    return;
  }
  for (; i < n; i += warps) {
    const float* xr = x + (long long)i * ldx;
    const float* hr = xh + (long long)i * d;
    float acc = 0.f;
    if (vec) {
      for (int j = lane * 4; j < d; j += 128) {
        const float4 a = *reinterpret_cast<const float4*>(xr + j);
        const float4 b = *reinterpret_cast<const float4*>(hr + j);
        float4 r = make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w);
        r.x *= r.x; r.y *= r.y; r.z *= r.z; r.w *= r.w;
        acc += (r.x + r.y) + (r.z + r.w);
        if (roi) *reinterpret_cast<float4*>(roi + (long long)i * d + j) = r;
        if (zs) {
          const float4 mu = *reinterpret_cast<const float4*>(stats + j);
          const float4 sd = *reinterpret_cast<const float4*>(stats + d + j);
          *reinterpret_cast<float4*>(zs + (long long)i * d + j) =
              make_float4((r.x - mu.x) / sd.x, (r.y - mu.y) / sd.y, (r.z - mu.z) / sd.z, (r.w - mu.w) / sd.w);
        }
      }
    } else {
      for (int j = lane; j < d; j += 32) {
        float r = xr[j] - hr[j];
        r *= r;
        acc += r;
        if (roi) roi[(long long)i * d + j] = r;
        if (zs) zs[(long long)i * d + j] = (r - stats[j]) / stats[d + j];
      }
    }
    acc = warp_sum(acc);
    if (subj && lane == 0) subj[i] = acc / d;
  }
}

void launch_deviation(const SegTable& t, int n_seg, int max_rows, cudaStream_t st) {
  if (n_seg == 0 || max_rows == 0) return;
  int bx = (max_rows + 7) / 8;
  const int cap = max(1, (sm_count() * 8) / n_seg);   // ~8 CTAs per SM over all segments
  if (bx > cap) bx = cap;
  dim3 grid(bx, n_seg);
  deviation_kernel<<<grid, 256, 0, st>>>(t);
}

// ---- ROC-AUC by exact pair counting ------------------------------------------------------------
// CTA = (group of 8 adjacent columns, segment); warp w owns column 8 g + w.  The score table is row-major, so the 8
// columns of a row are 32 contiguous bytes = one DRAM sector: tiles of rows are read with every byte of every sector
// used (a CTA per column with stride n_cols used 4 of 32) and transposed into shared memory.  Per column the
// negatives of a tile are sorted (warp-level bitonic network in shared memory, +inf sentinels for the non-negatives)
// and every positive of every tile is binary-searched in them:  U2 += 2 #(neg < s) + #(neg == s).  Integer
// arithmetic -> bit-exact against oracle/deviation.py:auc_pairs (== sklearn roc_curve + auc).
constexpr int kAucCols = 8;       // columns per CTA = warps per CTA
constexpr int kAucTile = 1024;    // rows per tile (2 x 32 KB of shared memory)

// rows [r0, r0 + m) x columns [c0, c0 + 8) of a row-major table -> dst[col][row]; `pad` fills columns / rows beyond
// the table.  A warp reads 4 rows x 8 columns per instruction (4 full sectors).
__device__ __forceinline__ void auc_load_tile(const float* __restrict__ sc, int n_cols, int c0, int r0, int m, int p2,
                                              const uint8_t* __restrict__ lab, bool negatives_only, float* dst, int stride) {
  const int cj = threadIdx.x & 7;
  const bool col_ok = c0 + cj < n_cols;
  for (int i = threadIdx.x >> 3; i < p2; i += 32) {
    float v = __int_as_float(0x7f800000);                       // +inf sorts behind every real negative
    if (i < m && col_ok && (!negatives_only || lab[r0 + i] == 0)) v = sc[(long long)(r0 + i) * n_cols + c0 + cj];
    dst[cj * stride + i] = v;
  }
}

__device__ __forceinline__ void warp_bitonic_sort(float* v, int n_pow2, int lane) {
  for (int k = 2; k <= n_pow2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = lane; i < n_pow2; i += 32) {
        const int p = i ^ j;
        if (p > i) {
          const float a = v[i], b = v[p];
          const bool up = (i & k) == 0;
          if ((a > b) == up) { v[i] = b; v[p] = a; }
        }
      }
      __syncwarp();
    }
  }
}

// stride = rows per column array in shared memory: the launch sizes it to the largest segment (<= kAucTile), so that the
// usual 200-row test sets need 13 KB per CTA instead of 64 KB and five times as many CTAs are resident.
__global__ void __launch_bounds__(256) auc_kernel(AucTable t, int stride) {
  extern __shared__ float auc_smem[];
  float* negs = auc_smem;                                  // [8][stride] negatives of the current tile (sorted / compacted)
  float* vals = auc_smem + kAucCols * stride;               // [8][stride] raw scores of the current positive tile
  __shared__ int s_cnt[8];
  __shared__ unsigned short s_idx[kAucTile];                // single-tile path: negatives first, positives from the back
  const int s = blockIdx.y;
  const int n_cols = t.n_cols[s];
  const int c0 = blockIdx.x * kAucCols;
  if (c0 >= n_cols) return;
  const float* sc = t.scores[s];
  const uint8_t* lab = t.labels[s];
  const int n = t.n_rows[s];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const bool active = c0 + w < n_cols;                        // warp-uniform; barriers below stay CTA-uniform
  unsigned long long u2 = 0;
  int n_neg_total = 0, n_pos = 0;
  if (n <= kAucTile) {
    // ---- the usual case (a fold's test set): one tile holds every row.  No sort: the labels are shared by the 8
    // columns, so warp 0 compacts the row indices once (negatives from the front, positives from the back), every warp
    // gathers its column's negatives into a dense array and counts, for each of its positives, the negatives below /
    // equal by a broadcast scan (n_pos * n_neg / 32 steps per lane; cheaper than sorting for n <= 1024).
    auc_load_tile(sc, n_cols, c0, 0, n, n, lab, false, vals, stride);
    if (w == 0) {
      int nn = 0, np = 0;
      for (int i0 = 0; i0 < n; i0 += 32) {
        const int i = i0 + lane;
        const bool is_neg = i < n && lab[i] == 0, is_pos = i < n && lab[i] != 0;
        const unsigned bn = __ballot_sync(0xffffffffu, is_neg), bp = __ballot_sync(0xffffffffu, is_pos);
        const unsigned below = (1u << lane) - 1u;
        if (is_neg) s_idx[nn + __popc(bn & below)] = (unsigned short)i;
        if (is_pos) s_idx[n - 1 - (np + __popc(bp & below))] = (unsigned short)i;
        nn += __popc(bn); np += __popc(bp);
      }
      if (lane == 0) { s_cnt[0] = nn; s_cnt[1] = np; }
    }
    __syncthreads();
    const int n_neg = s_cnt[0];
    n_pos = s_cnt[1];
    n_neg_total = n_neg;
    if (!active) return;
    const float* pv = vals + w * stride;
    float* mine = negs + w * stride;
    // dense negatives, padded to a multiple of 4 with NaN (NaN < v and NaN == v are both false)
    const int n_neg4 = (n_neg + 3) & ~3;
    for (int q = lane; q < n_neg4; q += 32) mine[q] = q < n_neg ? pv[s_idx[q]] : __int_as_float(0x7fc00000);
    __syncwarp();
    for (int p = lane; p < n_pos; p += 32) {
      const float v = pv[s_idx[n - 1 - p]];
      // counts <= 1024 are exact in fp32: (u < v) as 1.0f / 0.0f is one FSET, accumulating it one FADD -- half the
      // instructions of the integer form (compare + predicated add compiled to ~10 instructions per negative); the four
      // negatives of a step come from one 16-byte broadcast read
      float less = 0.f, eq = 0.f;
      for (int q = 0; q < n_neg4; q += 4) {
        const float4 u = *reinterpret_cast<const float4*>(mine + q);
        less += (u.x < v ? 1.f : 0.f) + (u.y < v ? 1.f : 0.f);
        eq += (u.x == v ? 1.f : 0.f) + (u.y == v ? 1.f : 0.f);
        less += (u.z < v ? 1.f : 0.f) + (u.w < v ? 1.f : 0.f);
        eq += (u.z == v ? 1.f : 0.f) + (u.w == v ? 1.f : 0.f);
      }
      u2 += 2ull * (unsigned long long)less + (unsigned long long)eq;
    }
  } else {
    // ---- large tables: tiles of kAucTile rows; the negatives of a tile are sorted and the positives of every tile are
    // binary-searched in them
    for (int r0 = 0; r0 < n; r0 += kAucTile) {
      const int m = min(kAucTile, n - r0);
      int p2 = 32;
      while (p2 < m) p2 <<= 1;
      int cnt = 0;
      for (int i = threadIdx.x; i < m; i += 256) cnt += lab[r0 + i] == 0;
      for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
      __syncthreads();                                         // previous tile fully consumed (negs, s_cnt)
      if (lane == 0) s_cnt[w] = cnt;
      auc_load_tile(sc, n_cols, c0, r0, m, p2, lab, true, negs, stride);
      __syncthreads();
      int n_neg = 0;
      for (int k = 0; k < 8; ++k) n_neg += s_cnt[k];
      n_neg_total += n_neg;
      float* mine = negs + w * stride;
      if (active && n_neg > 0) warp_bitonic_sort(mine, p2, lane);
      if (n_neg == 0) continue;                                 // CTA-uniform
      for (int q0 = 0; q0 < n; q0 += kAucTile) {                // positives of every tile against this tile's negatives
        const int mq = min(kAucTile, n - q0);
        __syncthreads();
        auc_load_tile(sc, n_cols, c0, q0, mq, mq, lab, false, vals, stride);
        __syncthreads();
        if (!active) continue;
        const float* pv = vals + w * stride;
        for (int i = lane; i < mq; i += 32) {
          if (lab[q0 + i] == 0) continue;
          const float v = pv[i];
          int lo = 0, hi = n_neg;
          while (lo < hi) { const int mid = (lo + hi) >> 1; if (mine[mid] < v) lo = mid + 1; else hi = mid; }
          const int less = lo;
          hi = n_neg;
          while (lo < hi) { const int mid = (lo + hi) >> 1; if (mine[mid] <= v) lo = mid + 1; else hi = mid; }
          u2 += 2ull * less + (unsigned long long)(lo - less);
        }
      }
    }
    if (!active) return;
    for (int i = lane; i < n; i += 32) n_pos += lab[i] != 0;
    for (int o = 16; o > 0; o >>= 1) n_pos += __shfl_xor_sync(0xffffffffu, n_pos, o);
  }
  for (int o = 16; o > 0; o >>= 1) u2 += __shfl_xor_sync(0xffffffffu, u2, o);
  if (lane == 0) {
    const int col = c0 + w;
    if (t.out_u2 && t.out_u2[s]) t.out_u2[s][col] = u2;
    t.out_auc[s][col] = (n_pos > 0 && n_neg_total > 0)
        ? (double)u2 / (2.0 * (double)n_pos * (double)n_neg_total) : nan("");
  }
}

void launch_auc(const AucTable& t, int n_seg, int max_cols, int max_rows, cudaStream_t st) {
  if (n_seg == 0 || max_cols == 0) return;
  int stride = kAucTile;                          // multi-tile path: full tiles (power-of-two sort width <= kAucTile)
  if (max_rows <= kAucTile) stride = max_rows < 32 ? 32 : ((max_rows + 31) & ~31);
  const int smem = 2 * kAucCols * stride * (int)sizeof(float);
  cudaFuncSetAttribute(auc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * kAucCols * kAucTile * (int)sizeof(float));
  dim3 grid((max_cols + kAucCols - 1) / kAucCols, n_seg);
  auc_kernel<<<grid, 256, smem, st>>>(t, stride);
}

// ---- latent-space normative deviation (utils_vae.py:155-161) --------------------------------------------------
// grid = (row slices, n_seg).  Every CTA first reduces the reference rows of its segment to per-dimension mean and
// population variance (fp64 accumulation, fixed combine order; the [n_train][Z] table is tiny and L2-resident, so the
// row slices of one segment recompute it instead of synchronising), then streams its slice of the scored rows.
constexpr int kLatMaxZ = 128;
__global__ void __launch_bounds__(256) latent_deviation_kernel(LatentTable t) {
  __shared__ double s_sum[256], s_sq[256];
  __shared__ float s_mean[kLatMaxZ], s_var[kLatMaxZ];
  const int s = blockIdx.y;
  const int Z = t.latent[s], nt = t.n_train[s], n = t.n_rows[s];
  const float* mt = t.mu_train[s];
  // thread = (row group rg, dimension k): 256 / Zp row groups, Zp = Z rounded up to a power of two <= 128
  int zp = 1;
  while (zp < Z) zp <<= 1;
  const int k = threadIdx.x % zp, rg = threadIdx.x / zp, n_rg = 256 / zp;
  double sum = 0.0, sq = 0.0;
  if (k < Z)
    for (int i = rg; i < nt; i += n_rg) { const double v = (double)mt[(long long)i * Z + k]; sum += v; sq += v * v; }
  s_sum[threadIdx.x] = sum; s_sq[threadIdx.x] = sq;
  __syncthreads();
  if (rg == 0 && k < Z) {
    double a = 0.0, b = 0.0;
    for (int g = 0; g < n_rg; ++g) { a += s_sum[g * zp + k]; b += s_sq[g * zp + k]; }
    const double mean = nt > 0 ? a / nt : 0.0;
    double var = nt > 0 ? b / nt - mean * mean : 0.0;
    if (var < 0.0) var = 0.0;
    s_mean[k] = (float)mean; s_var[k] = (float)var;
  }
  __syncthreads();
  const float* mu = t.mu[s];
  const float* lv = t.logvar[s];
  float* oz = t.out_z ? t.out_z[s] : nullptr;
  float* od = t.out_dev ? t.out_dev[s] : nullptr;
  // one row per group of zp threads (zp <= 32: sub-warp shuffle reduction; larger: shared memory).  The loop is
  // uniform over the CTA (row index guarded inside), so the barriers of the wide case are never divergent.
  __shared__ float s_row[8];
  for (int base = blockIdx.x * n_rg; base < n; base += gridDim.x * n_rg) {
    const int i = base + rg;
    float a = 0.f;
    if (i < n && k < Z) {
      const float zz = (mu[(long long)i * Z + k] - s_mean[k]) / sqrtf(s_var[k] + expf(lv[(long long)i * Z + k]));
      if (oz) oz[(long long)i * Z + k] = zz;
      a = fabsf(zz);
    }
    if (zp <= 32) {
      for (int o = zp >> 1; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
      if (od && i < n && k == 0) od[i] = a / Z;
    } else {
      a = warp_sum(a);
      if ((threadIdx.x & 31) == 0) s_row[threadIdx.x >> 5] = a;
      __syncthreads();
      if (od && i < n && k == 0) {
        float tot = 0.f;
        for (int w = 0; w < zp / 32; ++w) tot += s_row[rg * (zp / 32) + w];
        od[i] = tot / Z;
      }
      __syncthreads();
    }
  }
}

void launch_latent_deviation(const LatentTable& t, int n_seg, int max_rows, cudaStream_t st) {
  if (n_seg == 0) return;
  int bx = (max_rows + 255) / 256;
  const int cap = max(1, (sm_count() * 4) / n_seg);
  if (bx > cap) bx = cap;
  if (bx < 1) bx = 1;
  dim3 grid(bx, n_seg);
  latent_deviation_kernel<<<grid, 256, 0, st>>>(t);
}

// ---- misc -----------------------------------------------------------------------------------
// Fixed-size float64 record per segment {subject AUC | per-ROI mean | std | AUC (d_max each) | per-subject deviation
// (n_test_max, NaN-padded)}: the payload of the multi-GPU all-gather, assembled in one pass from the
// scorer's flat result buffers.  CTA = one segment.
__global__ void __launch_bounds__(256) member_records_kernel(const float* __restrict__ stats, const long long* __restrict__ o_stats,
                                                             const double* __restrict__ auc_roi, const long long* __restrict__ o_auc,
                                                             const double* __restrict__ auc_subj, const float* __restrict__ subj,
                                                             const long long* __restrict__ o_subj, const int* __restrict__ seg_d,
                                                             const int* __restrict__ n_test, int d_max, int n_test_max,
                                                             double* __restrict__ out) {
  const int s = blockIdx.x, d = seg_d[s], n = n_test[s];
  const long long w = 1 + 3LL * d_max + n_test_max;
  double* row = out + (long long)s * w;
  const float* st = stats + o_stats[s];
  const double* au = auc_roi + o_auc[s];
  const float* sj = subj + o_subj[s];
  if (threadIdx.x == 0) row[0] = auc_subj[s];
  for (int j = threadIdx.x; j < d_max; j += blockDim.x) {
    row[1 + j] = j < d ? (double)st[j] : 0.0;
    row[1 + d_max + j] = j < d ? (double)st[d + j] : 0.0;
    row[1 + 2 * d_max + j] = j < d ? au[j] : 0.0;
  }
  const double nan = __longlong_as_double(0x7ff8000000000000LL);
  for (int j = threadIdx.x; j < n_test_max; j += blockDim.x) row[1 + 3 * d_max + j] = j < n ? (double)sj[j] : nan;
}

void launch_member_records(const float* stats, const long long* o_stats, const double* auc_roi, const long long* o_auc,
                           const double* auc_subj, const float* subj, const long long* o_subj, const int* seg_d, const int* n_test,
                           int n_seg, int d_max, int n_test_max, double* out, cudaStream_t st) {
  if (n_seg > 0) member_records_kernel<<<n_seg, 256, 0, st>>>(stats, o_stats, auc_roi, o_auc, auc_subj, subj, o_subj, seg_d, n_test,
                                                               d_max, n_test_max, out);
}

__global__ void mean_rows_kernel(PtrTable16 src, int k, long long n, float* out) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    float a = 0.f;
    for (int j = 0; j < k; ++j) a += src.p[j][i];
    out[i] = a / k;
  }
}

void launch_mean_rows(const PtrTable16& src, int k, long long n, float* out, cudaStream_t st) {
  if (n == 0) return;
  const int blocks = (int)min((n + 255) / 256, (long long)sm_count() * 8);
  mean_rows_kernel<<<blocks, 256, 0, st>>>(src, k, n, out);
}

__global__ void adam_kernel(float* p, const float* g, float* m, float* v, long long n, float step_size,
                            float bc2_sqrt, float b1, float b2, float eps) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float grad = g[i];
    const float m1 = b1 * m[i] + (1.f - b1) * grad;
    const float v1 = b2 * v[i] + (1.f - b2) * grad * grad;
    m[i] = m1; v[i] = v1;
    p[i] -= step_size * (m1 / (sqrtf(v1) / bc2_sqrt + eps));
  }
}

void launch_adam(float* p, const float* g, float* m, float* v, long long n, float step_size, float bc2_sqrt,
                 float b1, float b2, float eps, cudaStream_t st) {
  if (n == 0) return;
  const int blocks = (int)min((n + 255) / 256, (long long)sm_count() * 8);
  adam_kernel<<<blocks, 256, 0, st>>>(p, g, m, v, n, step_size, bc2_sqrt, b1, b2, eps);
}

__global__ void philox_kernel(unsigned long long seed, unsigned long long step, uint32_t stream_id, long long n,
                              float* out) {
  for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g * 4 < n;
       g += (long long)gridDim.x * blockDim.x) {
    float v[4];
    philox_normal4(seed, step, stream_id, (uint32_t)g, v);
    for (int j = 0; j < 4; ++j)
      if (g * 4 + j < n) out[g * 4 + j] = v[j];
  }
}

void launch_philox(unsigned long long seed, unsigned long long step, uint32_t stream_id, long long n, float* out,
                   cudaStream_t st) {
  if (n == 0) return;
  const long long groups = (n + 3) / 4;
  const int blocks = (int)min((groups + 255) / 256, (long long)sm_count() * 8);
  philox_kernel<<<blocks, 256, 0, st>>>(seed, step, stream_id, n, out);
}

}  // namespace nmb
