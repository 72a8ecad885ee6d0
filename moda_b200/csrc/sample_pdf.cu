// Importance sampling of ray depths (nnutils/rendering.py:582-623) and the sorted union with the coarse
// depths (rendering.py:103-110).  One CTA per ray: the CDF is built in shared memory, every thread
// inverts it for one uniform sample (binary search = torch.searchsorted(right=True)), and the merged depth
// list is ordered with a bitonic sort in shared memory.
#include "common.cuh"

namespace moda {

constexpr int PDF_THREADS = 128;
constexpr int PDF_MAX = 1024;  // max bins and max merged samples per ray

__device__ __forceinline__ float linspace01_pdf(int i, int n) {
  if (n == 1) return 0.f;
  const float step = 1.0f / (float)(n - 1);
  return (i < n / 2) ? step * (float)i : 1.0f - step * (float)(n - 1 - i);
}

// merge: bins are mid-points of z (S-1 of them), weights = w[1:S-1]; out = sort(z ++ samples) (R, S+NI)
// plain: bins (R, n+1), weights (R, n); out = samples (R, NI)
__global__ void __launch_bounds__(PDF_THREADS) sample_pdf_kernel(const float* z, const float* bins,
                                                                 const float* weights, const float* u, float* out,
                                                                 int S, int n, int NI, int det, float eps,
                                                                 int merge) {
  __shared__ float s_bins[PDF_MAX + 1];
  __shared__ float s_cdf[PDF_MAX + 1];
  __shared__ float s_sort[PDF_MAX];
  __shared__ float s_tot;
  const int r = blockIdx.x;
  const int tid = threadIdx.x;
  const float* w = merge ? weights + (size_t)r * S + 1 : weights + (size_t)r * n;
  for (int i = tid; i < n + 1; i += PDF_THREADS)
    s_bins[i] = merge ? 0.5f * (z[(size_t)r * S + i] + z[(size_t)r * S + i + 1]) : bins[(size_t)r * (n + 1) + i];
  if (tid == 0) {
    float tot = 0.f;
    for (int i = 0; i < n; ++i) tot += w[i] + eps;
    s_tot = tot;
    float c = 0.f;
    s_cdf[0] = 0.f;
    for (int i = 0; i < n; ++i) {
      c += (w[i] + eps) / tot;
      s_cdf[i + 1] = c;
    }
  }
  __syncthreads();
  for (int j = tid; j < NI; j += PDF_THREADS) {
    const float uj = det ? linspace01_pdf(j, NI) : u[(size_t)r * NI + j];
    // number of cdf entries <= uj  (searchsorted, right=True) over n+1 entries
    int lo = 0, hi = n + 1;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (s_cdf[mid] <= uj) lo = mid + 1; else hi = mid;
    }
    const int below = max(lo - 1, 0), above = min(lo, n);
    float denom = s_cdf[above] - s_cdf[below];
    if (denom < eps) denom = 1.f;
    const float smp = s_bins[below] + (uj - s_cdf[below]) / denom * (s_bins[above] - s_bins[below]);
    if (merge) s_sort[S + j] = smp; else out[(size_t)r * NI + j] = smp;
  }
  if (!merge) return;
  const int total = S + NI;
  int pow2 = 1;
  while (pow2 < total) pow2 <<= 1;
  for (int i = tid; i < S; i += PDF_THREADS) s_sort[i] = z[(size_t)r * S + i];
  for (int i = total + tid; i < pow2; i += PDF_THREADS) s_sort[i] = INFINITY;
  __syncthreads();
  for (int k = 2; k <= pow2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < pow2; i += PDF_THREADS) {
        const int p = i ^ j;
        if (p > i) {
          const float a = s_sort[i], b = s_sort[p];
          const bool up = (i & k) == 0;
          if ((a > b) == up) { s_sort[i] = b; s_sort[p] = a; }
        }
      }
      __syncthreads();
    }
  }
  for (int i = tid; i < total; i += PDF_THREADS) out[(size_t)r * total + i] = s_sort[i];
}

}  // namespace moda

using namespace moda;

extern "C" int moda_sample_pdf(const float* z, const float* bins, const float* weights, const float* u,
                               float* out, int R, int S, int n, int NI, int det, float eps, int merge,
                               cudaStream_t stream) {
  MODA_REQUIRE(weights && out && (det || u), "sample_pdf: null pointer");
  MODA_REQUIRE(merge ? (z != nullptr && n == S - 2) : (bins != nullptr), "sample_pdf: bad mode arguments");
  MODA_REQUIRE(n >= 1 && n <= PDF_MAX && NI >= 1 && (!merge || S + NI <= PDF_MAX), "sample_pdf: sizes out of range");
  if (R == 0) return 0;
  sample_pdf_kernel<<<R, PDF_THREADS, 0, stream>>>(z, bins, weights, u, out, S, n, NI, det, eps, merge);
  return check_launch("sample_pdf");
}
