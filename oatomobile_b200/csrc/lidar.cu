// LIDAR point cloud -> 2-channel 200x200 bird's-eye-view histogram (sm_100a).
//
// Replaces `carla_lidar_measurement_to_ndarray` (oatomobile/utils/carla.py:165-233):
// points [N,3] float32 are split at z = -2.5 (a point exactly at -2.5 lands in both
// halves, as written), each half is binned with np.histogramdd over
// np.linspace(-50, 51, 201) edges in x and y, clipped at 5 hits and divided by 5.
// Bit-exact with NumPy: edges are rebuilt as `i * step + start` in float64 with separate
// rounding of the product and the sum (no FMA), the last edge is exactly `stop`,
// bin = (#edges <= x) - 1 with the right-most edge closed.  Integer atomics in
// global memory (L2), one thread per point; 80 000 counters.
#include "common.cuh"

namespace oat {
namespace {

constexpr int kBins = 200;

__device__ __forceinline__ double edge_at(int i, double start, double stop, double step) {
  return i == kBins ? stop : __dadd_rn(__dmul_rn((double)i, step), start);
}

__device__ __forceinline__ int bin_of(float v, double start, double stop, double step) {
  const double x = (double)v;
  if (!(x >= start) || x > stop) return -1;  // also rejects NaN
  int i = (int)floor((x - start) / step);
  i = max(0, min(i, kBins));
  while (i < kBins && edge_at(i + 1, start, stop, step) <= x) ++i;
  while (i > 0 && edge_at(i, start, stop, step) > x) --i;
  if (i == kBins) i = kBins - 1;             // x == right-most edge: closed last bin
  return i;
}

__global__ void __launch_bounds__(256) lidar_count_kernel(const float* __restrict__ pts, int64_t n,
                                                          double start, double stop, double step,
                                                          float z_split,
                                                          unsigned int* __restrict__ counts) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float x = __ldg(pts + 3 * i), y = __ldg(pts + 3 * i + 1), z = __ldg(pts + 3 * i + 2);
  const int bx = bin_of(x, start, stop, step), by = bin_of(y, start, stop, step);
  if (bx < 0 || by < 0) return;
  if (z <= z_split) atomicAdd(counts + (bx * kBins + by) * 2 + 0, 1u);  // "below"
  if (z >= z_split) atomicAdd(counts + (bx * kBins + by) * 2 + 1, 1u);  // "above"
}

__global__ void __launch_bounds__(256) lidar_finalize_kernel(const unsigned int* __restrict__ counts,
                                                             float* __restrict__ out,
                                                             unsigned int hist_max) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= kBins * kBins * 2) return;
  const unsigned int c = min(counts[i], hist_max);
  out[i] = (float)((double)c / (double)hist_max);
}

}  // namespace

int launch_lidar_bev(const float* points, int64_t n, int pixels_per_meter, int hist_max,
                     int meters_max, unsigned int* counts, float* out, cudaStream_t stream) {
  const int bins = meters_max * 2 * pixels_per_meter;
  if (bins != kBins) return fail("oat_lidar_bev: only the 200x200 grid (50 m, 2 px/m) is built");
  if (hist_max < 1) return fail("oat_lidar_bev: hist_max_per_pixel must be >= 1");
  const double start = -(double)meters_max, stop = (double)meters_max + 1.0;
  const double step = (stop - start) / (double)bins;  // numpy.linspace: delta / div
  OAT_CUDA(cudaMemsetAsync(counts, 0, sizeof(unsigned int) * kBins * kBins * 2, stream));
  if (n > 0) {
    lidar_count_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(points, n, start, stop, step,
                                                                       -2.5f, counts);
    OAT_LAUNCH_CHECK();
  }
  lidar_finalize_kernel<<<(kBins * kBins * 2 + 255) / 256, 256, 0, stream>>>(counts, out,
                                                                            (unsigned int)hist_max);
  OAT_LAUNCH_CHECK();
  return 0;
}

}  // namespace oat
