// shq_scan.cuh -- device-wide exclusive scans (reduce / scan-of-sums / apply), generic over an
// associative operator.  Used for bin offsets, compaction offsets and the max-plus scan that turns
// (quotient, encoded length) pairs into canonical CQF slot positions.
#pragma once
#include "shq_common.cuh"

namespace shq {

constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;

struct SumU32 { typedef uint32_t T; __device__ __host__ static T id() { return 0; } __device__ __host__ static T op(T a, T b) { return a + b; } };
struct SumU64 { typedef uint64_t T; __device__ __host__ static T id() { return 0; } __device__ __host__ static T op(T a, T b) { return a + b; } };

// f(e) = max(e + a, b): composition is associative; scanning these maps over the sorted entries
// yields the first free slot before each entry (see layout kernels).
struct MaxPlus { long long a, b; };
constexpr long long kNegInf = -(1LL << 60);
struct MaxPlusOp {
  typedef MaxPlus T;
  __device__ __host__ static T id() { return MaxPlus{0, kNegInf}; }
  __device__ __host__ static T op(T x, T y) {  // apply x first, then y
    long long b1 = x.b + y.a;
    return MaxPlus{x.a + y.a, b1 > y.b ? b1 : y.b};
  }
};

template <class T>
__device__ __forceinline__ T shfl_up_any(T v, int delta) {
  static_assert(sizeof(T) % 4 == 0, "shfl size");
  T r;
  const uint32_t* s = reinterpret_cast<const uint32_t*>(&v);
  uint32_t* d = reinterpret_cast<uint32_t*>(&r);
#pragma unroll
  for (int i = 0; i < (int)(sizeof(T) / 4); i++) d[i] = __shfl_up_sync(0xffffffffu, s[i], delta);
  return r;
}

// Exclusive scan of one value per thread across a block of kScanThreads threads.
// Returns the exclusive prefix; *total receives the block aggregate (valid in all threads).
template <class Op>
__device__ typename Op::T block_exclusive_scan(typename Op::T v, typename Op::T* total) {
  typedef typename Op::T T;
  __shared__ T warp_tot[32];
  __shared__ T block_tot;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  T inc = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    T o = shfl_up_any(inc, d);
    if (lane >= d) inc = Op::op(o, inc);
  }
  if (lane == 31) warp_tot[wid] = inc;
  __syncthreads();
  if (wid == 0) {
    T w = lane < nw ? warp_tot[lane] : Op::id();
    T winc = w;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      T o = shfl_up_any(winc, d);
      if (lane >= d) winc = Op::op(o, winc);
    }
    T wexc = shfl_up_any(winc, 1);
    if (lane == 0) wexc = Op::id();
    if (lane < nw) warp_tot[lane] = wexc;
    if (lane == 31) block_tot = winc;
  }
  __syncthreads();
  T exc = shfl_up_any(inc, 1);
  if (lane == 0) exc = Op::id();
  T res = Op::op(warp_tot[wid], exc);
  *total = block_tot;
  __syncthreads();
  return res;
}

template <class Op>
__global__ void __launch_bounds__(kScanThreads) scan_reduce_kernel(const typename Op::T* in, typename Op::T* sums, size_t n) {
  typedef typename Op::T T;
  size_t base = (size_t)blockIdx.x * kScanTile + (size_t)threadIdx.x * kScanItems;
  T acc = Op::id();
#pragma unroll
  for (int i = 0; i < kScanItems; i++)
    if (base + i < n) acc = Op::op(acc, in[base + i]);
  T tot;
  block_exclusive_scan<Op>(acc, &tot);
  if (threadIdx.x == 0) sums[blockIdx.x] = tot;
}

// in-place exclusive scan inside each tile, seeded with offs[blockIdx.x] (or identity if offs==null)
template <class Op>
__global__ void __launch_bounds__(kScanThreads) scan_apply_kernel(typename Op::T* data, const typename Op::T* offs, size_t n,
                                                                 typename Op::T* grand_total) {
  typedef typename Op::T T;
  size_t base = (size_t)blockIdx.x * kScanTile + (size_t)threadIdx.x * kScanItems;
  T loc[kScanItems];
  T acc = Op::id();
#pragma unroll
  for (int i = 0; i < kScanItems; i++) {
    loc[i] = base + i < n ? data[base + i] : Op::id();
    acc = Op::op(acc, loc[i]);
  }
  T tot;
  T exc = block_exclusive_scan<Op>(acc, &tot);
  T run = offs ? Op::op(offs[blockIdx.x], exc) : exc;
#pragma unroll
  for (int i = 0; i < kScanItems; i++) {
    if (base + i < n) data[base + i] = run;
    run = Op::op(run, loc[i]);
  }
  if (grand_total && blockIdx.x == gridDim.x - 1 && threadIdx.x == kScanThreads - 1) *grand_total = run;
}

// Host driver.  scratch must hold ceil(n/tile) + ceil(n/tile^2) + ... elements (+ a few).
// If total != nullptr, *total (device pointer) receives the reduction of all n elements.
template <class Op>
void exclusive_scan_inplace(typename Op::T* d, size_t n, DevBuf& scratch, cudaStream_t st, typename Op::T* d_total,
                            uint64_t* launches) {
  typedef typename Op::T T;
  if (n == 0) {
    if (d_total) { T z = Op::id(); SHQ_CUDA(cudaMemcpyAsync(d_total, &z, sizeof(T), cudaMemcpyHostToDevice, st)); }
    return;
  }
  // level sizes
  std::vector<size_t> sizes;
  size_t m = n;
  while (m > (size_t)kScanTile) { m = (m + kScanTile - 1) / kScanTile; sizes.push_back(m); }
  size_t tot_elems = 0;
  for (size_t s : sizes) tot_elems += s;
  scratch.reserve((tot_elems + 16) * sizeof(T));
  std::vector<T*> lvl;
  T* p = scratch.as<T>();
  for (size_t s : sizes) { lvl.push_back(p); p += s; }
  // reduce up
  const T* src = d; size_t cur = n;
  for (size_t i = 0; i < sizes.size(); i++) {
    scan_reduce_kernel<Op><<<(unsigned)sizes[i], kScanThreads, 0, st>>>(src, lvl[i], cur);
    if (launches) (*launches)++;
    src = lvl[i]; cur = sizes[i];
  }
  // scan down
  for (int i = (int)sizes.size() - 1; i >= 0; i--) {
    size_t len = sizes[i];
    const T* offs = (i + 1 < (int)sizes.size()) ? lvl[i + 1] : nullptr;
    scan_apply_kernel<Op><<<(unsigned)((len + kScanTile - 1) / kScanTile), kScanThreads, 0, st>>>(lvl[i], offs, len, nullptr);
    if (launches) (*launches)++;
  }
  scan_apply_kernel<Op><<<(unsigned)((n + kScanTile - 1) / kScanTile), kScanThreads, 0, st>>>(d, sizes.empty() ? nullptr : lvl[0], n, d_total);
  if (launches) (*launches)++;
  SHQ_CUDA(cudaGetLastError());
}

}  // namespace shq
