// clp_sparse.cuh -- compact-row copy of the affinity matrix and its sweep.
//
// The consistency graph is sparse (14.9 % at BASELINE.json's config 2), so after the dense build the
// non-neutral entries (everything except the -0.0 "inconsistent" code) of every local row are compacted,
// in column order, into
//      val[]  : the stored element itself (fp32 / fp64; sign bit = constraint bit, as in the dense store)
//      off16[]: 8 * (column - first column of its segment) -- the byte offset of v[column] inside the
//               shared-memory copy of the segment (segments are <= 4096 columns wide)
// = 6 bytes per kept entry (fp32 storage) instead of 4 bytes per matrix element.
// Layout is SEGMENT-major: all row slices of column segment 0, then segment 1, ...; inside a segment the
// slices of consecutive rows are adjacent, so the 32 rows x 1 segment a CTA works on are one contiguous
// range of HBM.  Every slice is padded to a multiple of 4 entries (16-byte value loads); padding entries
// point at a zero slot behind the staged segment.  slice(row, seg) in units of 4 entries:
//      [ ptr4[seg * (rows_pad + 1) + row] , ptr4[seg * (rows_pad + 1) + row + 1] )
// The sweep keeps the first-generation decomposition (segments of v staged in shared memory, 32-row tiles,
// per-segment partial products, fixed-order combine); only the inner row sweep changes.
// "plain" matrices (every kept entry has M > 0 and C = 1 -- always true after scorePairwiseConsistency) take
// a shorter path: Chat v is then just the sum of the gathered v.
// Algorithmic bytes per objective evaluation: 6 * stored entries + 4 * NSEG * (rows + 1).
#pragma once

namespace clp {

constexpr unsigned int kZeroSlot = kSegMax * 8;  // byte offset of the zero element behind the staged segment

struct SparseView {
  const void* val;             // T [4 * n4]
  const unsigned short* off16; // [4 * n4]
  const unsigned int* ptr4;    // [NSEG][rows_pad + 1]
  int rows_pad;
  int plain;                   // every kept entry has M > 0 and C = 1
};

template <typename T> __device__ __forceinline__ bool is_neutral(T s);
template <> __device__ __forceinline__ bool is_neutral<float>(float s) { return __float_as_uint(s) == 0x80000000u; }
template <> __device__ __forceinline__ bool is_neutral<double>(double s) {
  return (unsigned long long)__double_as_longlong(s) == 0x8000000000000000ULL;
}
template <typename T> __device__ __forceinline__ bool is_plain(T s) { return s > T(0); }

// pass 1: one warp per (padded) local row counts the kept entries per column segment:
//   cnt4[seg * (rows_pad + 1) + row] = number of 4-entry units of slice (row, seg)
//   totals[0] += kept entries, totals[1] += kept entries that are not "plain"
template <typename T>
__global__ void sparse_count_kernel(const T* M, long long ld, int m, int rows, int rows_pad, int W, int nseg,
                                    unsigned int* cnt4, unsigned long long* totals) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= rows_pad) return;
  unsigned int real = 0, odd = 0;
  for (int s = 0; s < nseg; ++s) {
    unsigned int c = 0, o = 0;
    if (warp < rows) {
      const int c0 = s * W, c1 = min(m, c0 + W);
      for (int j = c0 + lane * 4; j < c1; j += 128) {
        const T* p = M + (size_t)warp * ld + j;
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (j + e < c1 && !is_neutral<T>(p[e])) { ++c; o += is_plain<T>(p[e]) ? 0u : 1u; }
      }
    }
#pragma unroll
    for (int k = 16; k > 0; k >>= 1) { c += __shfl_xor_sync(0xffffffffu, c, k); o += __shfl_xor_sync(0xffffffffu, o, k); }
    if (lane == 0) cnt4[(size_t)s * (rows_pad + 1) + warp] = (c + 3u) >> 2;
    real += c; odd += o;
  }
  if (lane == 0) {
    if (real) atomicAdd(&totals[0], (unsigned long long)real);
    if (odd) atomicAdd(&totals[1], (unsigned long long)odd);
  }
}

// pass 2a: one block per column segment: exclusive scan (in place) of that segment's row counts, walking the
// array in coalesced tiles of 1024; the segment total goes to segtot[seg]
__global__ void sparse_scan_seg_kernel(unsigned int* cnt4, int n, unsigned long long* segtot) {
  __shared__ unsigned int wsum[32];
  __shared__ unsigned int carry_s;
  unsigned int* a = cnt4 + (size_t)blockIdx.x * n;
  const int t = threadIdx.x, lane = t & 31, w = t >> 5;
  unsigned long long carry = 0;
  for (int base = 0; base < n; base += 1024) {
    const int i = base + t;
    const unsigned int x = (i < n) ? a[i] : 0u;
    unsigned int inc = x;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const unsigned int y = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += y; }
    if (lane == 31) wsum[w] = inc;
    __syncthreads();
    if (w == 0) {
      unsigned int v = wsum[lane], s = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const unsigned int y = __shfl_up_sync(0xffffffffu, s, o); if (lane >= o) s += y; }
      wsum[lane] = s - v;                 // exclusive prefix of the warp sums
      if (lane == 31) carry_s = s;        // tile total
    }
    __syncthreads();
    if (i < n) a[i] = (unsigned int)(carry + wsum[w] + (inc - x));
    carry += carry_s;
    __syncthreads();
  }
  if (t == 0) segtot[blockIdx.x] = carry;
}

// pass 2b: add the start of each segment (prefix of the segment totals); block 0 also publishes the grand total
__global__ void sparse_scan_fix_kernel(unsigned int* cnt4, int n, int nseg, const unsigned long long* segtot,
                                       unsigned long long* total4) {
  unsigned long long base = 0;
  for (int s = 0; s < (int)blockIdx.x; ++s) base += segtot[s];
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    unsigned long long tot = 0;
    for (int s = 0; s < nseg; ++s) tot += segtot[s];
    *total4 = tot;
  }
  if (base == 0) return;
  unsigned int* a = cnt4 + (size_t)blockIdx.x * n;
  for (int i = threadIdx.x; i < n; i += blockDim.x) a[i] += (unsigned int)base;
}

// pass 3: one warp per row writes the kept entries of each segment in column order, then the padding.
// ptr4 is the scanned array; the slot [seg][rows_pad] of every segment holds the start of the next one.
template <typename T>
__global__ void sparse_fill_kernel(const T* M, long long ld, int m, int rows, int rows_pad, int W, int nseg,
                                   const unsigned int* ptr4, T* val, unsigned short* off16) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= rows) return;
  for (int s = 0; s < nseg; ++s) {
    const unsigned int* pp = ptr4 + (size_t)s * (rows_pad + 1) + warp;
    unsigned long long pos = 4ull * pp[0];
    const unsigned long long slice_end = 4ull * pp[1];
    const int c0 = s * W, c1 = min(m, c0 + W);
    for (int j0 = c0; j0 < c1; j0 += 128) {
      const int j = j0 + lane * 4;
      T x[4];
      unsigned int keep = 0;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        x[e] = (j + e < c1) ? M[(size_t)warp * ld + j + e] : encode<T>(0.0, false);
        keep |= (!is_neutral<T>(x[e]) ? 1u : 0u) << e;
      }
      const unsigned int cnt = __popc(keep);
      unsigned int pre = cnt;  // inclusive scan over lanes
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned int y = __shfl_up_sync(0xffffffffu, pre, o);
        if (lane >= o) pre += y;
      }
      const unsigned int total = __shfl_sync(0xffffffffu, pre, 31);
      unsigned long long w = pos + (pre - cnt);
#pragma unroll
      for (int e = 0; e < 4; ++e)
        if (keep & (1u << e)) { val[w] = x[e]; off16[w] = (unsigned short)(8 * (j + e - c0)); ++w; }
      pos += total;
    }
    for (unsigned long long w = pos + lane; w < slice_end; w += 32) { val[w] = encode<T>(0.0, false); off16[w] = (unsigned short)kZeroSlot; }
  }
}

// 4 entries of one row slice
template <typename T> struct Entry4;
template <> struct Entry4<float> {
  float4 x; uint2 k;
  __device__ __forceinline__ void load(const float* val, const unsigned short* off, unsigned long long at) {
    x = ldg_stream(reinterpret_cast<const float4*>(val + at));
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(k.x), "=r"(k.y) : "l"(off + at));
  }
  __device__ __forceinline__ float get(int e) const { return e == 0 ? x.x : e == 1 ? x.y : e == 2 ? x.z : x.w; }
  __device__ __forceinline__ void neutral() { x = make_float4(-0.f, -0.f, -0.f, -0.f); k = make_uint2(kZeroSlot | (kZeroSlot << 16), kZeroSlot | (kZeroSlot << 16)); }
};
template <> struct Entry4<double> {
  double2 a, b; uint2 k;
  __device__ __forceinline__ void load(const double* val, const unsigned short* off, unsigned long long at) {
    a = ldg_stream(reinterpret_cast<const double2*>(val + at));
    b = ldg_stream(reinterpret_cast<const double2*>(val + at) + 1);
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(k.x), "=r"(k.y) : "l"(off + at));
  }
  __device__ __forceinline__ double get(int e) const { return e == 0 ? a.x : e == 1 ? a.y : e == 2 ? b.x : b.y; }
  __device__ __forceinline__ void neutral() { a = make_double2(-0.0, -0.0); b = a; k = make_uint2(kZeroSlot | (kZeroSlot << 16), kZeroSlot | (kZeroSlot << 16)); }
};
__device__ __forceinline__ unsigned int off_of(const uint2& k, int e) {
  return e == 0 ? (k.x & 0xffffu) : e == 1 ? (k.x >> 16) : e == 2 ? (k.y & 0xffffu) : (k.y >> 16);
}
__device__ __forceinline__ double vs_at(const double* vs, unsigned int byte_off) {
  return *reinterpret_cast<const double*>(reinterpret_cast<const char*>(vs) + byte_off);
}

// rows [lr, lr+4) x the staged segment.  The four row slices are walked together; every lane takes chunks of
// 4 consecutive entries (4 x (16 + 8) bytes per lane in flight).  Few registers on purpose: the sweep is
// latency-bound, so the kernels that contain it run 3 CTAs per SM.
// vs holds the segment of v in natural order, vs[kSegMax] == 0.
template <typename T, bool PLAIN>
__device__ __forceinline__ void sparse_rows4(const SparseView& sp, const unsigned int (&a)[5], const double* vs,
                                             double (&acc)[8]) {
  const int lane = threadIdx.x & 31;
  const T* val = reinterpret_cast<const T*>(sp.val);
  unsigned int beg[4], n4[4], nmax = 0;
#pragma unroll
  for (int r = 0; r < 4; ++r) { beg[r] = a[r]; n4[r] = a[r + 1] - a[r]; nmax = max(nmax, n4[r]); }
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.0;
  for (unsigned int c = lane; c < nmax; c += 32) {
    Entry4<T> E[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      if (c < n4[r]) E[r].load(val, sp.off16, 4ull * (beg[r] + c));
      else E[r].neutral();
    }
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const double v = vs_at(vs, off_of(E[r].k, e));
        if (PLAIN) {
          acc[r] = fma((double)E[r].get(e), v, acc[r]);  // padding: -0.0 * 0.0
          acc[4 + r] += v;
        } else {
          double dummyM = 0.0, dummyC = 0.0;
          apply_elem<false>(E[r].get(e), v, 0.0, acc[r], acc[4 + r], dummyM, dummyC);
        }
      }
  }
}

// whole sparse pass of one CTA: same decomposition and partial layout as matvec_phase.
// vs must hold kSegMax + 1 doubles.
// (Tried: a plain-only instance capped at 64 registers for 4 CTAs/SM -- the sweep gained 4 %, the two
// synchronisation steps of the evaluation lost it again with 592 CTAs; not kept.)
template <typename T, bool PLAIN_ONLY>
__device__ void sparse_phase(const MatView& mv, const Plan& p, const StageArgs& st, const SparseView& sp,
                             double* partM, double* partC, double* vs, double* red_smem) {
  const int sg = blockIdx.x % p.SG, rg = blockIdx.x / p.SG;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) vs[kSegMax] = 0.0;
  for (int seg = sg; seg < p.NSEG; seg += p.SG) {
    stage_segment<double>(st, p, mv.m, seg, rg == 0, vs, red_smem);  // <double>: natural (unpermuted) order
    // slice pointers of the next item are fetched while the current one is processed
    const unsigned int* pseg = sp.ptr4 + (size_t)seg * (sp.rows_pad + 1) + warp * kRowsPerWarp;
    unsigned int nxt[5] = {0u, 0u, 0u, 0u, 0u};
    if (rg < p.NRT) {
#pragma unroll
      for (int r = 0; r < 5; ++r) nxt[r] = pseg[(size_t)rg * kRowTile + r];
    }
    for (int rt = rg; rt < p.NRT; rt += p.RG) {
      const int lr = rt * kRowTile + warp * kRowsPerWarp;
      unsigned int cur[5];
#pragma unroll
      for (int r = 0; r < 5; ++r) cur[r] = nxt[r];
      if (rt + p.RG < p.NRT) {
#pragma unroll
        for (int r = 0; r < 5; ++r) nxt[r] = pseg[(size_t)(rt + p.RG) * kRowTile + r];
      }
      double acc[8];
      if (PLAIN_ONLY || sp.plain) sparse_rows4<T, true>(sp, cur, vs, acc);
      else sparse_rows4<T, false>(sp, cur, vs, acc);
      const double tot = warp_reduce8(acc);
      if ((lane & 3) == 0) {
        const int qv = lane >> 2;  // 0..3: M of row qv, 4..7: C of row qv-4
        double* dst = (qv >> 2) ? partC : partM;
        dst[(size_t)seg * mv.rows_pad + lr + (qv & 3)] = tot;
      }
    }
    __syncthreads();  // vs is re-staged by the next segment pass
  }
}

}  // namespace clp
