// clp_sparse.cuh -- compact-row ("bit-dropped") copy of the affinity matrix and its sweep.
//
// SURVEY section 8f rank 3.  The consistency graph is sparse (14.9 % at BASELINE.json's config 2), so
// after the dense build the non-neutral entries (everything except the -0.0 "inconsistent" code) of every
// local row are compacted, in column order, into
//      val[]   : the stored element itself (fp32 / fp64, sign bit = constraint bit, as in the dense store)
//      col16[] : the column, relative to the start of its column segment (segments are <= 4096 wide)
// = 6 bytes per kept entry (fp32 storage) instead of 4 bytes per matrix element.  The sweep keeps the
// first-generation decomposition (column segments staged in shared memory, 32-row tiles, per-segment
// partial products, fixed-order combine): only the inner row sweep changes, from a dense stream to a
// gather over the row's slice of entries that fall into the staged segment.
//      slice(row, seg) = [ row_ptr[row] + seg_off[row][seg] , row_ptr[row] + seg_off[row][seg+1] )
// Algorithmic bytes per objective evaluation: 6 * nnz (+ 4 (NSEG+1) + 8 bytes of offsets per row).
#pragma once

namespace clp {

struct SparseView {
  const void* val;                  // T [nnz]
  const unsigned short* col16;      // [nnz] column - seg * W
  const unsigned long long* row_ptr;  // [rows_pad + 1]
  const unsigned int* seg_off;      // [rows_pad][NSEG + 1] offsets of the row's segment slices
  int nseg;
};

template <typename T> __device__ __forceinline__ bool is_neutral(T s);
template <> __device__ __forceinline__ bool is_neutral<float>(float s) { return __float_as_uint(s) == 0x80000000u; }
template <> __device__ __forceinline__ bool is_neutral<double>(double s) {
  return (unsigned long long)__double_as_longlong(s) == 0x8000000000000000ULL;
}

// pass 1: one warp per local row counts the kept entries per column segment.  Every (row, segment) slice is
// padded to a multiple of 4 entries so that the sweep can use 16-byte value loads; seg_off holds the padded
// exclusive prefix inside the row, row_cnt the padded row total, *real_total the number of real entries.
template <typename T>
__global__ void sparse_count_kernel(const T* M, long long ld, int m, int rows, int rows_pad, int W, int nseg,
                                    unsigned int* seg_off, unsigned long long* row_cnt, unsigned long long* real_total) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= rows_pad) return;
  unsigned int run = 0, real = 0;
  for (int s = 0; s < nseg; ++s) {
    unsigned int c = 0;
    if (warp < rows) {
      const int c0 = s * W, c1 = min(m, c0 + W);
      for (int j = c0 + lane * 4; j < c1; j += 128) {
        const T* p = M + (size_t)warp * ld + j;
#pragma unroll
        for (int e = 0; e < 4; ++e) c += (j + e < c1 && !is_neutral<T>(p[e])) ? 1u : 0u;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if (lane == 0) seg_off[(size_t)warp * (nseg + 1) + s] = run;
    real += c;
    run += (c + 3u) & ~3u;
  }
  if (lane == 0) {
    seg_off[(size_t)warp * (nseg + 1) + nseg] = run;
    row_cnt[warp] = run;
    if (real) atomicAdd(real_total, (unsigned long long)real);
  }
}

// pass 2: exclusive scan of the row counts (single block; rows_pad <= 262144)
__global__ void sparse_scan_kernel(const unsigned long long* row_cnt, int n, unsigned long long* row_ptr) {
  __shared__ unsigned long long part[1024];
  const int t = threadIdx.x, nt = blockDim.x;
  const int per = (n + nt - 1) / nt;
  const int b = t * per, e = min(n, b + per);
  unsigned long long s = 0;
  for (int i = b; i < e; ++i) s += row_cnt[i];
  part[t] = s;
  __syncthreads();
  if (t == 0) {
    unsigned long long run = 0;
    for (int i = 0; i < nt; ++i) { const unsigned long long v = part[i]; part[i] = run; run += v; }
    row_ptr[n] = run;
  }
  __syncthreads();
  unsigned long long run = part[t];
  for (int i = b; i < e; ++i) { row_ptr[i] = run; run += row_cnt[i]; }
}

// pass 3: one warp per row writes the kept entries of each segment in column order, then the padding
template <typename T>
__global__ void sparse_fill_kernel(const T* M, long long ld, int m, int rows, int W, int nseg,
                                   const unsigned long long* row_ptr, const unsigned int* seg_off, T* val,
                                   unsigned short* col16) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const unsigned long long base = row_ptr[warp];
  const unsigned int* so = seg_off + (size_t)warp * (nseg + 1);
  for (int s = 0; s < nseg; ++s) {
    unsigned long long pos = base + so[s];
    const unsigned long long slice_end = base + so[s + 1];
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
        if (keep & (1u << e)) { val[w] = x[e]; col16[w] = (unsigned short)(j + e - c0); ++w; }
      pos += total;
    }
    for (unsigned long long w = pos + lane; w < slice_end; w += 32) { val[w] = encode<T>(0.0, false); col16[w] = 0; }
  }
}

// 4 entries of one row slice
template <typename T> struct Entry4;
template <> struct Entry4<float> {
  float4 x; uint2 k;
  __device__ __forceinline__ void load(const float* val, const unsigned short* col, unsigned long long at) {
    x = ldg_stream(reinterpret_cast<const float4*>(val + at));
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(k.x), "=r"(k.y) : "l"(col + at));
  }
  __device__ __forceinline__ void neutral() { x = make_float4(-0.f, -0.f, -0.f, -0.f); k = make_uint2(0u, 0u); }
  __device__ __forceinline__ float get(int e) const { return e == 0 ? x.x : e == 1 ? x.y : e == 2 ? x.z : x.w; }
};
template <> struct Entry4<double> {
  double2 a, b; uint2 k;
  __device__ __forceinline__ void load(const double* val, const unsigned short* col, unsigned long long at) {
    a = ldg_stream(reinterpret_cast<const double2*>(val + at));
    b = ldg_stream(reinterpret_cast<const double2*>(val + at) + 1);
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(k.x), "=r"(k.y) : "l"(col + at));
  }
  __device__ __forceinline__ void neutral() { a = make_double2(-0.0, -0.0); b = a; k = make_uint2(0u, 0u); }
  __device__ __forceinline__ double get(int e) const { return e == 0 ? a.x : e == 1 ? a.y : e == 2 ? b.x : b.y; }
};
__device__ __forceinline__ unsigned int col_of(const uint2& k, int e) {
  return e == 0 ? (k.x & 0xffffu) : e == 1 ? (k.x >> 16) : e == 2 ? (k.y & 0xffffu) : (k.y >> 16);
}

// 2 x 4 rows (rows [lra, lra+4) and [lrb, lrb+4); lrb < 0: none) x the staged segment.  The eight row slices
// are walked together; every lane takes chunks of 4 consecutive entries, so up to 8 x (16 + 8) bytes per lane
// are in flight.  vs holds the segment of v in natural order.
template <typename T>
__device__ __forceinline__ void sparse_rows8(const SparseView& sp, int lra, int lrb, int seg, const double* vs,
                                             double (&accA)[8], double (&accB)[8]) {
  const int lane = threadIdx.x & 31;
  const T* val = reinterpret_cast<const T*>(sp.val);
  unsigned long long beg[8];
  unsigned int n4[8], nmax = 0;
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const int row = (r < 4) ? lra + r : lrb + (r - 4);
    n4[r] = 0; beg[r] = 0;
    if (r < 4 || lrb >= 0) {
      const unsigned long long base = sp.row_ptr[row];
      const unsigned int* so = sp.seg_off + (size_t)row * (sp.nseg + 1) + seg;
      const unsigned int o0 = so[0], o1 = so[1];
      beg[r] = base + o0;
      n4[r] = (o1 - o0) >> 2;
      nmax = max(nmax, n4[r]);
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) { accA[i] = 0.0; accB[i] = 0.0; }
  for (unsigned int c = lane; c < nmax; c += 32) {
    Entry4<T> E[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      if (c < n4[r]) E[r].load(val, sp.col16, beg[r] + 4ull * c);
      else E[r].neutral();
    }
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const double v = vs[col_of(E[r].k, e)];
        double dummyM = 0.0, dummyC = 0.0;
        if (r < 4) apply_elem<false>(E[r].get(e), v, 0.0, accA[r], accA[4 + r], dummyM, dummyC);
        else apply_elem<false>(E[r].get(e), v, 0.0, accB[r - 4], accB[r], dummyM, dummyC);
      }
  }
}

// whole sparse pass of one CTA: same decomposition and partial layout as matvec_phase; a warp works on the
// row slices of two of its row tiles at a time
template <typename T>
__device__ void sparse_phase(const MatView& mv, const Plan& p, const StageArgs& st, const SparseView& sp,
                             double* partM, double* partC, double* vs, double* red_smem) {
  const int sg = blockIdx.x % p.SG, rg = blockIdx.x / p.SG;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int seg = sg; seg < p.NSEG; seg += p.SG) {
    stage_segment<double>(st, p, mv.m, seg, rg == 0, vs, red_smem);  // <double>: natural (unpermuted) order
    for (int rt = rg; rt < p.NRT; rt += 2 * p.RG) {
      const int lra = rt * kRowTile + warp * kRowsPerWarp;
      const int lrb = (rt + p.RG < p.NRT) ? (rt + p.RG) * kRowTile + warp * kRowsPerWarp : -1;
      double accA[8], accB[8];
      sparse_rows8<T>(sp, lra, lrb, seg, vs, accA, accB);
      const double totA = warp_reduce8(accA);
      const double totB = warp_reduce8(accB);
      if ((lane & 3) == 0) {
        const int qv = lane >> 2;  // 0..3: M of row qv, 4..7: C of row qv-4
        double* dst = (qv >> 2) ? partC : partM;
        dst[(size_t)seg * mv.rows_pad + lra + (qv & 3)] = totA;
        if (lrb >= 0) dst[(size_t)seg * mv.rows_pad + lrb + (qv & 3)] = totB;
      }
    }
    __syncthreads();  // vs is re-staged by the next segment pass
  }
}

}  // namespace clp
