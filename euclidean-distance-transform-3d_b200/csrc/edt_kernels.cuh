// edt_kernels.cuh -- sm_100a kernels of the multi-label squared EDT.
//
// What is computed (bit-for-bit the reference's result, see DESIGN.md "Arithmetic"):
//
//   first axis  (reference squared_edt_1d_multi_seg, src/edt.hpp:70-119)
//       out[p] = T[min(kL, kR)],  kL/kR = steps to the nearest voxel of a different label
//       (or the volume face when black_border) on either side, T[k] = fl32(a_k * a_k),
//       a_k = a_{k-1} (+) w in float32 (the reference's sequential adds), +inf if no such
//       voxel, 0 for background.
//   later axes  (reference squared_edt_1d_parabolic_multi_seg + squared_edt_1d_parabolic,
//       src/edt.hpp:168-377) for a voxel i inside a run [a,b) of equal labels:
//       out[i] = min( min_{v in [a,b)} fl32(w2*(i-v)^2 + f[v]),
//                     fl32(w2*(i-a+1)^2) if the run has a low border,
//                     fl32(w2*(b-i)^2)   if the run has a high border )
//       with w2 = fl32(w*w).  The reference finds the same minimum with a lower-envelope
//       scan; only the minimum VALUE matters, so the kernels are
//       free to search it any way they like.
//
// No tensor cores: there is no contraction here, the passes are HBM-bound streaming
// kernels over the label volume and one float32 volume that is transformed in place.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace edtb200 {

// kSqrt / kNegate act in a pass's store; kZeroLabel makes the first-axis pass treat
// background (label 0) as an ordinary label instead of forcing its distance to 0.
enum : int { kSqrt = 1, kNegate = 2, kZeroLabel = 4 };

constexpr int kNoBoundary = 0x3fffffff;

template <int Bytes> struct LabelOf;
template <> struct LabelOf<1> { using type = uint8_t;  using wide = uint32_t; };
template <> struct LabelOf<2> { using type = uint16_t; using wide = uint32_t; };
template <> struct LabelOf<4> { using type = uint32_t; using wide = uint32_t; };
template <> struct LabelOf<8> { using type = uint64_t; using wide = unsigned long long; };

__device__ __forceinline__ float finish_value(float v, bool background, int flags) {
  if (flags & kSqrt) v = __fsqrt_rn(v);                 // np.sqrt / std::sqrt: IEEE-correct
  if ((flags & kNegate) && background) v = -v;          // sdf: f(data) - f(data==0)
  return v;
}

// T[k] for k = 0..count-1 (see header).  One thread: the adds are sequential by definition
// (src/edt.hpp:92-118 accumulates d[i] = d[i-1] + w in float32).
__global__ void step_table_kernel(float w, int count, float* __restrict__ table) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  float a = 0.0f;
  table[0] = 0.0f;
  for (int k = 1; k < count; ++k) {
    a = __fadd_rn(a, w);
    table[k] = __fmul_rn(a, a);
  }
}

// ---------------------------------------------------------------------------------------
// First-axis pass.  One warp per line of `sx` contiguous voxels.
//   sweep 1: 32 labels per step, one per lane (coalesced), neighbour label by shuffle,
//            __ballot -> one 32-bit word of "label changes here" bits B[0..sx] and one of
//            "is background" bits, kept in shared memory (sx/8 bytes per line);
//   sweep 2: per word, the nearest set bit below / above it (warp scan over the words);
//   sweep 3: per voxel, nearest boundary on each side with clz/ffs on its own word, then
//            the table lookup and a coalesced float store.
// B[j] (1 <= j < sx) says labels j-1 and j differ; B[0] and B[sx] are the volume faces.
// ---------------------------------------------------------------------------------------
template <int Bytes>
__global__ void __launch_bounds__(256)
first_axis_kernel(const typename LabelOf<Bytes>::type* __restrict__ labels,
                  float* __restrict__ out, int64_t nlines, int sx,
                  const float* __restrict__ table, int border, int flags) {
  using LT = typename LabelOf<Bytes>::type;
  using WT = typename LabelOf<Bytes>::wide;
  extern __shared__ uint32_t smem_u32[];

  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int warps = blockDim.x >> 5;
  const int nchunks = (sx + 31) >> 5;
  const int nwords = (sx >> 5) + 1;

  uint32_t* bnd = smem_u32 + (size_t)warp * (4 * (size_t)nwords);
  uint32_t* bkg = bnd + nwords;
  int* below = reinterpret_cast<int*>(bkg + nwords);   // nearest set bit in words < c
  int* above = below + nwords;                          // nearest set bit in words > c

  for (int64_t line = (int64_t)blockIdx.x * warps + warp; line < nlines;
       line += (int64_t)gridDim.x * warps) {
    const LT* __restrict__ src = labels + line * sx;
    float* __restrict__ dst = out + line * sx;

    // ---- sweep 1: boundary / background bit words ----
    WT carry = 0;
    for (int c0 = 0; c0 < nchunks; c0 += 4) {
      WT v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int p = ((c0 + u) << 5) + lane;
        v[u] = (p < sx) ? (WT)src[p] : (WT)0;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int c = c0 + u;
        if (c < nchunks) {                       // warp-uniform
          const int p = (c << 5) + lane;
          WT up = __shfl_up_sync(full, v[u], 1);
          if (lane == 0) up = carry;
          carry = __shfl_sync(full, v[u], 31);
          bool edge;
          if (p == 0) edge = border != 0;
          else if (p < sx) edge = (v[u] != up);
          else edge = (p == sx) && (border != 0);
          const uint32_t wb = __ballot_sync(full, edge);
          const uint32_t wz = __ballot_sync(full, (p < sx) && (v[u] == 0));
          if (lane == 0) { bnd[c] = wb; bkg[c] = wz; }
        }
      }
    }
    if ((sx & 31) == 0 && lane == 0) { bnd[nchunks] = border ? 1u : 0u; bkg[nchunks] = 0u; }
    __syncwarp();

    // ---- sweep 2: nearest set bit strictly below / above each word ----
    {
      int run = -1;
      for (int base = 0; base < nwords; base += 32) {
        const int c = base + lane;
        const uint32_t w = (c < nwords) ? bnd[c] : 0u;
        int hi = w ? ((c << 5) + 31 - __clz(w)) : -1;
#pragma unroll
        for (int s = 1; s < 32; s <<= 1) {
          const int o = __shfl_up_sync(full, hi, s);
          if (lane >= s) hi = max(hi, o);
        }
        int excl = __shfl_up_sync(full, hi, 1);
        if (lane == 0) excl = -1;
        excl = max(excl, run);
        if (c < nwords) below[c] = excl;
        run = max(run, __shfl_sync(full, hi, 31));
      }
      int nxt = kNoBoundary;
      for (int base = ((nwords - 1) >> 5) << 5; base >= 0; base -= 32) {
        const int c = base + lane;
        const uint32_t w = (c < nwords) ? bnd[c] : 0u;
        int lo = w ? ((c << 5) + __ffs(w) - 1) : kNoBoundary;
#pragma unroll
        for (int s = 1; s < 32; s <<= 1) {
          const int o = __shfl_down_sync(full, lo, s);
          if (lane + s < 32) lo = min(lo, o);
        }
        int excl = __shfl_down_sync(full, lo, 1);
        if (lane == 31) excl = kNoBoundary;
        excl = min(excl, nxt);
        if (c < nwords) above[c] = excl;
        nxt = min(nxt, __shfl_sync(full, lo, 0));
      }
    }
    __syncwarp();

    // ---- sweep 3: distances ----
    for (int c = 0; c < nchunks; ++c) {
      const int p = (c << 5) + lane;
      const uint32_t w = bnd[c];
      const uint32_t mle = w & (0xffffffffu >> (31 - lane));
      const uint32_t mgt = (lane == 31) ? 0u : (w & (0xfffffffeu << lane));
      const int jl = mle ? ((c << 5) + 31 - __clz(mle)) : below[c];
      const int jr = mgt ? ((c << 5) + __ffs(mgt) - 1) : above[c];
      const int kl = (jl >= 0) ? (p - jl + 1) : kNoBoundary;
      const int kr = (jr != kNoBoundary) ? (jr - p) : kNoBoundary;
      const int k = min(kl, kr);
      if (p < sx) {
        const bool background = (bkg[c] >> lane) & 1u;
        float val = (k >= kNoBoundary) ? __int_as_float(0x7f800000) : __ldg(table + k);
        if (background && !(flags & kZeroLabel)) val = 0.0f;
        dst[p] = finish_value(val, background, flags);
      }
    }
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------------
// Later-axis pass (Y: line stride sx, Z: line stride sx*sy), in place on f.
//
// A CTA owns a tile of 32 adjacent lines (32 consecutive x, i.e. 128 B per row, every
// global access a full coalesced line) times the whole line length n, staged in shared
// memory:  fs[n][32] float32  +  one 32-bit "run starts here" word per (32 rows, line).
// Thread (lane = line, warp = 32-row chunk) builds the word of its chunk from the labels,
// then produces the 32 outputs of its chunk.  For each output the search is a two-sided
// scan outwards over its own run that stops as soon as w2*d^2 alone reaches the best value
// found so far; candidates are evaluated with one fused multiply-add each, which is the
// correctly rounded value of w2*d^2 + f[v] -- what the reference computes in double and
// rounds once (src/edt.hpp:225-230).  Run borders enter as the two closed-form terms.
// ---------------------------------------------------------------------------------------
struct LineGeom {
  int64_t outer_count;     // Y pass: sz            Z pass: 1
  int64_t outer_stride;    // Y pass: sx*sy         Z pass: 0
  int64_t inner_count;     // Y pass: sx            Z pass: sx*sy   (adjacent lines)
  int64_t line_stride;     // Y pass: sx            Z pass: sx*sy
  int n;                   // line length
  int tiles_per_outer;     // ceil(inner_count / 32)
};

template <int Bytes>
__global__ void __launch_bounds__(512)
later_axis_kernel(const typename LabelOf<Bytes>::type* __restrict__ labels,
                  float* __restrict__ f, LineGeom g, float w2,
                  int border_lo, int border_hi, int flags) {
  using LT = typename LabelOf<Bytes>::type;
  extern __shared__ __align__(16) unsigned char smem_raw[];

  const int n = g.n;
  const int nchunks = (n + 31) >> 5;
  float* fs = reinterpret_cast<float*>(smem_raw);                      // [n][32]
  uint32_t* startw = reinterpret_cast<uint32_t*>(fs + (size_t)n * 32); // [nchunks][32]
  uint32_t* zerow = startw + (size_t)nchunks * 32;                     // [nchunks][32]

  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int warps = blockDim.x >> 5;

  const int64_t tile = blockIdx.x;
  const int64_t outer = tile / g.tiles_per_outer;
  const int64_t inner0 = (tile - outer * g.tiles_per_outer) * 32;
  const bool live = (inner0 + lane) < g.inner_count;
  const int64_t base = outer * g.outer_stride + inner0 + lane;
  const int64_t ls = g.line_stride;

  // ---- stage: f tile and run-start words ----
  for (int c = warp; c < nchunks; c += warps) {
    const int i0 = c << 5;
    uint32_t wstart = 0, wzero = 0;
    if (live) {
      LT prev = (i0 > 0) ? labels[base + (int64_t)(i0 - 1) * ls] : (LT)0;
#pragma unroll 8
      for (int r = 0; r < 32; ++r) {
        const int i = i0 + r;
        if (i < n) {
          const int64_t at = base + (int64_t)i * ls;
          const LT here = labels[at];
          fs[(size_t)i * 32 + lane] = f[at];
          if (i > 0 && here != prev) wstart |= (1u << r);
          if (here == 0) wzero |= (1u << r);
          prev = here;
        }
      }
    }
    startw[(size_t)c * 32 + lane] = wstart;
    zerow[(size_t)c * 32 + lane] = wzero;
  }
  __syncthreads();
  if (!live) return;

  // ---- compute ----
  for (int c = warp; c < nchunks; c += warps) {
    const int i0 = c << 5;
    const uint32_t wstart = startw[(size_t)c * 32 + lane];
    const uint32_t wzero = zerow[(size_t)c * 32 + lane];

    // start of the run that is open when this chunk begins (0 = line start)
    int run_lo = 0;
    for (int cc = c - 1; cc >= 0; --cc) {
      const uint32_t w = startw[(size_t)cc * 32 + lane];
      if (w) { run_lo = (cc << 5) + 31 - __clz(w); break; }
    }
    // first run start after this chunk (n = line end)
    int next_hi = n;
    for (int cc = c + 1; cc < nchunks; ++cc) {
      const uint32_t w = startw[(size_t)cc * 32 + lane];
      if (w) { next_hi = (cc << 5) + __ffs(w) - 1; break; }
    }

    for (int r = 0; r < 32; ++r) {
      const int i = i0 + r;
      if (i >= n) break;
      if ((wstart >> r) & 1u) run_lo = i;
      const uint32_t later = (r == 31) ? 0u : (wstart & (0xfffffffeu << r));
      const int run_hi = later ? (i0 + __ffs(later) - 1) : next_hi;   // exclusive

      const int dl = i - run_lo;          // in-run candidates below i
      const int dr = run_hi - 1 - i;      // in-run candidates above i
      float best = fs[(size_t)i * 32 + lane];
      if (run_lo > 0 || border_lo) {
        const float e = (float)(dl + 1);
        best = fminf(best, __fmul_rn(w2, __fmul_rn(e, e)));
      }
      if (run_hi < n || border_hi) {
        const float e = (float)(dr + 1);
        best = fminf(best, __fmul_rn(w2, __fmul_rn(e, e)));
      }
      const int dmax = max(dl, dr);
      float fd = 1.0f;
      for (int d = 1; d <= dmax; ++d, fd += 1.0f) {
        const float t = __fmul_rn(fd, fd);
        if (!(__fmul_rn(w2, t) < best)) break;
        if (d <= dl) best = fminf(best, __fmaf_rn(w2, t, fs[(size_t)(i - d) * 32 + lane]));
        if (d <= dr) best = fminf(best, __fmaf_rn(w2, t, fs[(size_t)(i + d) * 32 + lane]));
      }
      const bool background = (wzero >> r) & 1u;
      f[base + (int64_t)i * ls] = finish_value(best, background, flags);
    }
  }
}

// ---------------------------------------------------------------------------------------
// Later-axis pass for lines too long for a shared-memory tile: one thread per line, lanes
// on adjacent lines (coalesced), reading f_in / labels through L1/L2 and writing f_out
// (out of place, so no tile-wide synchronisation is needed).  Same arithmetic; distances
// of 4096 voxels and more, whose squares are not exact in float32, go through double.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ float parabola_at(float w2, int d, float height) {
  if (d < 4096) {
    const float e = (float)d;
    return __fmaf_rn(w2, __fmul_rn(e, e), height);
  }
  const double e = (double)d;
  return (float)__dadd_rn(__dmul_rn((double)w2, e * e), (double)height);
}

template <int Bytes>
__global__ void __launch_bounds__(128)
later_axis_long_kernel(const typename LabelOf<Bytes>::type* __restrict__ labels,
                       const float* __restrict__ fin, float* __restrict__ fout,
                       LineGeom g, float w2, int border_lo, int border_hi, int flags) {
  using LT = typename LabelOf<Bytes>::type;
  const int64_t lines_per_outer = g.inner_count;
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= lines_per_outer * g.outer_count) return;
  const int64_t outer = gid / lines_per_outer;
  const int64_t base = outer * g.outer_stride + (gid - outer * lines_per_outer);
  const int64_t ls = g.line_stride;
  const int n = g.n;

  int run_lo = 0;
  int run_hi = 0;            // exclusive end of the current run; recomputed when i reaches it
  LT mine = 0;
  for (int i = 0; i < n; ++i) {
    if (i == run_hi) {
      run_lo = i;
      mine = labels[base + (int64_t)i * ls];
      int j = i + 1;
      while (j < n && labels[base + (int64_t)j * ls] == mine) ++j;
      run_hi = j;
    }
    const int dl = i - run_lo;
    const int dr = run_hi - 1 - i;
    float best = fin[base + (int64_t)i * ls];
    if (run_lo > 0 || border_lo) best = fminf(best, parabola_at(w2, dl + 1, 0.0f));
    if (run_hi < n || border_hi) best = fminf(best, parabola_at(w2, dr + 1, 0.0f));
    const int dmax = max(dl, dr);
    for (int d = 1; d <= dmax; ++d) {
      if (!(parabola_at(w2, d, 0.0f) < best)) break;
      if (d <= dl) best = fminf(best, parabola_at(w2, d, fin[base + (int64_t)(i - d) * ls]));
      if (d <= dr) best = fminf(best, parabola_at(w2, d, fin[base + (int64_t)(i + d) * ls]));
    }
    fout[base + (int64_t)i * ls] = finish_value(best, mine == 0, flags);
  }
}

}  // namespace edtb200
