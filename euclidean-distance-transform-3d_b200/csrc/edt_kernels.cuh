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
#include <cuda.h>          // CUtensorMap (type only; the encoder is fetched at run time)
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
// First-axis pass, register-resident variant for lines of sx <= 128*K voxels with sx % 4 == 0
// (rows then start on 16-byte boundaries for the float4 stores and on 4*Bytes boundaries for
// the vector label loads).  One warp per line; a lane owns 4 consecutive voxels of each
// 128-voxel block, so every global access is a full-width vector (4*Bytes per lane in,
// 16 bytes per lane out).  Boundaries are found in registers: per lane a 4-bit mask per
// block, then a warp max-scan (nearest boundary at or below) and a warp min-scan (nearest
// boundary above) with carries across the K blocks.  The step table lives in shared memory.
// ---------------------------------------------------------------------------------------
template <int Bytes> struct Vec4Labels;
template <> struct Vec4Labels<1> {
  static __device__ __forceinline__ void load(const uint8_t* p, uint32_t v[4]) {
    const uint32_t w = __ldg(reinterpret_cast<const uint32_t*>(p));
    v[0] = w & 0xffu; v[1] = (w >> 8) & 0xffu; v[2] = (w >> 16) & 0xffu; v[3] = w >> 24;
  }
};
template <> struct Vec4Labels<2> {
  static __device__ __forceinline__ void load(const uint16_t* p, uint32_t v[4]) {
    const uint2 w = __ldg(reinterpret_cast<const uint2*>(p));
    v[0] = w.x & 0xffffu; v[1] = w.x >> 16; v[2] = w.y & 0xffffu; v[3] = w.y >> 16;
  }
};
template <> struct Vec4Labels<4> {
  static __device__ __forceinline__ void load(const uint32_t* p, uint32_t v[4]) {
    const uint4 w = __ldg(reinterpret_cast<const uint4*>(p));
    v[0] = w.x; v[1] = w.y; v[2] = w.z; v[3] = w.w;
  }
};
template <> struct Vec4Labels<8> {
  static __device__ __forceinline__ void load(const uint64_t* p, unsigned long long v[4]) {
    const ulonglong2 a = __ldg(reinterpret_cast<const ulonglong2*>(p));
    const ulonglong2 b = __ldg(reinterpret_cast<const ulonglong2*>(p) + 1);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
  }
};

// Plain = true: squared EDT, background forced to 0, no sqrt / sign (the hot configuration);
// Plain = false: behaviour selected by `flags` at run time.
template <int Bytes, int K, bool Plain>
__global__ void __launch_bounds__(256)
first_axis_vec_kernel(const typename LabelOf<Bytes>::type* __restrict__ labels,
                      float* __restrict__ out, int64_t nlines, int sx,
                      const float* __restrict__ table, int border, int flags) {
  using LT = typename LabelOf<Bytes>::type;
  using WT = typename LabelOf<Bytes>::wide;
  extern __shared__ float table_s[];                 // T[0..sx], then +inf at sx + 1
  for (int i = threadIdx.x; i <= sx; i += blockDim.x) table_s[i] = __ldg(table + i);
  if (threadIdx.x == 0) table_s[sx + 1] = __int_as_float(0x7f800000);
  __syncthreads();

  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int warps = blockDim.x >> 5;
  const unsigned lt_mask = (1u << lane) - 1u;
  const unsigned gt_mask = (lane == 31) ? 0u : (0xfffffffeu << lane);
  const bool keep_background = Plain ? false : (flags & kZeroLabel) != 0;

  for (int64_t line = (int64_t)blockIdx.x * warps + warp; line < nlines;
       line += (int64_t)gridDim.x * warps) {
    const LT* __restrict__ src = labels + line * sx;
    float* __restrict__ dst = out + line * sx;

    // ---- labels -> per-block 4-bit "label changes here" / "is background" masks ----
    WT v[K][4];
#pragma unroll
    for (int b = 0; b < K; ++b) {
      const int q0 = (b << 7) + (lane << 2);
      if (q0 < sx) Vec4Labels<Bytes>::load(src + q0, v[b]);
      else { v[b][0] = v[b][1] = v[b][2] = v[b][3] = 0; }
    }
    uint32_t edges = 0, zeros = 0;                   // 4 bits per block
    uint32_t occ[K];                                 // lanes owning at least one boundary
#pragma unroll
    for (int b = 0; b < K; ++b) {
      const int q0 = (b << 7) + (lane << 2);
      WT up = __shfl_up_sync(full, v[b][3], 1);
      if (b > 0) {                                   // compile-time: warp-uniform
        const WT tail = __shfl_sync(full, v[b > 0 ? b - 1 : 0][3], 31);
        if (lane == 0) up = tail;
      }
      uint32_t m = 0, z = 0;
      if (q0 < sx) {
        if ((q0 == 0) ? (border != 0) : (v[b][0] != up)) m |= 1u;
        if (v[b][1] != v[b][0]) m |= 2u;
        if (v[b][2] != v[b][1]) m |= 4u;
        if (v[b][3] != v[b][2]) m |= 8u;
        if (v[b][0] == 0) z |= 1u;
        if (v[b][1] == 0) z |= 2u;
        if (v[b][2] == 0) z |= 4u;
        if (v[b][3] == 0) z |= 8u;
      }
      occ[b] = __ballot_sync(full, m != 0);
      edges |= m << (4 * b);
      zeros |= z << (4 * b);
    }

    // ---- nearest boundary strictly below / above the lane's 4 voxels, per block ----
    // The ballots give the boundary-owning lanes; one shuffle fetches that lane's masks.
    int kl_in[K], kr_in[K];          // kL of the voxel just below q0, kR seed for q0 + 3
    {
      uint32_t prev_occ = 0; int prev_b = 0;
#pragma unroll
      for (int b = 0; b < K; ++b) {
        const int q0 = (b << 7) + (lane << 2);
        const uint32_t mine = occ[b] & lt_mask;
        const uint32_t pick = mine ? mine : prev_occ;
        const int pb = mine ? b : prev_b;
        const int sl = 31 - __clz(pick | 1u);
        const uint32_t nib = (__shfl_sync(full, edges, sl) >> (4 * pb)) & 15u;
        const int pos = (pb << 7) + (sl << 2) + 31 - __clz(nib | 1u);
        kl_in[b] = pick ? (q0 - pos) : kNoBoundary;          // = kL(q0 - 1) + ... see below
        if (occ[b]) { prev_occ = occ[b]; prev_b = b; }
      }
      uint32_t next_occ = 0; int next_b = 0;
#pragma unroll
      for (int b = K - 1; b >= 0; --b) {
        const int q0 = (b << 7) + (lane << 2);
        const uint32_t mine = occ[b] & gt_mask;
        const uint32_t pick = mine ? mine : next_occ;
        const int pb = mine ? b : next_b;
        const int sl = __ffs(pick | 0x80000000u) - 1;
        const uint32_t nib = (__shfl_sync(full, edges, sl) >> (4 * pb)) & 15u;
        const int pos = (pb << 7) + (sl << 2) + __ffs(nib | 8u) - 1;
        const int far = border ? (sx - (q0 + 3)) : kNoBoundary;
        kr_in[b] = pick ? (pos - (q0 + 3)) : far;
        if (occ[b]) { next_occ = occ[b]; next_b = b; }
      }
    }

    // ---- distances (chains over the 4 voxels), table lookup, vector store ----
#pragma unroll
    for (int b = 0; b < K; ++b) {
      const int q0 = (b << 7) + (lane << 2);
      if (q0 >= sx) continue;
      const uint32_t m = (edges >> (4 * b)) & 15u;
      const uint32_t z = (zeros >> (4 * b)) & 15u;
      // kL(q) = q - (nearest boundary position <= q) + 1:  1 at a boundary, else previous + 1
      int kl[4], kr[4];
      int run = kl_in[b];                    // q0 - pos: kL the voxel q0 would have without its own bit
#pragma unroll
      for (int e = 0; e < 4; ++e) { run = ((m >> e) & 1u) ? 1 : run + 1; kl[e] = run; }
      // kR(q) = (nearest boundary position > q) - q:  1 if q + 1 is a boundary, else next + 1
      run = kr_in[b];
      kr[3] = run;
#pragma unroll
      for (int e = 2; e >= 0; --e) { run = ((m >> (e + 1)) & 1u) ? 1 : run + 1; kr[e] = run; }
      float r[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        int k = min(min(kl[e], kr[e]), sx + 1);          // sx + 1 -> +inf (no boundary at all)
        const bool background = (z >> e) & 1u;
        if (background && !keep_background) k = 0;        // T[0] = 0
        float val = table_s[k];
        if (!Plain) val = finish_value(val, background, flags);
        r[e] = val;
      }
      *reinterpret_cast<float4*>(dst + q0) = make_float4(r[0], r[1], r[2], r[3]);
    }
  }
}

// ---------------------------------------------------------------------------------------
// Later-axis pass (Y: line stride sx, Z: line stride sx*sy), in place on f.
//
// A CTA owns a tile of TX adjacent lines (TX = 32, 16 or 8 consecutive x: 128/64/32 B per
// row, so every global access is made of whole 32-byte sectors) times the whole line length
// n, staged in shared memory:
//     fs[n][TX] float32   the distance tile, fetched by the TMA unit (cp.async.bulk.tensor over
//                         a 3-D tensor map of the volume; ragged tiles are zero-filled by the
//                         hardware) while the threads work on the labels;
//     startw[n/32][TX]    one 32-bit "a run of equal labels starts here" word per 32 rows;
//     hull_own/in[n/32][TX]  lower-envelope membership bits (the scan's vertex stack, 1 bit/voxel);
//     sq[n+2]             w2*e^2, the closed-form border terms.
// Thread (x = lane % TX, chunk = 32 consecutive rows) turns its label column into a run-start
// word, then produces the outputs of the runs that START in its chunk:
//   * runs of length one (the common case in dense segmentations, the only case for iid
//     labels): out = min(f, w2) -- a predicated straight-line loop over the 32 rows;
//   * every other run: Felzenszwalb-Huttenlocher lower envelope over the run, restated for
//     one thread per run: vertices with f = +inf are not sites; a vertex is dropped when
//     its intersection with the newcomer lies at or left of its intersection with the
//     vertex below it (compared by cross-multiplication in double, no division); the
//     read-out walks the hull and evaluates each candidate with one fused multiply-add,
//     which is the correctly rounded w2*d^2 + f[v] the reference computes in double and
//     rounds once (src/edt.hpp:225-230); run borders enter as sq[] terms.
// ---------------------------------------------------------------------------------------
struct LineGeom {
  int64_t outer_count;     // Y pass: sz            Z pass: 1
  int64_t outer_stride;    // Y pass: sx*sy         Z pass: 0
  int64_t inner_count;     // Y pass: sx            Z pass: sx*sy   (adjacent lines)
  int64_t line_stride;     // Y pass: sx            Z pass: sx*sy
  int n;                   // line length
  int tiles_per_outer;     // ceil(inner_count / TX)
};

struct TileBoxes {
  int box_rows;   // rows per TMA box (<= 256)
  int nboxes;     // boxes per tile; box_rows * nboxes >= n
};

__device__ __forceinline__ uint32_t smem_addr(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, unsigned parity) {
  unsigned ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_addr(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a tile load that never completes (bad tensor map) traps instead of hanging.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  for (unsigned spin = 0; spin < (1u << 26); ++spin)
    if (mbar_try_wait(bar, parity)) return;
  __trap();
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, int c0, int c1, int c2,
                                            uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
      ::"r"(smem_addr(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2),
        "r"(smem_addr(bar))
      : "memory");
}

// Hull membership bits of one line: bit (pos & 31) of word (pos >> 5) says "pos is a vertex of
// the lower envelope".  Two arrays keep writers apart without atomics: a run's bits go to
// `own` for the rows of the chunk it starts in and to `in` for the rows of later chunks (at
// most one run can enter a chunk from below, and the runs starting in a chunk are all
// processed, one after the other, by that chunk's thread).
template <int TX>
struct HullBits {
  uint32_t* own;   // &hull_own[0][x], row stride TX
  uint32_t* in;    // &hull_in[0][x]
  int ca;          // chunk in which the current run starts
  __device__ __forceinline__ uint32_t load(int wi) const { return (wi == ca ? own : in)[wi * TX]; }
  __device__ __forceinline__ void store(int wi, uint32_t v) const { (wi == ca ? own : in)[wi * TX] = v; }
};

// Smallest hull vertex position in (after, b), or -1.
template <int TX>
__device__ __forceinline__ int hull_next_vertex(const HullBits<TX> hb, int after, int b) {
  const int pos = after + 1;
  if (pos >= b) return -1;
  const int wlast = (b - 1) >> 5;
  int wi = pos >> 5;
  uint32_t m = hb.load(wi) & (0xffffffffu << (pos & 31));
  while (m == 0u) {
    if (wi == wlast) return -1;
    ++wi;
    m = hb.load(wi);
  }
  const int v = (wi << 5) + __ffs(m) - 1;
  return v < b ? v : -1;
}

// Lower envelope of the parabolas rooted at the finite samples of one run [a, b) of one line.
//   fcol : &fs[0][x] (row stride TX floats)
//   out  : byte address of the line's row 0 in global memory, `pitch` bytes between rows
template <int TX, bool Epilogue>
__device__ __forceinline__ void envelope_run(const float* __restrict__ fcol, HullBits<TX> hb, int a, int b,
                                             float w2f, bool lo_border, bool hi_border,
                                             const float* __restrict__ sq, char* __restrict__ out,
                                             size_t pitch, bool background, int flags) {
  const float inf = __int_as_float(0x7f800000);
  const double w2 = (double)w2f;
  const int wa = a >> 5;
  const uint32_t amask = 0xffffffffu << (a & 31);

  // ---- build: drop every vertex hidden by its neighbours ----
  int cw = wa;                   // word currently held in `cur`
  uint32_t cur = hb.load(wa);    // bits of earlier runs (positions < a) are left alone
  int cnt = 0;                   // vertices on the hull
  int q = 0, p = 0;              // top vertex and the one below it
  double fq = 0.0;               // f[q]
  double num = 0.0, den = 1.0;   // s(p,q) = num / (2*w2*den), kept as the pair (num, den)
  for (int r = a; r < b; ++r) {
    if ((r >> 5) != cw) { hb.store(cw, cur); cw = r >> 5; cur = 0u; }
    const float frf = fcol[r * TX];
    if (!(frf < inf)) continue;                       // +inf: not a site
    const double fr = (double)frf;
    double num_r = 0.0, den_r = 1.0;
    while (cnt >= 1) {
      den_r = (double)(r - q);
      num_r = (fr - fq) + w2 * (den_r * (double)(r + q));
      if (cnt == 1) break;                            // the bottom vertex is never dropped
      if (num_r * den > num * den_r) break;           // s(q,r) > s(p,q): q stays
      {                                               // drop q
        const int wq = q >> 5;
        const uint32_t bit = 1u << (q & 31);
        if (wq == cw) cur &= ~bit; else hb.store(wq, hb.load(wq) & ~bit);
      }
      --cnt;
      q = p;
      fq = (double)fcol[q * TX];
      if (cnt >= 2) {                                 // vertex below the new top
        int wi = q >> 5;
        uint32_t m = (wi == cw ? cur : hb.load(wi)) & ((1u << (q & 31)) - 1u);
        if (wi == wa) m &= amask;
        while (m == 0u) { --wi; m = hb.load(wi); if (wi == wa) m &= amask; }
        p = (wi << 5) + 31 - __clz(m);
        den = (double)(q - p);
        num = (fq - (double)fcol[p * TX]) + w2 * (den * (double)(q + p));
      }
    }
    cur |= 1u << (r & 31);
    ++cnt;
    p = q; num = num_r; den = den_r;
    q = r; fq = fr;
  }
  hb.store(cw, cur);

  // ---- read out: walk the hull, one fused multiply-add per candidate ----
  int v = hull_next_vertex<TX>(hb, a - 1, b);
  int v1 = (v >= 0) ? hull_next_vertex<TX>(hb, v, b) : -1;
  float fv = inf, fv1 = inf, dv = 0.0f, dv1 = 0.0f;   // dv = (float)(i - v)
  if (v >= 0) { fv = fcol[v * TX]; dv = (float)(a - v); }
  if (v1 >= 0) { fv1 = fcol[v1 * TX]; dv1 = (float)(a - v1); }
  char* dst = out + (size_t)a * pitch;
  for (int i = a; i < b; ++i) {
    float best = inf;
    if (v >= 0) {
      best = __fmaf_rn(w2f, __fmul_rn(dv, dv), fv);
      while (v1 >= 0) {
        const float cand = __fmaf_rn(w2f, __fmul_rn(dv1, dv1), fv1);
        if (!(cand <= best)) break;
        best = cand; v = v1; fv = fv1; dv = dv1;
        v1 = hull_next_vertex<TX>(hb, v1, b);
        if (v1 >= 0) { fv1 = fcol[v1 * TX]; dv1 = (float)(i - v1); }
      }
      dv += 1.0f; dv1 += 1.0f;
    }
    if (lo_border) best = fminf(best, sq[i - a + 1]);
    if (hi_border) best = fminf(best, sq[b - i]);
    if (Epilogue) best = finish_value(best, background, flags);   // a run has one label
    *reinterpret_cast<float*>(dst) = best;
    dst += pitch;
  }
}

template <int Bytes, int TX, bool Epilogue, bool UseTMA>
__global__ void __launch_bounds__(512)
later_axis_tile_kernel(const __grid_constant__ CUtensorMap fmap,
                       const typename LabelOf<Bytes>::type* __restrict__ labels,
                       float* __restrict__ f, LineGeom g, TileBoxes tb, float w2,
                       int border_lo, int border_hi, int flags) {
  using LT = typename LabelOf<Bytes>::type;
  extern __shared__ __align__(128) unsigned char smem_tile[];
  constexpr int SUBS = 32 / TX;                       // chunks handled side by side by one warp

  const int n = g.n;
  const int nchunks = (n + 31) >> 5;
  const int rows_alloc = UseTMA ? tb.box_rows * tb.nboxes : n;
  float* fs = reinterpret_cast<float*>(smem_tile);                               // [rows_alloc][TX]
  uint32_t* startw = reinterpret_cast<uint32_t*>(fs + (size_t)rows_alloc * TX);  // [nchunks][TX]
  uint32_t* zerow = startw + (size_t)nchunks * TX;                               // [nchunks][TX]
  uint32_t* hull_own = zerow + (size_t)nchunks * TX;                             // [nchunks][TX]
  uint32_t* hull_in = hull_own + (size_t)nchunks * TX;                           // [nchunks][TX]
  float* sq = reinterpret_cast<float*>(hull_in + (size_t)nchunks * TX);          // [n + 2]
  uint64_t* bar = reinterpret_cast<uint64_t*>(sq + ((n + 2 + 1) & ~1));

  const int lane = threadIdx.x & 31;
  const int x = lane & (TX - 1);
  const int chunk0 = (threadIdx.x >> 5) * SUBS + (lane / TX);
  const int chunk_step = (blockDim.x >> 5) * SUBS;

  const int64_t tile = blockIdx.x;
  const int64_t outer = tile / g.tiles_per_outer;
  const int64_t inner0 = (tile - outer * g.tiles_per_outer) * TX;
  const bool live = (inner0 + x) < g.inner_count;
  // CTA-uniform tile origin + 32-bit element offsets (the host guarantees n * line_stride < 2^32)
  const LT* __restrict__ tl = labels + (outer * g.outer_stride + inner0);
  float* __restrict__ tf = f + (outer * g.outer_stride + inner0);
  const uint32_t ls = (uint32_t)g.line_stride;

  if (UseTMA && threadIdx.x == 0) {
    mbar_init(bar, 1);
    mbar_expect_tx(bar, (unsigned)rows_alloc * (unsigned)(TX * sizeof(float)));
    for (int bx = 0; bx < tb.nboxes; ++bx)
      tma_load_3d(fs + (size_t)bx * tb.box_rows * TX, &fmap, (int)inner0, bx * tb.box_rows, (int)outer, bar);
  }

  // ---- labels -> run-start / background words; border-term table; (plain loads of f) ----
  for (int i = threadIdx.x; i < n + 2; i += blockDim.x) {
    const float e = (float)i;
    sq[i] = __fmul_rn(w2, __fmul_rn(e, e));
  }
  for (int c = chunk0; c < nchunks; c += chunk_step) {
    const int i0 = c << 5;
    uint32_t wstart = 0, wzero = 0;
    if (live) {
      uint32_t idx = (uint32_t)i0 * ls + (uint32_t)x;
      LT prev = (i0 > 0) ? tl[idx - ls] : (LT)0;
      if (i0 + 32 <= n) {
#pragma unroll
        for (int r = 0; r < 32; ++r) {
          const LT here = tl[idx];
          if (!UseTMA) fs[(size_t)(i0 + r) * TX + x] = tf[idx];
          idx += ls;
          if (here != prev) wstart |= (1u << r);
          if (Epilogue && here == 0) wzero |= (1u << r);
          prev = here;
        }
      } else {
        for (int r = 0; r < n - i0; ++r) {
          const LT here = tl[idx];
          if (!UseTMA) fs[(size_t)(i0 + r) * TX + x] = tf[idx];
          idx += ls;
          if (here != prev) wstart |= (1u << r);
          if (Epilogue && here == 0) wzero |= (1u << r);
          prev = here;
        }
        wstart |= 1u << (n - i0);          // pretend a run starts at row n (line end)
      }
      if (i0 == 0) wstart |= 1u;           // a run starts at row 0 by definition
    }
    startw[(size_t)c * TX + x] = wstart;
    if (Epilogue) zerow[(size_t)c * TX + x] = wzero;
    hull_own[(size_t)c * TX + x] = 0u;
    hull_in[(size_t)c * TX + x] = 0u;
  }
  __syncthreads();         // words, table (and plain-loaded tile) visible; orders the mbarrier init
  if (UseTMA) mbar_wait(bar, 0);       // float tile has landed
  if (!live) return;

  for (int c = chunk0; c < nchunks; c += chunk_step) {
    const int i0 = c << 5;
    const int rows = min(32, n - i0);
    const uint32_t wstart = startw[(size_t)c * TX + x];
    const uint32_t wzero = Epilogue ? zerow[(size_t)c * TX + x] : 0u;
    // bit r of `nextw`: a run starts at row i0 + r + 1 (the line end counts as a start)
    uint32_t ext = 1u;
    if (i0 + 32 < n) ext = startw[(size_t)(c + 1) * TX + x] & 1u;
    const uint32_t nextw = (wstart >> 1) | (ext << 31);
    uint32_t single = wstart & nextw;                    // runs of length one
    if (!border_lo && c == 0) single &= ~1u;             // rows lacking a border term go the long way
    if (!border_hi && i0 + 32 >= n) single &= ~(1u << (n - 1 - i0));

    const float* fp0 = fs + (size_t)i0 * TX + x;
    char* const op0 = reinterpret_cast<char*>(tf + ((uint32_t)i0 * ls + (uint32_t)x));
    const size_t pitch = (size_t)ls * sizeof(float);

    // (1) runs of length one: min(f, w2); value computed unconditionally, store predicated
    if (rows == 32) {
      char* op = op0;
#pragma unroll
      for (int r = 0; r < 32; ++r) {
        float v = fminf(fp0[r * TX], w2);
        if (Epilogue) v = finish_value(v, (wzero >> r) & 1u, flags);
        if (single & (1u << r)) *reinterpret_cast<float*>(op) = v;
        op += pitch;
      }
    } else {
      for (int r = 0; r < rows; ++r) {
        float v = fminf(fp0[r * TX], w2);
        if (Epilogue) v = finish_value(v, (wzero >> r) & 1u, flags);
        if (single & (1u << r)) *reinterpret_cast<float*>(op0 + (size_t)r * pitch) = v;
      }
    }

    // (2) every other run that starts in this chunk: lower envelope over the whole run
    uint32_t starts = wstart & ~single & (rows == 32 ? 0xffffffffu : ((1u << rows) - 1u));
    if (starts) {
      // first run start after this chunk (line end if none)
      int next_hi = min(n, i0 + 32);
      if (!ext) {
        next_hi = n;
        for (int cc = c + 1; cc < nchunks; ++cc) {
          const uint32_t w = startw[(size_t)cc * TX + x];
          if (w) { next_hi = min(n, (cc << 5) + __ffs(w) - 1); break; }
        }
      }
      char* const line0 = reinterpret_cast<char*>(tf + x);
      while (starts) {
        const int r0 = __ffs(starts) - 1;
        starts &= starts - 1u;
        const int a = i0 + r0;
        const uint32_t mhi = nextw & (0xffffffffu << r0);
        const int b = mhi ? (i0 + __ffs(mhi)) : next_hi;              // exclusive
        HullBits<TX> hb;
        hb.own = hull_own + x; hb.in = hull_in + x; hb.ca = c;
        envelope_run<TX, Epilogue>(fs + x, hb, a, b, w2, a > 0 || border_lo, b < n || border_hi, sq,
                                   line0, pitch, (wzero >> r0) & 1u, flags);
      }
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
