// stage.cu - source staging kernels: b-spline prefilter (line-parallel recursive filter),
// brace, cubemap internal-representation support. They turn an uploaded raster into the
// coefficient container the render kernel gathers from, with the reference's arithmetic:
//   zimt::prefilter / iir_filter         zimt/prefilter.h:125-198, zimt/recursive.h:321-729
//   spherical_prefilter                   environment.h:356-522
//   bracer                                zimt/brace.h:151-338
//   cubemap_t::fill_support / prefilter   cubemap.h:607-946
#include <algorithm>

#include "eu_device.cuh"
#include "kernels.h"

// ---- the recursive filter on one line, addressed through an accessor ----------------------
// Acc::operator()(n) -> float& of element n of the line.
template <typename Acc>
__device__ __forceinline__ float iir_icc(const IirDev& f, Acc& c, int M, int k) {  // recursive.h:321-583
  float z = f.pole[k], zn, z2n, iz, Sum;
  int n, hz = f.horizon[k];
  switch (f.bc) {
    case EU_BC_MIRROR:
      if (hz < M) {
        zn = z; Sum = c(0);
        for (n = 1; n < hz; n++) { Sum += zn * c(n); zn *= z; }
      } else {
        zn = z; iz = 1.0f / z;
        z2n = f.pole_pow[k];
        Sum = c(0) + z2n * c(M - 1);
        z2n *= z2n * iz;
        for (n = 1; n <= M - 2; n++) { Sum += (zn + z2n) * c(n); zn *= z; z2n *= iz; }
        Sum /= (1.0f - zn * zn);
      }
      return Sum;
    case EU_BC_NATURAL:
      if (hz < M) {
        float c02 = c(0) + c(0);
        zn = z; Sum = c(0);
        for (n = 1; n < hz; n++) { Sum += zn * (c02 - c(n)); zn *= z; }
        return Sum;
      } else {
        zn = z; iz = 1.0f / z;
        z2n = f.pole_pow[k];
        Sum = ((1.0f + z) / (1.0f - z)) * (c(0) - z2n * c(M - 1));
        z2n *= z2n * iz;
        for (n = 1; n <= M - 2; n++) { Sum -= (zn - z2n) * c(n); zn *= z; z2n *= iz; }
        return Sum / (1.0f - zn * zn);
      }
    case EU_BC_REFLECT:
      if (hz < M) {
        zn = z; Sum = c(0);
        for (n = 0; n < hz; n++) { Sum += zn * c(n); zn *= z; }
        return Sum;
      } else {
        zn = z; iz = 1.0f / z;
        z2n = f.pole_pow[k];
        Sum = 0.0f;
        for (n = 0; n < M - 1; n++) { Sum += (zn + z2n) * c(n); zn *= z; z2n *= iz; }
        Sum += (zn + z2n) * c(n);
        return c(0) + Sum / (1.0f - zn * zn);
      }
    default:  // EU_BC_PERIODIC
      if (hz < M) {
        zn = z; Sum = c(0);
        for (n = M - 1; n > (M - hz); n--) { Sum += zn * c(n); zn *= z; }
      } else {
        zn = z; Sum = c(0);
        for (n = M - 1; n > 0; n--) { Sum += zn * c(n); zn *= z; }
        Sum /= (1.0f - zn);
      }
      return Sum;
  }
}

template <typename Acc>
__device__ __forceinline__ float iir_iacc(const IirDev& f, Acc& c, int M, int k) {
  float z = f.pole[k], zn, Sum;
  switch (f.bc) {
    case EU_BC_MIRROR: return (z / (z * z - 1.0f)) * (c(M - 1) + z * c(M - 2));
    case EU_BC_NATURAL: return -(z / ((1.0f - z) * (1.0f - z))) * (c(M - 1) - z * c(M - 2));
    case EU_BC_REFLECT: return c(M - 1) / (1.0f - 1.0f / z);
    default:
      if (f.horizon[k] < M) {
        zn = z; Sum = c(M - 1) * z;
        for (int n = 0; n < f.horizon[k]; n++) { zn *= z; Sum += zn * c(n); }
        Sum = -Sum;
      } else {
        zn = z; Sum = c(M - 1);
        for (int n = 0; n < M - 1; n++) { Sum += zn * c(n); zn *= z; }
        Sum = z * Sum / (zn - 1.0f);
      }
      return Sum;
  }
}

// solve_gain_inlined (recursive.h:631-729) on one line, in place. The recursion itself is serial,
// but the loads are not: a sweep is software-pipelined in chunks of U elements - the loads of the
// next chunk are issued before the dependent chain runs over the current one, so every thread
// keeps U..2U loads in flight all the time (left to the compiler, the in-place stores would
// serialise the loads behind them). Elements start, start+DIR, ... (count of them); step(v) is one
// step of the recursion. Full chunks run without bounds checks: with few lines per SM (the
// row-resident x sweep) the instruction count per element is what bounds the kernel.
template <int U, int DIR, typename Acc, typename Step>
__device__ __forceinline__ void iir_sweep(Acc& c, int start, int count, Step step) {
  float a[U], b[U];
  int j0 = 0;
  if (count >= 2 * U) {
#pragma unroll
    for (int u = 0; u < U; u++) a[u] = c(start + DIR * u);
    for (; j0 + 3 * U <= count; j0 += 2 * U) {  // chunks j0 (in a), j0+U and j0+2U are complete
      const int n = start + DIR * j0;
#pragma unroll
      for (int u = 0; u < U; u++) b[u] = c(n + DIR * (U + u));
#pragma unroll
      for (int u = 0; u < U; u++) a[u] = step(a[u]);
#pragma unroll
      for (int u = 0; u < U; u++) c(n + DIR * u) = a[u];
#pragma unroll
      for (int u = 0; u < U; u++) a[u] = c(n + DIR * (2 * U + u));
#pragma unroll
      for (int u = 0; u < U; u++) b[u] = step(b[u]);
#pragma unroll
      for (int u = 0; u < U; u++) c(n + DIR * (U + u)) = b[u];
    }
    {  // chunk j0 is loaded and complete
      const int n = start + DIR * j0;
#pragma unroll
      for (int u = 0; u < U; u++) a[u] = step(a[u]);
#pragma unroll
      for (int u = 0; u < U; u++) c(n + DIR * u) = a[u];
      j0 += U;
    }
  }
  for (; j0 < count; j0++) {
    const int n = start + DIR * j0;
    c(n) = step(c(n));
  }
}

template <int U, typename Acc>
__device__ __forceinline__ void iir_line(const IirDev& f, Acc& c, int M) {
  if (M == 1 || f.npoles < 1) return;
  for (int k = 0; k < f.npoles; k++) {
    const float p = f.pole[k], g = f.gain;
    float X = iir_icc(f, c, M, k);
    if (k == 0) X = g * X;
    c(0) = X;
    if (k == 0)
      iir_sweep<U, 1>(c, 1, M - 1, [&](float v) { X = g * v + p * X; return X; });
    else
      iir_sweep<U, 1>(c, 1, M - 1, [&](float v) { X = v + p * X; return X; });
    X = iir_iacc(f, c, M, k);
    c(M - 1) = X;
    iir_sweep<U, -1>(c, M - 2, M - 1, [&](float v) { X = p * (X - v); return X; });
  }
}

struct StrideAcc {
  float* base;
  ptrdiff_t st;
  __device__ __forceinline__ float& operator()(int n) { return base[(ptrdiff_t)n * st]; }
  __device__ __forceinline__ float* ptr(int n) { return base + (ptrdiff_t)n * st; }
  __device__ __forceinline__ bool linear(int, int) { return true; }  // elements n0..n1 are evenly spaced
  __device__ __forceinline__ ptrdiff_t delta(int) { return st; }     // floats from element n to n+1
};
// [left-half column top->bottom ; right-half column bottom->top], environment.h:425-447
struct PoleAcc {
  float* up;    // (x, 0)
  float* down;  // (x + w/2, h-1)
  ptrdiff_t st;
  int h;
  __device__ __forceinline__ float& operator()(int n) {
    return n < h ? up[(ptrdiff_t)n * st] : down[-(ptrdiff_t)(n - h) * st];
  }
  __device__ __forceinline__ float* ptr(int n) { return n < h ? up + (ptrdiff_t)n * st : down - (ptrdiff_t)(n - h) * st; }
  __device__ __forceinline__ bool linear(int n0, int n1) { return (n0 < h) == (n1 < h); }
  __device__ __forceinline__ ptrdiff_t delta(int n) { return n < h ? st : -st; }
};

// ---- lines along y through a shared-memory prefetch ring -----------------------------------
// One thread per column float, as before, but the loads no longer pass through registers: every
// thread keeps IIRY_D elements of its line in flight as 4-byte cp.async copies into its own column
// of a [IIRY_D][blockDim.x] ring (committed in groups of IIRY_G rows), waits for the oldest group,
// runs the recursion over it, stores the results and refills the slots. A thread only ever reads
// what it copied itself, so cp.async.wait_group is all the synchronisation there is. Bytes in
// flight per SM = lines x IIRY_D x 4 - set by shared memory, not by the register file, which is
// what held the register-prefetching version at a third of the HBM roofline.
#define IIRY_D 64
#define IIRY_G 8
#define IIRY_THREADS 64
__device__ __forceinline__ void cp_async4(float* smem_dst, const float* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc)
               : "memory");
}
template <int DIR, typename Acc, typename Step>
__device__ __forceinline__ void iir_sweep_ring(Acc& c, int start, int count, Step step, float* ring) {
  // element j of the sweep = c(start + DIR * j); its ring slot = ring[(j % IIRY_D) * IIRY_THREADS]
#pragma unroll 1
  for (int g0 = 0; g0 < IIRY_D; g0 += IIRY_G) {
#pragma unroll
    for (int u = 0; u < IIRY_G; u++)
      if (g0 + u < count) cp_async4(ring + (g0 + u) * IIRY_THREADS, c.ptr(start + DIR * (g0 + u)));
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
#pragma unroll 1
  for (int j0 = 0; j0 < count; j0 += IIRY_G) {
    asm volatile("cp.async.wait_group %0;" ::"n"(IIRY_D / IIRY_G - 1) : "memory");
    float* slot = ring + (j0 & (IIRY_D - 1)) * IIRY_THREADS;
    float v[IIRY_G];
#pragma unroll
    for (int u = 0; u < IIRY_G; u++) v[u] = slot[u * IIRY_THREADS];  // stale beyond count: never used
    const int ns = start + DIR * j0, nl = start + DIR * (j0 + IIRY_D);
    if (j0 + IIRY_D + IIRY_G <= count && c.linear(ns, ns + DIR * (IIRY_G - 1)) && c.linear(nl, nl + DIR * (IIRY_G - 1))) {
      // steady state: the group and its refill are complete and evenly spaced - two pointers
      // and immediate multiples of the line's step, no bounds checks
      float* ps = c.ptr(ns);
      const float* pl = c.ptr(nl);
      const int ds = (int)c.delta(ns) * DIR, dl = (int)c.delta(nl) * DIR;  // 32-bit: one IMAD.WIDE per address
#pragma unroll
      for (int u = 0; u < IIRY_G; u++) v[u] = step(v[u]);
#pragma unroll
      for (int u = 0; u < IIRY_G; u++) ps[u * ds] = v[u];
#pragma unroll
      for (int u = 0; u < IIRY_G; u++) cp_async4(slot + u * IIRY_THREADS, pl + u * dl);
    } else {
#pragma unroll
      for (int u = 0; u < IIRY_G; u++)
        if (j0 + u < count) v[u] = step(v[u]);
#pragma unroll
      for (int u = 0; u < IIRY_G; u++)
        if (j0 + u < count) *c.ptr(start + DIR * (j0 + u)) = v[u];
#pragma unroll
      for (int u = 0; u < IIRY_G; u++)
        if (j0 + IIRY_D + u < count) cp_async4(slot + u * IIRY_THREADS, c.ptr(start + DIR * (j0 + IIRY_D + u)));
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}
template <typename Acc>
__device__ __forceinline__ void iir_line_ring(const IirDev& f, Acc& c, int M, float* ring) {
  if (M == 1 || f.npoles < 1) return;
  for (int k = 0; k < f.npoles; k++) {
    const float p = f.pole[k], g = f.gain;
    float X = iir_icc(f, c, M, k);
    if (k == 0) X = g * X;
    c(0) = X;
    if (k == 0)
      iir_sweep_ring<1>(c, 1, M - 1, [&](float v) { X = g * v + p * X; return X; }, ring);
    else
      iir_sweep_ring<1>(c, 1, M - 1, [&](float v) { X = v + p * X; return X; }, ring);
    X = iir_iacc(f, c, M, k);
    c(M - 1) = X;
    iir_sweep_ring<-1>(c, M - 2, M - 1, [&](float v) { X = p * (X - v); return X; }, ring);
  }
}

// lines along x: one thread per (row, channel). A warp covers 32 consecutive rows of one
// channel; each thread walks its row, so a 32-B sector fetched for texel n is reused for the
// next texels of the same row out of L1.
__global__ void k_iir_x(float* core, int stride, int nch, int w, int h, IirDev f) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= h * nch) return;
  int c = i / h, y = i % h;
  StrideAcc a{core + (ptrdiff_t)y * stride + c, nch};
  iir_line<8>(f, a, w);
}

// lines along x, tiled: a block owns IIR_R rows (all channels: IIR_R*NCH lines, one thread
// each). The row segments of a tile are moved between HBM and shared memory by all threads
// together (coalesced 4-byte accesses along the rows); the recursion runs on the tile in shared
// memory, carrying X from tile to tile: a forward sweep over the tiles for the causal filter, a
// backward sweep for the anticausal one. The initial coefficients read the line in HBM directly
// (they touch `horizon` elements). Tile pitch = IIR_TW*NCH + NCH floats, so that thread
// (row r, channel ch) = lane r*NCH+ch reads bank (lane + n*NCH) mod 32: conflict-free.
#define IIR_R 32
#define IIR_TW 64
template <int NCH>
__global__ void __launch_bounds__(IIR_R* NCH) k_iir_x_tiled(float* core, int stride, int w, int h, IirDev f) {
  constexpr int ROWF = IIR_TW * NCH;  // floats per tile row
  constexpr int PITCH = ROWF + NCH;
  constexpr int NT = IIR_R * NCH;
  __shared__ float tile[IIR_R * PITCH];
  const int tid = threadIdx.x;
  const int y0 = blockIdx.x * IIR_R;
  const int r = tid / NCH, ch = tid % NCH;
  const bool active = (y0 + r) < h;
  const int rows = min(IIR_R, h - y0);
  const int ntiles = (w + IIR_TW - 1) / IIR_TW;
  StrideAcc line{core + (ptrdiff_t)(y0 + (active ? r : 0)) * stride + ch, NCH};
  float* const trow = tile + r * PITCH + ch;
  float* const base = core + (ptrdiff_t)y0 * stride;
  auto load_tile = [&](int t) {
    const int x0f = t * ROWF, nf = min(ROWF, w * NCH - x0f);
    for (int i = tid; i < rows * ROWF; i += NT) {
      int rr = i / ROWF, kk = i - rr * ROWF;
      if (kk < nf) tile[rr * PITCH + kk] = base[(ptrdiff_t)rr * stride + x0f + kk];
    }
  };
  auto store_tile = [&](int t) {
    const int x0f = t * ROWF, nf = min(ROWF, w * NCH - x0f);
    for (int i = tid; i < rows * ROWF; i += NT) {
      int rr = i / ROWF, kk = i - rr * ROWF;
      if (kk < nf) base[(ptrdiff_t)rr * stride + x0f + kk] = tile[rr * PITCH + kk];
    }
  };
  if (w == 1 || f.npoles < 1) return;
  for (int k = 0; k < f.npoles; k++) {
    const float p = f.pole[k], g = f.gain;
    float X = 0.0f;
    if (active) {
      X = iir_icc(f, line, w, k);
      if (k == 0) X = g * X;
    }
    __syncthreads();  // every line has read its initial sum before the sweep rewrites the rows
    for (int t = 0; t < ntiles; t++) {
      load_tile(t);
      __syncthreads();
      if (active) {
        const int n0 = t * IIR_TW, cnt = min(IIR_TW, w - n0);
        for (int n = 0; n < cnt; n++) {
          if (n0 + n > 0) X = (k == 0) ? g * trow[n * NCH] + p * X : trow[n * NCH] + p * X;
          trow[n * NCH] = X;
        }
      }
      __syncthreads();
      store_tile(t);
      __syncthreads();
    }
    if (active) X = iir_iacc(f, line, w, k);  // reads what this block has just stored
    for (int t = ntiles - 1; t >= 0; t--) {
      load_tile(t);
      __syncthreads();
      if (active) {
        const int n0 = t * IIR_TW, cnt = min(IIR_TW, w - n0);
        for (int n = cnt - 1; n >= 0; n--) {
          if (n0 + n < w - 1) X = p * (X - trow[n * NCH]);
          trow[n * NCH] = X;
        }
      }
      __syncthreads();
      store_tile(t);
      __syncthreads();
    }
  }
}

// lines along x, row-resident: a block pulls its rows WHOLE into shared memory - one bulk
// asynchronous copy per row (cp.async.bulk -> mbarrier complete_tx, the TMA engine) - runs every
// pole's causal and anticausal recursion there, one thread per (row, channel) line, and writes
// the rows back with one bulk copy each. HBM sees every float once in each direction (the tiled
// kernel above moves it twice per pole), and no thread ever waits for HBM inside the recursion.
// What bounds it is the recursion's dependent multiply-add chain (2 x w steps per pole) times
// the lines that fit an SM's shared memory, so the launcher sizes blocks for two per SM: one
// computes while the other one's copies are in flight.
// Rows are copied as the 16-byte-aligned span around the core row (`lead` floats before it, the
// span rounded up to a granule): those few brace floats belong to the same container row and
// come back unchanged. Row pitch in shared memory = span + 4 floats, so that the lines of
// consecutive rows start four banks apart.
__device__ __forceinline__ uint32_t st_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int NCH>
struct SmemLineAcc {
  float* base;
  __device__ __forceinline__ float& operator()(int n) { return base[n * NCH]; }
};
template <int NCH>
__global__ void k_iir_x_rows(float* core, int stride, int w, int h, IirDev f, int rpb, int lead, int span) {
  extern __shared__ __align__(128) float srows[];
  __shared__ __align__(8) uint64_t mbar;
  const int tid = threadIdx.x;
  const int y0 = blockIdx.x * rpb;
  const int nrows = min(rpb, h - y0);
  const int spitch = span + 4;
  const uint32_t bar = st_smem_u32(&mbar);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)(nrows * span) * 4u)
                 : "memory");
  }
  __syncthreads();
  float* const g0 = core - lead + (ptrdiff_t)y0 * stride;
  for (int r = tid; r < nrows; r += blockDim.x)
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     st_smem_u32(srows + r * spitch)),
                 "l"(g0 + (ptrdiff_t)r * stride), "r"((uint32_t)span * 4u), "r"(bar)
                 : "memory");
  uint32_t landed;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(landed)
        : "r"(bar)
        : "memory");
  } while (!landed);
  if (tid < nrows * NCH) {
    SmemLineAcc<NCH> line{srows + (tid / NCH) * spitch + lead + (tid % NCH)};
    iir_line<8>(f, line, w);
  }
  // the rows were written through the generic proxy; the bulk store reads them through the async one
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  for (int r = tid; r < nrows; r += blockDim.x)
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(g0 + (ptrdiff_t)r * stride),
                 "r"(st_smem_u32(srows + r * spitch)), "r"((uint32_t)span * 4u)
                 : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // shared memory must outlive the reads
}

// lines along y: one thread per float of a row (column x channel): consecutive threads touch
// consecutive addresses at every step of the recursion -> fully coalesced. n_sections > 1:
// the container is a stack of sections of height h that are filtered separately (cubemap IR).
__global__ void __launch_bounds__(IIRY_THREADS) k_iir_y(float* core, int stride, int rowfloats, int h, int n_sections, IirDev f) {
  __shared__ float ring[IIRY_D * IIRY_THREADS];
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rowfloats * n_sections) return;
  int s = i / rowfloats, x = i % rowfloats;
  StrideAcc a{core + (ptrdiff_t)s * h * stride + x, stride};
  iir_line_ring(f, a, h, ring + threadIdx.x);
}

__global__ void __launch_bounds__(IIRY_THREADS) k_iir_y_spherical(float* core, int stride, int nch, int w, int h, IirDev f) {
  __shared__ float ring[IIRY_D * IIRY_THREADS];
  int half = w / 2;
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= half * nch) return;
  PoleAcc a{core + i, core + (ptrdiff_t)(h - 1) * stride + (ptrdiff_t)half * nch + i, stride, h};
  iir_line_ring(f, a, 2 * h, ring + threadIdx.x);
}

// brace: every container texel outside the core is a copy of a core texel (PERIODIC / REFLECT
// index maps of zimt/brace.h:189-215; spherical: rows beyond the poles continue on the
// opposite meridian, environment.h:473-516, then the periodic x brace over all rows)
// zimt's bracer fills the two braces in lock-step from the core outwards, each slice from the slice its mirror
// image (or period) points at - a brace slice filled a few steps earlier when the brace is wider than the core
// (a 2-px raster under a quintic spline). The closed form of that is folding the index with period 2n
// (REFLECT) or n (PERIODIC).
__device__ __forceinline__ int brace_map(int i, int n, int bc) {
  if (i >= 0 && i < n) return i;
  if (bc == EU_BC_PERIODIC) {
    i %= n;
    return i < 0 ? i + n : i;
  }
  i %= 2 * n;
  if (i < 0) i += 2 * n;
  return i < n ? i : 2 * n - 1 - i;
}
// The launch covers the frame only: (ly + ry) full container rows, then (lx + rx) columns beside the
// core rows - a few thousand texels, not the whole container.
__global__ void k_brace(float* core, int stride, int nch, int w, int h, int lx, int rx, int ly, int ry, int bc0,
                        int bc1, int spherical) {
  const int cw = w + lx + rx;
  const long long n_rows = (long long)cw * (ly + ry), n_cols = (long long)(lx + rx) * h;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_rows + n_cols) return;
  int X, Y;  // container coordinates
  if (i < n_rows) {
    int r = (int)(i / cw);
    X = (int)(i - (long long)r * cw);
    Y = r < ly ? r : h + r;  // rows 0..ly-1 above the core, rows ly+h.. below it
  } else {
    long long j = i - n_rows;
    int r = (int)(j / (lx + rx)), c = (int)(j - (long long)r * (lx + rx));
    Y = ly + r;
    X = c < lx ? c : w + c;  // columns 0..lx-1 left of the core, columns lx+w.. right of it
  }
  int x = X - lx, y = Y - ly;
  int sx = brace_map(x, w, bc0), sy;
  if (spherical && (y < 0 || y >= h)) {
    sy = y < 0 ? -1 - y : 2 * h - 1 - y;
    int half = w / 2;
    sx = sx < half ? sx + half : sx - half;
  } else {
    sy = brace_map(y, h, bc1);
  }
  const float* s = core + (ptrdiff_t)sy * stride + (ptrdiff_t)sx * nch;
  float* d = core + (ptrdiff_t)y * stride + (ptrdiff_t)x * nch;
  for (int c = 0; c < nch; c++) d[c] = s[c];
}

// ---- cubemap IR support (cubemap.h:607-911) ------------------------------------------------
// 1-px mirrored ring around every cube face (mirror_around, :607-660). Corners are written by
// the column pass from the row pass' result in the reference; the value is the face corner.
__global__ void k_cm_ring(float* ir, int stride, int nch, int F, int S, int L, int R) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;  // position along the ring side, -1..F
  int face = blockIdx.y, side = blockIdx.z;
  int t = i - 1;
  if (t > F) return;
  int cmin = L > 0 ? -1 : 0, cmax = R > 0 ? F : F - 1;
  if (t < cmin || t > cmax) return;
  float* f0 = ir + (ptrdiff_t)(face * S + L) * stride + (ptrdiff_t)L * nch;  // face texel (0,0)
  auto px = [&](int x, int y) { return f0 + (ptrdiff_t)y * stride + (ptrdiff_t)x * nch; };
  int tc = t < 0 ? 0 : (t > F - 1 ? F - 1 : t);  // the row pass has filled (t,-1)/(t,F) from (t,0)/(t,F-1)
  const float* s;
  float* d;
  switch (side) {
    case 0: if (!L) return; d = px(t, -1); s = px(tc, 0); break;
    case 1: if (!R) return; d = px(t, F); s = px(tc, F - 1); break;
    case 2: if (!L) return; d = px(-1, t); s = px(0, tc); break;
    default: if (!R) return; d = px(F, t); s = px(F - 1, tc); break;
  }
  for (int c = 0; c < nch; c++) d[c] = s[c];
}

// one frame stripe of one section, by bilinear reprojection from the other sections
// (fill_frame_t::eval, cubemap.h:733-810). Coordinates are doubled integers relative to the
// section centre (:867-868).
// fill_frame_t::eval for frame pixel (x, y) of section `face` (cubemap.h:733-810)
template <int NCH, int SPACE>
__device__ __forceinline__ void dev_cm_fill_pixel(const SourceDev& S, int face, int x, int y, int section_px, int ithird,
                                                  double refc_md, float model_to_px, float px[NCH]) {
  int ishift = section_px - 1;
  int c0 = 2 * x - ishift, c1 = 2 * y - ishift;
  float ray[3];
  switch (face) {
    case CM_FRONT: ray[0] = (float)c0; ray[1] = (float)c1; ray[2] = (float)ithird; break;
    case CM_BACK: ray[0] = (float)(-c0); ray[1] = (float)c1; ray[2] = (float)(-ithird); break;
    case CM_RIGHT: ray[0] = (float)ithird; ray[1] = (float)c1; ray[2] = (float)(-c0); break;
    case CM_LEFT: ray[0] = (float)(-ithird); ray[1] = (float)c1; ray[2] = (float)c0; break;
    case CM_BOTTOM: ray[0] = (float)(-c0); ray[1] = (float)ithird; ray[2] = (float)c1; break;
    default: ray[0] = (float)(-c0); ray[1] = (float)(-ithird); ray[2] = (float)(-c1); break;
  }
  int fv;
  float in_face[2], pk[2];
  dev_cubeface(ray, fv, in_face);
  // metrics_t::get_pickup_coordinate_px, cubemap.h:401-411 (refc_md is a double member there)
  pk[0] = (float)((double)in_face[0] + refc_md);
  pk[1] = (float)((double)in_face[1] + refc_md);
  pk[0] *= model_to_px;
  pk[1] *= model_to_px;
  pk[1] += (float)(fv * section_px);
  pk[0] -= .5f;
  pk[1] -= .5f;
  dev_spline_eval<NCH, NCH, 1, SPACE>(S, 1, nullptr, pk[0], pk[1], px);
}

template <int NCH>
__global__ void k_cm_fill(float* ir, SourceDev S, int face, int L, int R, int section_px, int ithird, double refc_md,
                          float model_to_px) {
  // the four stripes of the section's frame as one index space: top (S x L), bottom (S x R), left (L x mid),
  // right (R x mid), mid = S - L - R rows. With an EVEN face width a frame pixel's ray hits ANOTHER face, so the
  // stripes of one section do not read each other and one launch per face keeps the reference's result.
  const int Sx = section_px, mid = Sx - L - R;
  const long long n_top = (long long)Sx * L, n_bot = (long long)Sx * R, n_left = (long long)L * mid, n_right = (long long)R * mid;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  int x, y;
  if (i < n_top) { y = (int)(i / Sx); x = (int)(i % Sx); }
  else if ((i -= n_top) < n_bot) { y = Sx - R + (int)(i / Sx); x = (int)(i % Sx); }
  else if ((i -= n_bot) < n_left) { y = L + (int)(i / L); x = (int)(i % L); }
  else if ((i -= n_left) < n_right) { y = L + (int)(i / R); x = Sx - R + (int)(i % R); }
  else return;
  float px[NCH];
  dev_cm_fill_pixel<NCH, 0>(S, face, x, y, section_px, ithird, refc_md, model_to_px, px);
  float* d = ir + (ptrdiff_t)(face * section_px + y) * S.stride + (ptrdiff_t)x * NCH;
#pragma unroll
  for (int c = 0; c < NCH; c++) d[c] = px[c];
}

// ODD face widths. The face is then not centred in its section (left frame = right frame - 1) while the
// pixel-to-ray step assumes it is (ishift = section_px - 1, cubemap.h:861-868), so some frame pixels map onto
// their OWN section and read frame texels the same fill is rewriting: the result depends on the order in which
// zimt::process works - faces, stripes and lines in sequence, a line in vectors of 16 pixels, each vector
// evaluated from the raster as it is and then stored (zimt/wielding.h:317-455). One block walks exactly that
// order (lanes 0..15 = the pixels of a vector; loads bypass the non-coherent caches). Where the reference is
// deterministic this reproduces it; the LEFT/RIGHT column that reads the line above races in the reference
// itself (another thread works on that line) - there the lines are taken in order, as the oracle does.
// Slow by construction (a few microseconds per vector), used for odd widths only.
template <int NCH>
__global__ void k_cm_fill_ordered(float* ir, SourceDev S, int L, int R, int F, int section_px, int ithird, double refc_md,
                                  float model_to_px) {
  const int Sx = section_px, lane = threadIdx.x;
  for (int face = 0; face < 6; face++) {
    const int win[4][4] = {{0, 0, Sx, L}, {0, Sx - R, Sx, Sx}, {0, L, L, Sx - R}, {L + F, L, Sx, Sx - R}};
    const int on[4] = {L > 0, R > 0, L > 0, R > 0};
    for (int st = 0; st < 4; st++) {
      if (!on[st]) continue;
      const int x0 = win[st][0], y0 = win[st][1], x1 = win[st][2], y1 = win[st][3];
      for (int y = y0; y < y1; y++) {
        for (int v0 = x0; v0 < x1; v0 += EU_LANES) {
          const int x = v0 + lane;
          float px[NCH];
          const bool mine = lane < EU_LANES && x < x1;
          if (mine) dev_cm_fill_pixel<NCH, 2>(S, face, x, y, section_px, ithird, refc_md, model_to_px, px);
          __syncthreads();  // every pixel of the vector is evaluated before any is stored
          if (mine) {
            float* d = ir + (ptrdiff_t)(face * section_px + y) * S.stride + (ptrdiff_t)x * NCH;
#pragma unroll
            for (int c = 0; c < NCH; c++) __stcg(d + c, px[c]);
          }
          __threadfence_block();
          __syncthreads();  // ... and stored before the next vector is evaluated
        }
      }
    }
  }
}

// cubemap_t::load for a raster that is already in device memory (cubemap.h:607-640: the six faces go to the centres
// of their sections): ONE launch for all six faces instead of six pitched device-to-device copies with a stream
// drain between them. V = float4 when rows and offsets are 16-byte multiples, else float.
template <typename V>
__global__ void __launch_bounds__(256) k_cm_place(const V* __restrict__ src, V* __restrict__ dst, int row_v, int face_px,
                                                  int section_px, int left_px, int dst_pitch_v, int left_v) {
  const size_t n = (size_t)6 * face_px * row_v;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const size_t r = i / row_v;
    const int k = (int)(i - r * row_v);
    const int face = (int)(r / face_px), y = (int)(r - (size_t)face * face_px);
    dst[((size_t)face * section_px + left_px + y) * dst_pitch_v + left_v + k] = __ldg(src + i);
  }
}

// interleaved nch-float texels (rows of src_pitch floats) -> dense 16-byte texels
__global__ void k_pad_texels(const float* __restrict__ src, int src_pitch, float4* __restrict__ dst, int dst_pitch_texels,
                             int cw, int chh, int nch) {
  int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= cw) return;
  for (int y = blockIdx.y; y < chh; y += gridDim.y) {  // gridDim.y is limited to 65535 rows
    const float* s = src + (size_t)y * src_pitch + (size_t)x * nch;
    dst[(size_t)y * dst_pitch_texels + x] = make_float4(s[0], nch > 1 ? s[1] : 0.f, nch > 2 ? s[2] : 0.f, 0.f);
  }
}

// ---- alpha of masked / cropped facets (environment.h:703-890) -----------------------------
// 5-tap binomial (1 4 6 4 1)/16, REFLECT extrapolation (zimt/extrapolate.h:141-153). On 0/1 data
// the x pass yields multiples of 1/16, the y pass multiples of 1/256: all partial sums are exact,
// so the summation order of the reference's circular-buffer FIR (zimt/convolve.h) is immaterial.
__device__ __forceinline__ int dev_reflect(int i, int w) {
  if (i < 0) i = -1 - i;
  if (i >= w) {
    i %= 2 * w;
    if (i >= w) i = 2 * w - i - 1;
  }
  return i;
}
__global__ void k_alpha_feather_x(const unsigned char* __restrict__ in, float* __restrict__ out, int w, int h) {
  int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= w) return;
  const float k5[5] = {1.0f / 16.0f, 4.0f / 16.0f, 6.0f / 16.0f, 4.0f / 16.0f, 1.0f / 16.0f};
  for (int y = blockIdx.y; y < h; y += gridDim.y) {
    float s = 0.0f;
#pragma unroll
    for (int j = 0; j < 5; j++) s += k5[j] * (float)in[(size_t)y * w + dev_reflect(x - 2 + j, w)];
    out[(size_t)y * w + x] = s;
  }
}
__global__ void k_alpha_feather_y(const float* __restrict__ in, float* __restrict__ out, int w, int h) {
  int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= w) return;
  const float k5[5] = {1.0f / 16.0f, 4.0f / 16.0f, 6.0f / 16.0f, 4.0f / 16.0f, 1.0f / 16.0f};
  for (int y = blockIdx.y; y < h; y += gridDim.y) {
    float s = 0.0f;
#pragma unroll
    for (int j = 0; j < 5; j++) s += k5[j] * in[(size_t)dev_reflect(y - 2 + j, h) * w + x];
    out[(size_t)y * w + x] = s;
  }
}
// raster of native_nch channels -> nch channels (an added alpha channel is 1), times alpha
__global__ void k_alpha_apply(const float* __restrict__ raw, int native_nch, const float* __restrict__ alpha,
                              float* __restrict__ out, int nch, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float a = alpha[i];
  for (int c = 0; c < nch; c++) {
    float v = c < native_nch ? raw[i * native_nch + c] : 1.0f;
    out[i * nch + c] = v * a;
  }
}

// ---- launchers -----------------------------------------------------------------------------
template <int NCH>
static bool launch_iir_x_rows(float* core, int stride, int w, int h, const IirDev& f, cudaStream_t st) {
  if ((stride & 3) || ((uintptr_t)core & 3)) return false;
  const int lead = (int)(((uintptr_t)core & 15) / 4);
  const int span = (lead + w * NCH + 3) & ~3;
  const size_t row_bytes = (size_t)(span + 4) * sizeof(float);
  size_t budget = (228 * 1024 - 2 * 1024) / 2 - 64;  // two blocks per SM
  if (row_bytes > budget) budget = 226 * 1024;        // very wide rows: one block per SM, one row each
  int rpb = (int)(budget / row_bytes);
  if (rpb < 1) return false;  // a row does not fit an SM's shared memory: the tiled kernel streams it
  rpb = std::min(rpb, 256 / NCH);
  rpb = std::min(rpb, std::max(1, (h + 295) / 296));  // small rasters: spread the rows over the SMs
  const int threads = ((rpb * NCH + 31) / 32) * 32;
  const size_t smem = (size_t)rpb * row_bytes;
  // per launch, not once: the attribute belongs to the device the library is initialised on, and that
  // may change between eu_shutdown and the next eu_init
  if (cudaFuncSetAttribute(k_iir_x_rows<NCH>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024) != cudaSuccess) {
    cudaGetLastError();  // not sticky: the tiled kernel takes over
    return false;
  }
  k_iir_x_rows<NCH><<<(h + rpb - 1) / rpb, threads, smem, st>>>(core, stride, w, h, f, rpb, lead, span);
  return true;
}

cudaError_t eu_launch_iir_x(float* core, int stride, int nch, int w, int h, const IirDev& f, cudaStream_t st) {
  if (w == 1 || f.npoles < 1) return cudaSuccess;
  bool done = false;
  switch (nch) {
    case 1: done = launch_iir_x_rows<1>(core, stride, w, h, f, st); break;
    case 2: done = launch_iir_x_rows<2>(core, stride, w, h, f, st); break;
    case 3: done = launch_iir_x_rows<3>(core, stride, w, h, f, st); break;
    case 4: done = launch_iir_x_rows<4>(core, stride, w, h, f, st); break;
  }
  if (done) return cudaGetLastError();
  int nb = (h + IIR_R - 1) / IIR_R;
  switch (nch) {
    case 1: k_iir_x_tiled<1><<<nb, IIR_R * 1, 0, st>>>(core, stride, w, h, f); break;
    case 2: k_iir_x_tiled<2><<<nb, IIR_R * 2, 0, st>>>(core, stride, w, h, f); break;
    case 3: k_iir_x_tiled<3><<<nb, IIR_R * 3, 0, st>>>(core, stride, w, h, f); break;
    case 4: k_iir_x_tiled<4><<<nb, IIR_R * 4, 0, st>>>(core, stride, w, h, f); break;
    default: {
      int n = h * nch;
      k_iir_x<<<(n + 63) / 64, 64, 0, st>>>(core, stride, nch, w, h, f);
    }
  }
  return cudaGetLastError();
}
cudaError_t eu_launch_iir_y(float* core, int stride, int nch, int w, int h, int n_sections, const IirDev& f,
                            cudaStream_t st) {
  int n = w * nch * n_sections;
  k_iir_y<<<(n + IIRY_THREADS - 1) / IIRY_THREADS, IIRY_THREADS, 0, st>>>(core, stride, w * nch, h, n_sections, f);
  return cudaGetLastError();
}
cudaError_t eu_launch_iir_y_spherical(float* core, int stride, int nch, int w, int h, const IirDev& f,
                                      cudaStream_t st) {
  int n = (w / 2) * nch;
  k_iir_y_spherical<<<(n + IIRY_THREADS - 1) / IIRY_THREADS, IIRY_THREADS, 0, st>>>(core, stride, nch, w, h, f);
  return cudaGetLastError();
}
__global__ void k_brace_natural_1d(float* core, int n, int k) {
  int i = threadIdx.x;
  if (i < k) {
    core[-1 - i] = core[0] + core[0] - core[1 + i];
    core[n + i] = core[n - 1] + core[n - 1] - core[n - 2 - i];
  }
}
cudaError_t eu_launch_brace_natural_1d(float* core, int n, int k, cudaStream_t st) {
  k_brace_natural_1d<<<1, 32, 0, st>>>(core, n, k);
  return cudaGetLastError();
}

cudaError_t eu_launch_brace(float* core, int stride, int nch, int w, int h, int lx, int rx, int ly, int ry, int bc0,
                            int bc1, int spherical, cudaStream_t st) {
  const long long n = (long long)(w + lx + rx) * (ly + ry) + (long long)(lx + rx) * h;
  if (n <= 0) return cudaSuccess;
  k_brace<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(core, stride, nch, w, h, lx, rx, ly, ry, bc0, bc1, spherical);
  return cudaGetLastError();
}

cudaError_t eu_launch_cubemap_support(float* ir, int pitch, int nch, int F, int S, int L, int R, double refc_md,
                                      double model_to_px, int* n_launches, cudaStream_t st) {
  *n_launches = 0;
  if (L == 0 && R == 0) return cudaSuccess;
  dim3 rgrid((F + 2 + 127) / 128, 6, 4);
  k_cm_ring<<<rgrid, 128, 0, st>>>(ir, pitch, nch, F, S, L, R);
  ++*n_launches;
  SourceDev src;
  src.core = ir;
  src.stride = pitch;
  src.tstride = nch;
  src.nch = nch;
  src.w = S;
  src.h = 6 * S;
  src.bc0 = src.bc1 = EU_BC_REFLECT;
  src.upper_x = (float)((long double)(S - 1) + 0.5L);
  src.upper_y = (float)((long double)(6 * S - 1) + 0.5L);
  int ithird = (int)(model_to_px * 2);
  if (F & 1) {  // odd face width: the fill reads what it writes - one block in the reference's order
    switch (nch) {
      case 1: k_cm_fill_ordered<1><<<1, 32, 0, st>>>(ir, src, L, R, F, S, ithird, refc_md, (float)model_to_px); break;
      case 2: k_cm_fill_ordered<2><<<1, 32, 0, st>>>(ir, src, L, R, F, S, ithird, refc_md, (float)model_to_px); break;
      case 3: k_cm_fill_ordered<3><<<1, 32, 0, st>>>(ir, src, L, R, F, S, ithird, refc_md, (float)model_to_px); break;
      case 4: k_cm_fill_ordered<4><<<1, 32, 0, st>>>(ir, src, L, R, F, S, ithird, refc_md, (float)model_to_px); break;
      default: return cudaErrorInvalidValue;
    }
    ++*n_launches;
    return cudaGetLastError();
  }
  // the reference fills face after face, and later faces read ring pixels that earlier ones have overwritten
  // (cubemap.h:819-911): keep that order. Within a face the four stripes are independent (see k_cm_fill).
  const int mid = S - L - R;
  const long long n = (long long)S * (L + R) + (long long)(L + R) * (mid > 0 ? mid : 0);
  const unsigned blocks = (unsigned)((n + 255) / 256);
  for (int face = 0; face < 6; face++) {
    switch (nch) {
      case 1: k_cm_fill<1><<<blocks, 256, 0, st>>>(ir, src, face, L, R, S, ithird, refc_md, (float)model_to_px); break;
      case 2: k_cm_fill<2><<<blocks, 256, 0, st>>>(ir, src, face, L, R, S, ithird, refc_md, (float)model_to_px); break;
      case 3: k_cm_fill<3><<<blocks, 256, 0, st>>>(ir, src, face, L, R, S, ithird, refc_md, (float)model_to_px); break;
      case 4: k_cm_fill<4><<<blocks, 256, 0, st>>>(ir, src, face, L, R, S, ithird, refc_md, (float)model_to_px); break;
      default: return cudaErrorInvalidValue;
    }
    ++*n_launches;
  }
  return cudaGetLastError();
}

cudaError_t eu_launch_pad_texels(const float* src, int src_pitch, float* dst, int dst_pitch_texels, int cw, int chh, int nch,
                                 cudaStream_t st) {
  dim3 grid((cw + 255) / 256, chh < 65535 ? chh : 65535);
  k_pad_texels<<<grid, 256, 0, st>>>(src, src_pitch, reinterpret_cast<float4*>(dst), dst_pitch_texels, cw, chh, nch);
  return cudaGetLastError();
}

cudaError_t eu_launch_cubemap_place(const float* src, float* ir, int pitch, int nch, int F, int S, int L, cudaStream_t st) {
  const int rowf = F * nch, leftf = L * nch;
  const bool vec = !(rowf & 3) && !(leftf & 3) && !(pitch & 3) && !((uintptr_t)src & 15) && !((uintptr_t)ir & 15);
  const size_t n = (size_t)6 * F * (vec ? rowf / 4 : rowf);
  const size_t want = (n + 255) / 256;
  const int blocks = (int)(want < (size_t)148 * 32 ? want : (size_t)148 * 32);
  if (vec)
    k_cm_place<float4><<<blocks, 256, 0, st>>>(reinterpret_cast<const float4*>(src), reinterpret_cast<float4*>(ir), rowf / 4, F, S, L,
                                               pitch / 4, leftf / 4);
  else
    k_cm_place<float><<<blocks, 256, 0, st>>>(src, ir, rowf, F, S, L, pitch, leftf);
  return cudaGetLastError();
}

cudaError_t eu_launch_alpha_apply(const unsigned char* mask, float* tmp_a, float* tmp_b, const float* raw, int native_nch,
                                  float* out, int nch, int w, int h, cudaStream_t st) {
  dim3 grid((w + 255) / 256, h < 65535 ? h : 65535);
  k_alpha_feather_x<<<grid, 256, 0, st>>>(mask, tmp_a, w, h);
  k_alpha_feather_y<<<grid, 256, 0, st>>>(tmp_a, tmp_b, w, h);
  size_t n = (size_t)w * h;
  k_alpha_apply<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(raw, native_nch, tmp_b, out, nch, n);
  return cudaGetLastError();
}
