// kernels_vec.cu -- Krylov vector kernels (replaces PETSc VecAXPY/VecAYPX/VecWAXPY/VecMDot/VecMAXPY/
// VecNorm/VecPointwiseMult reached from KSPSolve, reference call site src/SaddlePointProblem.c:70).
//
// All kernels are HBM-bound streams: 128-bit loads/stores, grids sized as a multiple of the SM count,
// reductions by warp-shuffle trees with a deterministic fixed-order second stage (dev.cuh).
// Algorithmic bytes: copy 16n, scale 16n, AXPY 24n, WAXPY 24n, dot 16n, norm 8n,
// MDot(k) 8(k+1)n, MAXPY(k) 8(k+2)n, fused MAXPY+norm 8(k+2)n.
#include "dev.cuh"

namespace b200sp {

namespace {

enum MapOp { OP_SET, OP_SCALE, OP_AXPY, OP_AYPX, OP_WAXPY, OP_AXPBYPCZ, OP_CHEB, OP_CHEB_D, OP_PMULT, OP_RECIP, OP_COPY };

struct MapArgs {
  double a, b, c;
  const double *x, *y, *z, *d;
  double *w;
};

template <int OP>
__device__ __forceinline__ double map_one(const MapArgs &m, int64_t i) {
  switch (OP) {
  case OP_SET: return m.a;
  case OP_COPY: return m.x[i];
  case OP_SCALE: return m.a * m.w[i];
  case OP_AXPY: return m.w[i] + m.a * m.x[i];
  case OP_AYPX: return m.x[i] + m.a * m.w[i];
  case OP_WAXPY: return m.a * m.x[i] + m.y[i];
  case OP_AXPBYPCZ: return m.a * m.x[i] + m.b * m.y[i] + m.c * m.z[i];
  case OP_CHEB: return m.a * m.x[i] + m.b * m.y[i] + m.c * m.z[i];
  case OP_CHEB_D: return m.a * m.x[i] + m.b * m.y[i] + m.c * (m.z[i] * m.d[i]);
  case OP_PMULT: return m.x[i] * m.y[i];
  case OP_RECIP: { double v = m.w[i]; return 1.0 / (v == 0.0 ? 1.0 : v); }
  }
  return 0.0;
}

// two elements per thread per step, 128-bit accesses when every pointer is 16-byte aligned
template <int OP, bool V2>
__global__ void __launch_bounds__(256) k_map(int64_t n, MapArgs m) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  if (V2) {
    const int64_t n2 = n >> 1;
    for (int64_t i2 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i2 < n2; i2 += stride) {
      double2 r;
      if (OP == OP_SET) { r.x = m.a; r.y = m.a; }
      else {
        // load as vectors, combine with the same scalar expression as map_one
        double2 x = {0, 0}, y = {0, 0}, z = {0, 0}, d = {0, 0}, w = {0, 0};
        if (OP == OP_COPY || OP == OP_AXPY || OP == OP_AYPX || OP == OP_WAXPY || OP == OP_AXPBYPCZ || OP == OP_CHEB || OP == OP_CHEB_D || OP == OP_PMULT)
          x = *reinterpret_cast<const double2 *>(m.x + 2 * i2);
        if (OP == OP_WAXPY || OP == OP_AXPBYPCZ || OP == OP_CHEB || OP == OP_CHEB_D || OP == OP_PMULT)
          y = *reinterpret_cast<const double2 *>(m.y + 2 * i2);
        if (OP == OP_AXPBYPCZ || OP == OP_CHEB || OP == OP_CHEB_D) z = *reinterpret_cast<const double2 *>(m.z + 2 * i2);
        if (OP == OP_CHEB_D) d = *reinterpret_cast<const double2 *>(m.d + 2 * i2);
        if (OP == OP_SCALE || OP == OP_AXPY || OP == OP_AYPX || OP == OP_RECIP) w = *reinterpret_cast<const double2 *>(m.w + 2 * i2);
        switch (OP) {
        case OP_COPY: r = x; break;
        case OP_SCALE: r.x = m.a * w.x; r.y = m.a * w.y; break;
        case OP_AXPY: r.x = w.x + m.a * x.x; r.y = w.y + m.a * x.y; break;
        case OP_AYPX: r.x = x.x + m.a * w.x; r.y = x.y + m.a * w.y; break;
        case OP_WAXPY: r.x = m.a * x.x + y.x; r.y = m.a * x.y + y.y; break;
        case OP_AXPBYPCZ:
        case OP_CHEB: r.x = m.a * x.x + m.b * y.x + m.c * z.x; r.y = m.a * x.y + m.b * y.y + m.c * z.y; break;
        case OP_CHEB_D: r.x = m.a * x.x + m.b * y.x + m.c * (z.x * d.x); r.y = m.a * x.y + m.b * y.y + m.c * (z.y * d.y); break;
        case OP_PMULT: r.x = x.x * y.x; r.y = x.y * y.y; break;
        case OP_RECIP: r.x = 1.0 / (w.x == 0.0 ? 1.0 : w.x); r.y = 1.0 / (w.y == 0.0 ? 1.0 : w.y); break;
        default: r.x = r.y = 0.0;
        }
      }
      *reinterpret_cast<double2 *>(m.w + 2 * i2) = r;
    }
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) m.w[n - 1] = map_one<OP>(m, n - 1);
  } else {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) m.w[i] = map_one<OP>(m, i);
  }
}

inline int stream_grid(Ctx *c, int64_t work_items) {
  int64_t want = (work_items + 255) / 256;
  int64_t cap = (int64_t)c->num_sms * 8; // 8 CTAs of 256 threads per SM = full occupancy
  if (want < 1) want = 1;
  return (int)(want < cap ? want : cap);
}

template <int OP>
void launch_map(Ctx *c, int64_t n, const MapArgs &m) {
  if (n <= 0) return;
  bool v2 = aligned16(m.w) && (!m.x || aligned16(m.x)) && (!m.y || aligned16(m.y)) && (!m.z || aligned16(m.z)) && (!m.d || aligned16(m.d));
  LaunchScope ls(c, "vec");
  if (v2) k_map<OP, true><<<stream_grid(c, (n + 1) / 2), 256, 0, c->stream>>>(n, m);
  else k_map<OP, false><<<stream_grid(c, n), 256, 0, c->stream>>>(n, m);
  check_launch("k_map");
}

__global__ void __launch_bounds__(256) k_hash(int64_t n, double *v) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    unsigned h = (unsigned)i * 2654435761u;
    h ^= h >> 16; h *= 0x85ebca6bu; h ^= h >> 13;
    v[i] = 0.5 + (double)(h >> 8) / 16777216.0;
  }
}
// the same hashed vector on a row-partitioned DMDA vector: entry (node (i,j), component c) gets the value its NATURAL
// global index ((j*M + i)*dof + c) has on one rank, so eigenvalue estimates do not depend on the partition
__global__ void __launch_bounds__(256) k_hash_natural(int xs, int ys, int xm, int ym, int M, int dof, double *v) {
  const int64_t n = (int64_t)xm * ym * dof;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
    const int node = (int)(t / dof), c = (int)(t % dof);
    const int i = xs + node % xm, j = ys + node / xm;
    unsigned h = (unsigned)(((int64_t)j * M + i) * dof + c) * 2654435761u;
    h ^= h >> 16; h *= 0x85ebca6bu; h ^= h >> 13;
    v[t] = 0.5 + (double)(h >> 8) / 16777216.0;
  }
}
__global__ void __launch_bounds__(256) k_scatter_set(int64_t n, const int *idx, double val, double *y) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) y[idx[i]] = val;
}

__global__ void __launch_bounds__(256) k_permute_scatter(int64_t n, const int *__restrict__ map, const double *__restrict__ in, double *out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    if (map[i] >= 0) out[map[i]] = in[i];
}
__global__ void __launch_bounds__(256) k_permute_gather(int64_t n, const int *__restrict__ map, const double *__restrict__ in, double *out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) out[i] = in[map[i]];
}

// ---- reductions -------------------------------------------------------------------------------
// out[j] = sum_i w[i] * V[j*ld + i], j < k <= KB.  One pass over w and the k basis vectors.
template <int KB, bool V2>
__global__ void __launch_bounds__(256) k_mdot(int64_t n, int k, const double *__restrict__ w, const double *__restrict__ V, int64_t ld,
                                              double *partials, unsigned *ticket, double *out) {
  double acc[KB];
#pragma unroll
  for (int j = 0; j < KB; ++j) acc[j] = 0.0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  if (V2) {
    const int64_t n2 = n >> 1;
    for (int64_t i2 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i2 < n2; i2 += stride) {
      const double2 wv = ld_stream_f64x2(w + 2 * i2);
#pragma unroll
      for (int j = 0; j < KB; ++j)
        if (j < k) {
          const double2 v = ld_stream_f64x2(V + (size_t)j * ld + 2 * i2);
          acc[j] = fma(wv.x, v.x, acc[j]);
          acc[j] = fma(wv.y, v.y, acc[j]);
        }
    }
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
#pragma unroll
      for (int j = 0; j < KB; ++j)
        if (j < k) acc[j] = fma(w[n - 1], V[(size_t)j * ld + n - 1], acc[j]);
    }
  } else {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
      const double wv = w[i];
#pragma unroll
      for (int j = 0; j < KB; ++j)
        if (j < k) acc[j] = fma(wv, V[(size_t)j * ld + i], acc[j]);
    }
  }
  grid_reduce_sum<KB>(acc, k, partials, ticket, out);
}

// w -= sum_j h[j] V_j  (j ascending, one fma each), out[0] = sum w_new^2.  h lives on the device.
template <int KB, bool V2, bool NORM>
__global__ void __launch_bounds__(256) k_maxpy(int64_t n, int k, double sign, double *__restrict__ w, const double *__restrict__ V, int64_t ld,
                                               const double *__restrict__ h, double *partials, unsigned *ticket, double *out) {
  __shared__ double s_h[KB];
  if (threadIdx.x < KB) s_h[threadIdx.x] = threadIdx.x < k ? sign * h[threadIdx.x] : 0.0;
  __syncthreads();
  double acc[1] = {0.0};
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  if (V2) {
    const int64_t n2 = n >> 1;
    for (int64_t i2 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i2 < n2; i2 += stride) {
      double2 wv = *reinterpret_cast<const double2 *>(w + 2 * i2);
#pragma unroll
      for (int j = 0; j < KB; ++j)
        if (j < k) {
          const double2 v = ld_stream_f64x2(V + (size_t)j * ld + 2 * i2);
          wv.x = fma(s_h[j], v.x, wv.x);
          wv.y = fma(s_h[j], v.y, wv.y);
        }
      *reinterpret_cast<double2 *>(w + 2 * i2) = wv;
      if (NORM) { acc[0] = fma(wv.x, wv.x, acc[0]); acc[0] = fma(wv.y, wv.y, acc[0]); }
    }
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
      double wv = w[n - 1];
#pragma unroll
      for (int j = 0; j < KB; ++j)
        if (j < k) wv = fma(s_h[j], V[(size_t)j * ld + n - 1], wv);
      w[n - 1] = wv;
      if (NORM) acc[0] = fma(wv, wv, acc[0]);
    }
  } else {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
      double wv = w[i];
#pragma unroll
      for (int j = 0; j < KB; ++j)
        if (j < k) wv = fma(s_h[j], V[(size_t)j * ld + i], wv);
      w[i] = wv;
      if (NORM) acc[0] = fma(wv, wv, acc[0]);
    }
  }
  if (NORM) grid_reduce_sum<1>(acc, 1, partials, ticket, out);
}

template <bool V2>
__global__ void __launch_bounds__(256) k_scale_inv_sqrt(int64_t n, const double *nrm2, const double *x, double *y) { // x may alias y
  const double s = 1.0 / sqrt(*nrm2);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  if (V2) {
    const int64_t n2 = n >> 1;
    for (int64_t i2 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i2 < n2; i2 += stride) {
      double2 v = *reinterpret_cast<const double2 *>(x + 2 * i2);
      v.x *= s; v.y *= s;
      *reinterpret_cast<double2 *>(y + 2 * i2) = v;
    }
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) y[n - 1] = x[n - 1] * s;
  } else {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) y[i] = x[i] * s;
  }
}

inline int reduce_grid(Ctx *c, int64_t work_items, int ctas_per_sm) {
  int64_t want = (work_items + 255) / 256;
  int64_t cap = (int64_t)c->num_sms * ctas_per_sm;
  if (cap > RED_MAX_BLOCKS) cap = RED_MAX_BLOCKS;
  if (want < 1) want = 1;
  return (int)(want < cap ? want : cap);
}

template <int KB>
void launch_mdot(Ctx *c, int64_t n, int k, const double *w, const double *V, int64_t ld, double *out) {
  bool v2 = aligned16(w) && aligned16(V) && (ld % 2 == 0);
  LaunchScope ls(c, "mdot");
  int grid = reduce_grid(c, v2 ? (n + 1) / 2 : n, KB >= 16 ? 4 : 6);
  if (v2) k_mdot<KB, true><<<grid, 256, 0, c->stream>>>(n, k, w, V, ld, c->d_partials, c->d_ticket, out);
  else k_mdot<KB, false><<<grid, 256, 0, c->stream>>>(n, k, w, V, ld, c->d_partials, c->d_ticket, out);
  check_launch("k_mdot");
}
template <int KB, bool NORM>
void launch_maxpy(Ctx *c, int64_t n, int k, double sign, double *w, const double *V, int64_t ld, const double *h, double *out) {
  bool v2 = aligned16(w) && aligned16(V) && (ld % 2 == 0);
  LaunchScope ls(c, "maxpy");
  int grid = reduce_grid(c, v2 ? (n + 1) / 2 : n, 6);
  if (v2) k_maxpy<KB, true, NORM><<<grid, 256, 0, c->stream>>>(n, k, sign, w, V, ld, h, c->d_partials, c->d_ticket, out);
  else k_maxpy<KB, false, NORM><<<grid, 256, 0, c->stream>>>(n, k, sign, w, V, ld, h, c->d_partials, c->d_ticket, out);
  check_launch("k_maxpy");
}

} // namespace

void vec_set(Ctx *c, int64_t n, double a, double *y) { MapArgs m{a, 0, 0, nullptr, nullptr, nullptr, nullptr, y}; launch_map<OP_SET>(c, n, m); }
void vec_copy(Ctx *c, int64_t n, const double *x, double *y) {
  if (x == y || n <= 0) return;
  MapArgs m{0, 0, 0, x, nullptr, nullptr, nullptr, y};
  launch_map<OP_COPY>(c, n, m);
}
void vec_scale(Ctx *c, int64_t n, double a, double *y) { MapArgs m{a, 0, 0, nullptr, nullptr, nullptr, nullptr, y}; launch_map<OP_SCALE>(c, n, m); }
void vec_axpy(Ctx *c, int64_t n, double a, const double *x, double *y) { MapArgs m{a, 0, 0, x, nullptr, nullptr, nullptr, y}; launch_map<OP_AXPY>(c, n, m); }
void vec_aypx(Ctx *c, int64_t n, double a, const double *x, double *y) { MapArgs m{a, 0, 0, x, nullptr, nullptr, nullptr, y}; launch_map<OP_AYPX>(c, n, m); }
void vec_waxpy(Ctx *c, int64_t n, double a, const double *x, const double *y, double *w) { MapArgs m{a, 0, 0, x, y, nullptr, nullptr, w}; launch_map<OP_WAXPY>(c, n, m); }
void vec_axpbypcz(Ctx *c, int64_t n, double a, const double *x, double b, const double *y, double cc, const double *z, double *w) {
  MapArgs m{a, b, cc, x, y, z, nullptr, w};
  launch_map<OP_AXPBYPCZ>(c, n, m);
}
void vec_cheb_update(Ctx *c, int64_t n, double a, const double *x, double b, const double *y, double cc, const double *d, const double *z, double *w) {
  MapArgs m{a, b, cc, x, y, z, d, w};
  if (d) launch_map<OP_CHEB_D>(c, n, m); else launch_map<OP_CHEB>(c, n, m);
}
void vec_pointwise_mult(Ctx *c, int64_t n, const double *x, const double *y, double *w) { MapArgs m{0, 0, 0, x, y, nullptr, nullptr, w}; launch_map<OP_PMULT>(c, n, m); }
void vec_reciprocal_safe(Ctx *c, int64_t n, double *d) { MapArgs m{0, 0, 0, nullptr, nullptr, nullptr, nullptr, d}; launch_map<OP_RECIP>(c, n, m); }
void vec_hash(Ctx *c, int64_t n, double *v) {
  if (n <= 0) return;
  LaunchScope ls(c, "vec");
  k_hash<<<stream_grid(c, n), 256, 0, c->stream>>>(n, v);
  check_launch("k_hash");
}
void vec_hash_natural(Ctx *c, int xs, int ys, int xm, int ym, int M, int dof, double *v) {
  const int64_t n = (int64_t)xm * ym * dof;
  if (n <= 0) return;
  LaunchScope ls(c, "vec");
  k_hash_natural<<<stream_grid(c, n), 256, 0, c->stream>>>(xs, ys, xm, ym, M, dof, v);
  check_launch("k_hash_natural");
}
void vec_scatter_set(Ctx *c, int64_t n, const int *idx, double val, double *y) {
  if (n <= 0) return;
  LaunchScope ls(c, "vec");
  k_scatter_set<<<stream_grid(c, n), 256, 0, c->stream>>>(n, idx, val, y);
  check_launch("k_scatter_set");
}

void vec_permute_scatter(Ctx *c, int64_t n, const int *map, const double *in, double *out) {
  if (n <= 0) return;
  LaunchScope ls(c, "vec");
  k_permute_scatter<<<stream_grid(c, n), 256, 0, c->stream>>>(n, map, in, out);
  check_launch("k_permute_scatter");
}
void vec_permute_gather(Ctx *c, int64_t n, const int *map, const double *in, double *out) {
  if (n <= 0) return;
  LaunchScope ls(c, "vec");
  k_permute_gather<<<stream_grid(c, n), 256, 0, c->stream>>>(n, map, in, out);
  check_launch("k_permute_gather");
}

void vec_mdot(Ctx *c, int64_t n, int k, const double *w, const double *V, int64_t ld, double *out) {
  // passes of at most 16 basis vectors (w is re-read once per pass)
  for (int j0 = 0; j0 < k; j0 += 16) {
    int kk = k - j0 < 16 ? k - j0 : 16;
    const double *Vj = V + (size_t)j0 * ld;
    if (kk <= 1) launch_mdot<1>(c, n, kk, w, Vj, ld, out + j0);
    else if (kk <= 2) launch_mdot<2>(c, n, kk, w, Vj, ld, out + j0);
    else if (kk <= 4) launch_mdot<4>(c, n, kk, w, Vj, ld, out + j0);
    else if (kk <= 8) launch_mdot<8>(c, n, kk, w, Vj, ld, out + j0);
    else launch_mdot<16>(c, n, kk, w, Vj, ld, out + j0);
  }
}
void vec_dot(Ctx *c, int64_t n, const double *x, const double *y, double *out) { launch_mdot<1>(c, n, 1, x, y, 0, out); }

void vec_maxpy_norm2(Ctx *c, int64_t n, int k, double *w, const double *V, int64_t ld, const double *h, double *out) {
  int j0 = 0;
  for (; k - j0 > 16; j0 += 16) launch_maxpy<16, false>(c, n, 16, -1.0, w, V + (size_t)j0 * ld, ld, h + j0, out);
  int kk = k - j0;
  const double *Vj = V + (size_t)j0 * ld;
  if (kk <= 1) launch_maxpy<1, true>(c, n, kk, -1.0, w, Vj, ld, h + j0, out);
  else if (kk <= 2) launch_maxpy<2, true>(c, n, kk, -1.0, w, Vj, ld, h + j0, out);
  else if (kk <= 4) launch_maxpy<4, true>(c, n, kk, -1.0, w, Vj, ld, h + j0, out);
  else if (kk <= 8) launch_maxpy<8, true>(c, n, kk, -1.0, w, Vj, ld, h + j0, out);
  else launch_maxpy<16, true>(c, n, kk, -1.0, w, Vj, ld, h + j0, out);
}
void vec_maxpy(Ctx *c, int64_t n, int k, double *w, const double *V, int64_t ld, const double *coef) {
  for (int j0 = 0; j0 < k; j0 += 16) {
    int kk = k - j0 < 16 ? k - j0 : 16;
    const double *Vj = V + (size_t)j0 * ld;
    if (kk <= 1) launch_maxpy<1, false>(c, n, kk, 1.0, w, Vj, ld, coef + j0, nullptr);
    else if (kk <= 4) launch_maxpy<4, false>(c, n, kk, 1.0, w, Vj, ld, coef + j0, nullptr);
    else if (kk <= 8) launch_maxpy<8, false>(c, n, kk, 1.0, w, Vj, ld, coef + j0, nullptr);
    else launch_maxpy<16, false>(c, n, kk, 1.0, w, Vj, ld, coef + j0, nullptr);
  }
}
void vec_scale_inv_sqrt(Ctx *c, int64_t n, const double *nrm2, const double *x, double *y) {
  if (n <= 0) return;
  bool v2 = aligned16(x) && aligned16(y);
  LaunchScope ls(c, "vec");
  if (v2) k_scale_inv_sqrt<true><<<stream_grid(c, (n + 1) / 2), 256, 0, c->stream>>>(n, nrm2, x, y);
  else k_scale_inv_sqrt<false><<<stream_grid(c, n), 256, 0, c->stream>>>(n, nrm2, x, y);
  check_launch("k_scale_inv_sqrt");
}

} // namespace b200sp
