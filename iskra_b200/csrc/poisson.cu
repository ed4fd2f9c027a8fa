// Device field solve rho -> phi -> E for the reference's 5-point operator
// (FiniteDifferenceMethod/src/generalized_poisson.jl).  The reference assembles a dense nn x nn
// matrix and LU-factors it every step (:34-68, :201-203, :372-378); that is 8.8 TB at 1025^2
// nodes.  The operator is separable whenever the Dirichlet nodes are whole edges:
//     A = (T_i (x) I + I (x) T_j) / dx^2          (dx == dy required, :65 "TODO")
// with T = path / ring / Dirichlet-lifted tridiagonals, so we solve it exactly by
//     (1) an orthonormal eigen-transform along one axis (closed-form sine/cosine bases;
//         FP64 GEMM, or an FFT-based DST-I when the interior length + 1 is a power of two),
//     (2) one (cyclic) tridiagonal Thomas solve per mode along the other axis, with the
//         elimination factors precomputed once per boundary structure,
//     (3) the inverse transform.
// Arbitrary Dirichlet masks on small grids fall back to a dense inverse computed once per
// boundary structure on the host and a device GEMV per step.
// Fully periodic / all-open operators are singular; the reference's LU then returns a solution
// polluted by the null vector.  Here the constant mode of the right-hand side is
// projected out and phi has zero mean -- E is unaffected by the gauge.
#include <cmath>
#include <cstring>

#include "common.cuh"

namespace {

constexpr int TPB = 256;
enum AxisKind { AX_PATH = 0, AX_RING = 1, AX_DD = 2, AX_DL = 3, AX_DR = 4 };

struct AxisInfo {
  AxisKind kind;
  int n;       // nodes
  int lo, m;   // unknown range [lo, lo+m)
};

// ---------------------------------------------------------------------------------------------
// host: dense assembly exactly as the reference (debug, parity and the dense fallback)
// ---------------------------------------------------------------------------------------------
void assemble_dense(const iskb_ctx *c, std::vector<double> &A, std::vector<double> &b) {
  const PoissonState &ps = c->ps;
  const int nx = c->g.nx, ny = c->g.ny;
  const int64_t nnodes = (int64_t)nx * ny;
  const int64_t nn = nnodes + ps.n_sigma;                // add_new_dof appends one row/column per sigma dof :217-230
  A.assign((size_t)(nn * nn), 0.0);
  b.assign((size_t)nn, 0.0);
  auto at = [&](int64_t r, int64_t col) -> double & { return A[(size_t)(r + col * nn)]; };
  if (ps.has_custom)                                     // operator assembled by the caller (axial grid :70-199)
    for (int64_t col = 0; col < nnodes; ++col)
      for (int64_t r = 0; r < nnodes; ++r) at(r, col) = ps.custom_A[(size_t)(r + col * nnodes)];
  for (int j = 0; j < ny && !ps.has_custom; ++j)
    for (int i = 0; i < nx; ++i) {                       // :42-61
      const int64_t r = i + (int64_t)j * nx;
      if (i < nx - 1) { at(r, r) -= 1.0; at(r, r + 1) += 1.0; }
      if (i > 0) { at(r, r) -= 1.0; at(r, r - 1) += 1.0; }
      if (j < ny - 1) { at(r, r) -= 1.0; at(r, r + nx) += 1.0; }
      if (j > 0) { at(r, r) -= 1.0; at(r, r - nx) += 1.0; }
    }
  const double d2 = c->g.dx * c->g.dx;                   // :65
  if (!ps.has_custom)
    for (auto &v : A) v /= d2;
  for (int k = 0; k < ps.n_sigma; ++k) {                 // :227-228
    at(nnodes + k, nnodes + k) = 1.0;
    b[(size_t)(nnodes + k)] = ps.sigma[(size_t)k];
  }
  if (ps.periodic_j && !ps.has_custom) {                 // apply_periodic(ps, 1) :291-306
    const double cc = (0.5 + 0.5) / (c->g.dx * c->g.dx);
    for (int jj = 0; jj < 2; ++jj) {
      const int j = jj == 0 ? 0 : ny - 1;
      for (int i = 0; i < nx; ++i) {
        const int64_t r = i + (int64_t)j * nx;
        if (j == ny - 1) { at(r, r) -= cc; at(r, i) += cc; }
        if (j == 0) { at(r, r) -= cc; at(r, i + (int64_t)(ny - 1) * nx) += cc; }
      }
    }
  }
  if (ps.periodic_i && !ps.has_custom) {                 // apply_periodic(ps, 2) :308-323
    const double cc = (0.5 + 0.5) / (c->g.dy * c->g.dy);
    for (int j = 0; j < ny; ++j)
      for (int ii = 0; ii < 2; ++ii) {
        const int i = ii == 0 ? 0 : nx - 1;
        const int64_t r = i + (int64_t)j * nx;
        if (i == nx - 1) { at(r, r) -= cc; at(r, (int64_t)j * nx) += cc; }
        if (i == 0) { at(r, r) -= cc; at(r, (nx - 1) + (int64_t)j * nx) += cc; }
      }
  }
  if (!ps.neu_kind.empty()) {                            // apply_neumann :235-269 (eps_r == 1)
    const double dx = c->g.dx, dy = c->g.dy;
    for (int64_t r = 0; r < nnodes; ++r) {
      const int kind = ps.neu_kind[(size_t)r];
      if (!kind) continue;
      const int i = (int)(r % nx), j = (int)(r / nx);
      const int64_t ri = ps.neu_i2[(size_t)r] + (int64_t)j * nx, s = nnodes + ps.neu_dof[(size_t)r];
      for (int64_t col = 0; col < nn; ++col) at(r, col) = 0.0;
      if (kind == 1) {                                   // :248-255
        at(r, r) -= 2 * 1.0 / dx;
        at(r, ri) += 2 * 1.0 / dx;
        at(r, s) += 2;
      } else {                                           // :256-267
        const int64_t rj = i + (int64_t)ps.neu_j2[(size_t)r] * nx;
        at(r, r) -= 4 * 1.0 / (3 * (dx * dx));
        at(r, r) -= 4 * 1.0 / (3 * (dy * dy));
        at(r, ri) += 4 * 1.0 / (3 * (dx * dx));
        at(r, rj) += 4 * 1.0 / (3 * (dy * dy));
        at(r, s) += (2.0 / 3.0) * (dx + dy) / (dx * dy);
      }
    }
  }
  for (int64_t r = 0; r < nnodes; ++r)
    if (ps.isdir[(size_t)r]) {                           // apply_dirichlet :205-215
      for (int64_t col = 0; col < nn; ++col) at(r, col) = 0.0;
      at(r, r) = 1.0;
      b[(size_t)r] = ps.dval[(size_t)r];
    }
}

// Gauss-Jordan inverse with partial pivoting, column-major.  Returns false if singular.
bool invert_dense(std::vector<double> &A, int64_t n) {
  std::vector<double> inv((size_t)(n * n), 0.0);
  for (int64_t k = 0; k < n; ++k) inv[(size_t)(k + k * n)] = 1.0;
  double amax = 0.0;
  for (double v : A) amax = std::fmax(amax, std::fabs(v));
  std::vector<double> colk((size_t)n);
  for (int64_t k = 0; k < n; ++k) {
    int64_t piv = k;
    double best = std::fabs(A[(size_t)(k + k * n)]);
    for (int64_t r = k + 1; r < n; ++r) {
      const double a = std::fabs(A[(size_t)(r + k * n)]);
      if (a > best) { best = a; piv = r; }
    }
    if (!(best > 1e-13 * amax)) return false;
    if (piv != k)
      for (int64_t col = 0; col < n; ++col) {
        std::swap(A[(size_t)(k + col * n)], A[(size_t)(piv + col * n)]);
        std::swap(inv[(size_t)(k + col * n)], inv[(size_t)(piv + col * n)]);
      }
    const double ip = 1.0 / A[(size_t)(k + k * n)];
    for (int64_t col = 0; col < n; ++col) {
      A[(size_t)(k + col * n)] *= ip;
      inv[(size_t)(k + col * n)] *= ip;
    }
    for (int64_t r = 0; r < n; ++r) colk[(size_t)r] = A[(size_t)(r + k * n)];
    for (int64_t col = 0; col < n; ++col) {
      const double ak = A[(size_t)(k + col * n)], ik = inv[(size_t)(k + col * n)];
      if (ak == 0.0 && ik == 0.0) continue;
      double *Ac = &A[(size_t)(col * n)], *Ic = &inv[(size_t)(col * n)];
      for (int64_t r = 0; r < n; ++r) {
        if (r == k) continue;
        const double f = colk[(size_t)r];
        if (f == 0.0) continue;
        Ac[r] -= f * ak;
        Ic[r] -= f * ik;
      }
    }
  }
  A.swap(inv);
  return true;
}

// ---- the same Gauss-Jordan on the device: same pivots (first largest |a| of the column), same arithmetic per element
// (one multiply and one subtract, never contracted), so the inverse is bit-identical to invert_dense above.  The host version
// is memory bound on one core (5 s at 2145 unknowns, 5 min at 8125 -- problem/13_seed.jl's grid); here one pivot is a
// rank-1 update of 2 n^2 elements at HBM speed.
__global__ void __launch_bounds__(1024) k_gj_pivot(const double *__restrict__ A, int64_t n, int64_t k, double amax, int64_t *piv,
                                                   double *pval, int *singular) {
  __shared__ double s_v[1024];
  __shared__ int64_t s_i[1024];
  double best = -1.0;
  int64_t bi = n;
  for (int64_t r = k + threadIdx.x; r < n; r += 1024) {
    const double a = fabs(A[r + k * n]);
    if (a > best) { best = a; bi = r; }          // rows ascend per thread: the first largest stays
  }
  s_v[threadIdx.x] = best; s_i[threadIdx.x] = bi;
  __syncthreads();
  for (int d = 512; d > 0; d >>= 1) {
    if ((int)threadIdx.x < d) {
      const double ov = s_v[threadIdx.x + d];
      const int64_t oi = s_i[threadIdx.x + d];
      if (ov > s_v[threadIdx.x] || (ov == s_v[threadIdx.x] && oi < s_i[threadIdx.x])) { s_v[threadIdx.x] = ov; s_i[threadIdx.x] = oi; }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    if (!(s_v[0] > 1e-13 * amax)) { *singular = 1; *piv = k; *pval = 1.0; }
    else { *piv = s_i[0]; *pval = A[s_i[0] + k * n]; }
  }
}
// swap rows k and piv, scale row k by 1/pivot; the scaled row is kept contiguous for the update
__global__ void k_gj_rows(double *A, double *inv, int64_t n, int64_t k, const int64_t *__restrict__ piv, const double *__restrict__ pval,
                          double *rowA, double *rowI) {
  const int64_t p = *piv;
  const double ip = 1.0 / *pval;
  for (int64_t col = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; col < n; col += (int64_t)gridDim.x * blockDim.x) {
    double a = A[p + col * n], b = inv[p + col * n];
    if (p != k) {
      A[p + col * n] = A[k + col * n];
      inv[p + col * n] = inv[k + col * n];
    }
    a = __dmul_rn(a, ip);
    b = __dmul_rn(b, ip);
    A[k + col * n] = a; inv[k + col * n] = b;
    rowA[col] = a; rowI[col] = b;
  }
}
__global__ void k_gj_colk(const double *__restrict__ A, int64_t n, int64_t k, double *colk) {
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x) colk[r] = A[r + k * n];
}
__global__ void __launch_bounds__(256) k_gj_update(double *A, double *inv, int64_t n, int64_t k, const double *__restrict__ colk,
                                                   const double *__restrict__ rowA, const double *__restrict__ rowI) {
  const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r >= n || r == k) return;
  const double f = colk[r];
  if (f == 0.0) return;
  for (int64_t col = blockIdx.y; col < n; col += gridDim.y) {
    const double ak = rowA[col], ik = rowI[col];
    if (ak != 0.0) A[r + col * n] = __dsub_rn(A[r + col * n], __dmul_rn(f, ak));
    if (ik != 0.0) inv[r + col * n] = __dsub_rn(inv[r + col * n], __dmul_rn(f, ik));
  }
}
__global__ void k_gj_identity(double *inv, int64_t n) {
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) inv[k + k * n] = 1.0;
}

// A (host, n*n column-major) -> its inverse in a new device buffer *d_inv_out
int32_t invert_dense_device(iskb_ctx *c, const std::vector<double> &A, int64_t n, double **d_inv_out) {
  double amax = 0.0;
  for (double v : A) amax = std::fmax(amax, std::fabs(v));
  double *dA = nullptr, *dI = nullptr, *d_work = nullptr, *d_pval = nullptr;
  int64_t *d_piv = nullptr;
  int *d_sing = nullptr;
  const size_t bytes = (size_t)(n * n) * sizeof(double);
  CU_TRY(cudaMalloc(&dA, bytes));
  CU_TRY(cudaMalloc(&dI, bytes));
  CU_TRY(cudaMalloc(&d_work, 3 * (size_t)n * sizeof(double)));
  CU_TRY(cudaMalloc(&d_pval, sizeof(double)));
  CU_TRY(cudaMalloc(&d_piv, sizeof(int64_t)));
  CU_TRY(cudaMalloc(&d_sing, sizeof(int)));
  cudaStream_t st = c->stream;
  CU_TRY(cudaMemcpyAsync(dA, A.data(), bytes, cudaMemcpyHostToDevice, st));
  CU_TRY(cudaMemsetAsync(dI, 0, bytes, st));
  CU_TRY(cudaMemsetAsync(d_sing, 0, sizeof(int), st));
  const int nb = (int)((n + 255) / 256);
  k_gj_identity<<<nb, 256, 0, st>>>(dI, n);
  double *colk = d_work, *rowA = d_work + n, *rowI = d_work + 2 * n;
  const dim3 ug((unsigned)nb, (unsigned)std::min<int64_t>(n, 1024));
  for (int64_t k = 0; k < n; ++k) {
    k_gj_pivot<<<1, 1024, 0, st>>>(dA, n, k, amax, d_piv, d_pval, d_sing);
    k_gj_rows<<<nb, 256, 0, st>>>(dA, dI, n, k, d_piv, d_pval, rowA, rowI);
    k_gj_colk<<<nb, 256, 0, st>>>(dA, n, k, colk);
    k_gj_update<<<ug, 256, 0, st>>>(dA, dI, n, k, colk, rowA, rowI);
  }
  int sing = 0;
  CU_TRY(cudaMemcpyAsync(&sing, d_sing, sizeof(int), cudaMemcpyDeviceToHost, st));
  CU_TRY(cudaStreamSynchronize(st));
  CU_TRY(cudaGetLastError());
  cudaFree(dA); cudaFree(d_work); cudaFree(d_pval); cudaFree(d_piv); cudaFree(d_sing);
  if (sing) { cudaFree(dI); return iskb_fail(ISKB_E_SINGULAR, "dense Poisson operator is singular"); }
  *d_inv_out = dI;
  return ISKB_OK;
}

// ---------------------------------------------------------------------------------------------
// host: closed-form orthonormal eigen-decompositions of the 1-D operators
// ---------------------------------------------------------------------------------------------
void axis_eigen(const AxisInfo &ax, std::vector<double> &V, std::vector<double> &lam) {
  const int m = ax.m;
  V.assign((size_t)m * m, 0.0);
  lam.assign((size_t)m, 0.0);
  const long double PI = 3.141592653589793238462643383279502884L;
  auto set = [&](int p, int k, long double v) { V[(size_t)p + (size_t)k * m] = (double)v; };
  // angles are reduced in exact integer arithmetic before the long-double sin/cos
  auto ang = [&](long long num, long long den) { return PI * (long double)(num % (2 * den)) / (long double)den; };
  switch (ax.kind) {
    case AX_RING: {
      const int n = m;
      int col = 0;
      for (int k = 0; k <= n / 2; ++k) {
        const long double l = -2.0L + 2.0L * cosl(ang(2LL * k, n));
        if (k == 0 || (n % 2 == 0 && k == n / 2)) {
          const long double s = 1.0L / sqrtl((long double)n);
          for (int p = 0; p < n; ++p) set(p, col, s * cosl(ang(2LL * k * p, n)));
          lam[(size_t)col++] = (double)l;
        } else {
          const long double s = sqrtl(2.0L / n);
          for (int p = 0; p < n; ++p) set(p, col, s * cosl(ang(2LL * k * p, n)));
          lam[(size_t)col++] = (double)l;
          for (int p = 0; p < n; ++p) set(p, col, s * sinl(ang(2LL * k * p, n)));
          lam[(size_t)col++] = (double)l;
        }
      }
      break;
    }
    case AX_PATH: {
      const int n = m;
      for (int k = 0; k < n; ++k) {
        const long double s = k == 0 ? 1.0L / sqrtl((long double)n) : sqrtl(2.0L / n);
        for (int p = 0; p < n; ++p) set(p, k, s * cosl(ang((long long)k * (2 * p + 1), 2LL * n)));
        lam[(size_t)k] = (double)(-2.0L + 2.0L * cosl(ang(k, n)));
      }
      break;
    }
    case AX_DD: {
      const long double s = sqrtl(2.0L / (m + 1));
      for (int k = 1; k <= m; ++k) {
        for (int p = 1; p <= m; ++p) set(p - 1, k - 1, s * sinl(ang((long long)k * p, m + 1)));
        lam[(size_t)(k - 1)] = (double)(-2.0L + 2.0L * cosl(ang(k, m + 1)));
      }
      break;
    }
    case AX_DL:
    case AX_DR: {
      const long double s = 2.0L / sqrtl(2.0L * m + 1.0L);
      for (int k = 1; k <= m; ++k) {
        for (int p = 1; p <= m; ++p) {
          const int row = ax.kind == AX_DL ? p - 1 : m - p;
          set(row, k - 1, s * sinl(ang((long long)(2 * k - 1) * p, 2LL * m + 1)));
        }
        lam[(size_t)(k - 1)] = (double)(-2.0L + 2.0L * cosl(ang(2 * k - 1, 2LL * m + 1)));
      }
      break;
    }
  }
}

// diagonal of the tridiagonal along an axis restricted to its unknown range (off-diagonals = 1)
void axis_diag(const AxisInfo &ax, std::vector<double> &d) {
  d.assign((size_t)ax.m, -2.0);
  if (ax.kind == AX_PATH) { d[0] = -1.0; d[(size_t)ax.m - 1] = -1.0; if (ax.m == 1) d[0] = 0.0; }
  if (ax.kind == AX_DL) d[(size_t)ax.m - 1] = -1.0;
  if (ax.kind == AX_DR) d[0] = -1.0;
}

// ---------------------------------------------------------------------------------------------
// device kernels
// ---------------------------------------------------------------------------------------------
struct SolveDims {
  int nx, ny;
  int transposed;   // working axis a is j when set
  int a0, ma, b0, mb;
  double scale;     // dx^2 / eps0
};

__device__ __forceinline__ int64_t node_of(const SolveDims &s, int a, int b) {
  return s.transposed ? (int64_t)b + (int64_t)a * s.nx : (int64_t)a + (int64_t)b * s.nx;
}

// R[a,b] = dx^2 * b_rhs - (Dirichlet neighbours), b_rhs = (-rho)/eps0   (:375 with f = -rho)
__global__ void k_build_rhs(SolveDims s, const double *__restrict__ rho, const uint8_t *__restrict__ isdir,
                            const double *__restrict__ dval, double *R) {
  const int64_t tot = (int64_t)s.ma * s.mb;
  const int na = s.transposed ? s.ny : s.nx, nb = s.transposed ? s.nx : s.ny;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < tot;
       t += (int64_t)gridDim.x * blockDim.x) {
    const int ka = (int)(t % s.ma), kb = (int)(t / s.ma);
    const int a = s.a0 + ka, b = s.b0 + kb;
    double r = -(rho[node_of(s, a, b)]) * s.scale;
    if (a - 1 >= 0 && isdir[node_of(s, a - 1, b)]) r -= dval[node_of(s, a - 1, b)];
    if (a + 1 < na && isdir[node_of(s, a + 1, b)]) r -= dval[node_of(s, a + 1, b)];
    if (b - 1 >= 0 && isdir[node_of(s, a, b - 1)]) r -= dval[node_of(s, a, b - 1)];
    if (b + 1 < nb && isdir[node_of(s, a, b + 1)]) r -= dval[node_of(s, a, b + 1)];
    R[t] = r;
  }
}

__global__ void k_store_phi(SolveDims s, const double *__restrict__ X, const uint8_t *__restrict__ isdir,
                            const double *__restrict__ dval, double *phi) {
  const int64_t nn = (int64_t)s.nx * s.ny;
  for (int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; n < nn;
       n += (int64_t)gridDim.x * blockDim.x) {
    if (isdir[n]) {
      phi[n] = dval[n];
      continue;
    }
    const int i = (int)(n % s.nx), j = (int)(n / s.nx);
    const int a = s.transposed ? j : i, b = s.transposed ? i : j;
    phi[n] = X[(int64_t)(a - s.a0) + (int64_t)(b - s.b0) * s.ma];
  }
}

// C[M x N] = A[M x K] * B[K x N], all column-major, FP64, 64x64x16 tiles, 4x4 per thread.
__global__ void __launch_bounds__(256) k_dgemm(int M, int N, int K, const double *__restrict__ A, int lda,
                                               const double *__restrict__ B, int ldb, double *C, int ldc) {
  __shared__ double As[16][64 + 4];
  __shared__ double Bs[16][64 + 4];
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
  const int m0 = blockIdx.x * 64, n0 = blockIdx.y * 64;
  double acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += 16) {
    for (int e = threadIdx.x; e < 64 * 16; e += 256) {
      const int r = e % 64, kk = e / 64;
      const int gm = m0 + r, gk = k0 + kk;
      As[kk][r] = (gm < M && gk < K) ? A[(int64_t)gm + (int64_t)gk * lda] : 0.0;
    }
    for (int e = threadIdx.x; e < 64 * 16; e += 256) {
      const int kk = e % 16, cidx = e / 16;
      const int gk = k0 + kk, gn = n0 + cidx;
      Bs[kk][cidx] = (gk < K && gn < N) ? B[(int64_t)gk + (int64_t)gn * ldb] : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      double a[4], b[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) a[q] = As[kk][tx + 16 * q];
#pragma unroll
      for (int q = 0; q < 4; ++q) b[q] = Bs[kk][ty + 16 * q];
#pragma unroll
      for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int r = 0; r < 4; ++r) acc[q][r] = fma(a[q], b[r], acc[q][r]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int q = 0; q < 4; ++q)
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int gm = m0 + tx + 16 * q, gn = n0 + ty + 16 * r;
      if (gm < M && gn < N) C[(int64_t)gm + (int64_t)gn * ldc] = acc[q][r];
    }
}

// Thomas solve along b, ONE WARP PER MODE k on mode-major data (W[q + k*mb], q contiguous).
// Both sweeps are first-order linear recurrences, y_q = A_q*y_{q-1} + B_q, so a tile of 32
// consecutive q is resolved with a 5-step warp scan over the composed affine maps
// (A2,B2) o (A1,B1) = (A2*A1, A2*B1 + B2) and the tiles are chained through a carried value:
//   forward   y_q = (r_q - y_{q-1}) * m_q          A = -m_q, B = r_q*m_q      (m precomputed)
//   backward  x_q = y_q - m_q * x_{q+1}            A = -m_q, B = y_q
//   cyclic    x_q -= qt_q * f, f = (x_0 + x_{n-1}/gamma)*qden   (Sherman-Morrison)
// The singular mode (all-periodic / all-open operator) pins x_0 = 0 on the mean-projected rhs and
// removes the mean of the result.  Depth per sweep: mb/32 tiles x ~60 cycles instead of mb steps.
__device__ __forceinline__ void affine_scan_up(double &A, double &B, int lane) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const double Ap = __shfl_up_sync(0xffffffffu, A, d), Bp = __shfl_up_sync(0xffffffffu, B, d);
    if (lane >= d) {
      B = fma(A, Bp, B);
      A = A * Ap;
    }
  }
}

__global__ void __launch_bounds__(128) k_thomas_warp(int ma, int mb, int cyclic, int singular_mode,
                                                     const double *__restrict__ mt, const double *__restrict__ msing,
                                                     const double *__restrict__ qt, const double *__restrict__ qden,
                                                     const double *__restrict__ gam, double *W) {
  const int lane = threadIdx.x & 31;
  const int k = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (k >= ma) return;
  double *w = W + (int64_t)k * mb;
  const bool sing = k == singular_mode;
  const double *m = sing ? msing : mt + (int64_t)k * mb;
  double mean = 0.0;
  if (sing) {
    for (int q = lane; q < mb; q += 32) mean += w[q];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) mean += __shfl_xor_sync(0xffffffffu, mean, d);
    mean /= mb;
  }
  const int ntile = (mb + 31) / 32;
  // forward sweep.  The recurrence is serial through `carry`, the loads are not: the operands of tile
  // t+1 are fetched before the scan of tile t so that their latency hides behind it.
  double carry = 0.0;
  double mq_n = 0.0, wq_n = 0.0;
  if (lane < mb) { mq_n = m[lane]; wq_n = w[lane]; }
  for (int t = 0; t < ntile; ++t) {
    const int q = t * 32 + lane;
    const double mq = mq_n, wq = wq_n;
    if (q + 32 < mb) { mq_n = m[q + 32]; wq_n = w[q + 32]; }
    double A = 1.0, B = 0.0;   // identity for padding lanes
    if (q < mb) {
      A = -mq;
      B = (wq - mean) * mq;
      if (sing && q == 0) { A = 0.0; B = 0.0; }   // pinned x_0 = 0
    }
    affine_scan_up(A, B, lane);
    const double y = fma(A, carry, B);
    if (q < mb) w[q] = y;
    carry = __shfl_sync(0xffffffffu, y, 31);
  }
  __syncwarp();
  // backward sweep (reversed index r = mb-1-q so that the recurrence runs upward in r)
  carry = 0.0;
  if (lane < mb) { mq_n = m[mb - 1 - lane]; wq_n = w[mb - 1 - lane]; }
  for (int t = 0; t < ntile; ++t) {
    const int r = t * 32 + lane, q = mb - 1 - r;
    const double mq = mq_n, wq = wq_n;
    if (r + 32 < mb) { mq_n = m[q - 32]; wq_n = w[q - 32]; }
    double A = 1.0, B = 0.0;
    if (r < mb) {
      A = (r == 0) ? 0.0 : -mq;
      B = wq;
      if (sing && q == 0) { A = 0.0; B = 0.0; }
    }
    affine_scan_up(A, B, lane);
    const double x = fma(A, carry, B);
    if (r < mb) w[q] = x;
    carry = __shfl_sync(0xffffffffu, x, 31);
  }
  __syncwarp();
  if (sing) {
    double sum = 0.0;
    for (int q = lane; q < mb; q += 32) sum += w[q];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, d);
    const double xm = sum / mb;
    for (int q = lane; q < mb; q += 32) w[q] -= xm;
  } else if (cyclic) {
    const double f = (w[0] + w[mb - 1] / gam[k]) * qden[k];
    __syncwarp();
    const double *qk = qt + (int64_t)k * mb;
    for (int q = lane; q < mb; q += 32) w[q] -= qk[q] * f;
  }
}

// out[c + r*cols] = in[r + c*rows]  (tiled through shared memory, coalesced on both sides)
__global__ void k_transpose(int rows, int cols, const double *__restrict__ in, double *__restrict__ out) {
  __shared__ double tile[32][33];
  const int r0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int dc = threadIdx.y; dc < 32; dc += blockDim.y) {
    const int r = r0 + threadIdx.x, c = c0 + dc;
    if (r < rows && c < cols) tile[dc][threadIdx.x] = in[(int64_t)r + (int64_t)c * rows];
  }
  __syncthreads();
  for (int dr = threadIdx.y; dr < 32; dr += blockDim.y) {
    const int r = r0 + dr, c = c0 + threadIdx.x;
    if (r < rows && c < cols) out[(int64_t)c + (int64_t)r * cols] = tile[threadIdx.x][dr];
  }
}

// DST-I along the contiguous axis by a complex FFT of length M = 2(m+1) in shared memory.
// Two real columns are packed into one complex sequence (odd extension): with
// Z = FFT(z_a + i z_b), the sine sums are S_a[k] = -Im Z[k]/2 and S_b[k] = Re Z[k]/2.
// out[k-1, col] = scale * S[k], scale = sqrt(2/(m+1)) (orthonormal, so forward == inverse).
// IN = 1: the right-hand side is formed on the fly from rho (what k_build_rhs writes: dx^2*(-rho)/eps0 minus the
//         Dirichlet neighbours, generalized_poisson.jl:375 with f = -rho) -- one pass over rho less per solve;
// IN = 2: mode-space input with the Sherman-Morrison correction of the cyclic Thomas solve applied on the way in
//         (x -= q*f, see k_thomas_seg);
// OUT = 1: the result goes straight into phi (k_store_phi's job for the unknown nodes).
// Shared-memory index of element i of the transform: one 16-byte slot is skipped after every 8, 64 and 512 elements.
// 128-bit accesses are served per quarter warp (8 lanes, 8 groups of 4 banks): with this padding the 8 lanes hit 8
// different groups both in the bit-reversed scatter of the load phase (addresses 128*brev(l) + c) and in the
// stride-4 accesses of the first butterfly passes -- unpadded, those were 8-way and 4-way conflicts.
__device__ __forceinline__ int fpad(int i) { return i + (i >> 3) + (i >> 6) + (i >> 9); }
static inline int fpad_host(int i) { return i + (i >> 3) + (i >> 6) + (i >> 9); }

struct FftIo {
  SolveDims s;
  const double *rho;
  const uint8_t *isdir;
  const double *dval;
  const double *qt;     // [k + q*ma] Sherman-Morrison vector (IN = 2, cyclic)
  const double *f;      // [ma] per-mode factor (IN = 2, cyclic); nullptr: no correction
  double *phi;
  // IN = 1 with rho still in the fixed-point sums of the fused tiled step (particles.cu, rho_materialize): rho is formed on
  // the way in with the expression of k_rho_fixed_local, so the solve reads the same bits whether or not rho was materialised
  int ns;               // 0: read io.rho
  const long long *u[8];
  long long z[8];
  const double *V;
  double q0s;
};

template <int IN, int OUT>
__global__ void __launch_bounds__(512) k_dst_fft(int m, int ncols, int log2M, const double2 *__restrict__ tw,
                                                 const double *__restrict__ in, double *__restrict__ out,
                                                 double scale, FftIo io) {
  extern __shared__ double2 zs[];
  const int M = 1 << log2M;
  const int ca = blockIdx.x * 2, cb = ca + 1;
  const bool has_b = cb < ncols;   // (mb odd: the last block transforms one column)
  auto fetch = [&](int col, int p) -> double {   // element p (1-based) of column col
    if (IN == 1) {
      const SolveDims &s = io.s;
      const int na = s.transposed ? s.ny : s.nx, nb = s.transposed ? s.nx : s.ny;
      const int a = s.a0 + p - 1, b = s.b0 + col;
      double rho_n;
      if (io.ns) {
        const int64_t n = node_of(s, a, b);
        long long acc = 0;
        for (int q = 0; q < io.ns; ++q) acc += io.z[q] * io.u[q][n];
        rho_n = __ddiv_rn(__dmul_rn((double)acc, io.q0s), io.V[n]);
      } else {
        rho_n = io.rho[node_of(s, a, b)];
      }
      double r = -rho_n * s.scale;
      if (p == 1 || p == m || col == 0 || col == ncols - 1) {   // only the rim of the unknown block can touch a Dirichlet node
        if (a - 1 >= 0 && io.isdir[node_of(s, a - 1, b)]) r -= io.dval[node_of(s, a - 1, b)];
        if (a + 1 < na && io.isdir[node_of(s, a + 1, b)]) r -= io.dval[node_of(s, a + 1, b)];
        if (b - 1 >= 0 && io.isdir[node_of(s, a, b - 1)]) r -= io.dval[node_of(s, a, b - 1)];
        if (b + 1 < nb && io.isdir[node_of(s, a, b + 1)]) r -= io.dval[node_of(s, a, b + 1)];
      }
      return r;
    }
    const int64_t e = (int64_t)col * m + (p - 1);
    double v = in[e];
    if (IN == 2 && io.f) v -= io.qt[e] * io.f[p - 1];
    return v;
  };
  // load with bit-reversed addressing; z[0] = z[m+1] = 0, z[M-p] = -z[p]
  for (int p = threadIdx.x; p <= m + 1; p += blockDim.x) {
    double2 v = make_double2(0.0, 0.0);
    if (p >= 1 && p <= m) {
      v.x = fetch(ca, p);
      v.y = has_b ? fetch(cb, p) : 0.0;
    }
    const int r0 = (int)(__brev((unsigned)p) >> (32 - log2M));
    zs[fpad(r0)] = v;
    if (p >= 1 && p <= m) {
      const int r1 = (int)(__brev((unsigned)(M - p)) >> (32 - log2M));
      zs[fpad(r1)] = make_double2(-v.x, -v.y);
    }
  }
  __syncthreads();
  // radix-2 stages fused in pairs (stage s on (a0,a1),(a2,a3), stage s+1 on (a0,a2),(a1,a3) of the same
  // four elements): half the shared-memory passes and barriers of a plain radix-2 loop
  int s = 0;
  for (; s + 1 < log2M; s += 2) {
    const int half = 1 << s;
    for (int t = threadIdx.x; t < M / 4; t += blockDim.x) {
      const int pos = t & (half - 1);
      const int i0 = ((t >> s) << (s + 2)) + pos;
      // one twiddle load per butterfly: w2a = W^(pos*2^(L-2-s)); the other two follow from it,
      // w2b = W^((pos+half)*2^(L-2-s)) = w2a * W^(M/4) = -i*w2a  and  w1 = W^(pos*2^(L-1-s)) = w2a^2
      const double2 w2a = __ldg(&tw[pos << (log2M - 2 - s)]);
      const double2 w2b = make_double2(w2a.y, -w2a.x);
      const double2 w1 = make_double2(fma(w2a.x, w2a.x, -(w2a.y * w2a.y)), 2.0 * (w2a.x * w2a.y));
      const int p0 = fpad(i0), p1 = fpad(i0 + half), p2 = fpad(i0 + 2 * half), p3 = fpad(i0 + 3 * half);
      double2 a0 = zs[p0], a1 = zs[p1], a2 = zs[p2], a3 = zs[p3];
      {
        const double br = fma(a1.x, w1.x, -(a1.y * w1.y)), bi = fma(a1.x, w1.y, a1.y * w1.x);
        a1 = make_double2(a0.x - br, a0.y - bi);
        a0 = make_double2(a0.x + br, a0.y + bi);
        const double cr = fma(a3.x, w1.x, -(a3.y * w1.y)), ci = fma(a3.x, w1.y, a3.y * w1.x);
        a3 = make_double2(a2.x - cr, a2.y - ci);
        a2 = make_double2(a2.x + cr, a2.y + ci);
      }
      {
        const double br = fma(a2.x, w2a.x, -(a2.y * w2a.y)), bi = fma(a2.x, w2a.y, a2.y * w2a.x);
        zs[p0] = make_double2(a0.x + br, a0.y + bi);
        zs[p2] = make_double2(a0.x - br, a0.y - bi);
        const double cr = fma(a3.x, w2b.x, -(a3.y * w2b.y)), ci = fma(a3.x, w2b.y, a3.y * w2b.x);
        zs[p1] = make_double2(a1.x + cr, a1.y + ci);
        zs[p3] = make_double2(a1.x - cr, a1.y - ci);
      }
    }
    __syncthreads();
  }
  for (; s < log2M; ++s) {   // odd log2M: one plain radix-2 stage left
    const int half = 1 << s;
    for (int t = threadIdx.x; t < M / 2; t += blockDim.x) {
      const int pos = t & (half - 1);
      const int i0 = ((t >> s) << (s + 1)) + pos, i1 = i0 + half;
      const double2 w = __ldg(&tw[pos << (log2M - 1 - s)]);
      const double2 a = zs[fpad(i0)], b = zs[fpad(i1)];
      const double br = fma(b.x, w.x, -(b.y * w.y)), bi = fma(b.x, w.y, b.y * w.x);
      zs[fpad(i0)] = make_double2(a.x + br, a.y + bi);
      zs[fpad(i1)] = make_double2(a.x - br, a.y - bi);
    }
    __syncthreads();
  }
  const double h = 0.5 * scale;
  for (int k = 1 + threadIdx.x; k <= m; k += blockDim.x) {
    const double2 z = zs[fpad(k)];
    if (OUT == 1) {
      io.phi[node_of(io.s, io.s.a0 + k - 1, io.s.b0 + ca)] = -h * z.y;
      if (has_b) io.phi[node_of(io.s, io.s.a0 + k - 1, io.s.b0 + cb)] = h * z.x;
    } else {
      out[(int64_t)ca * m + k - 1] = -h * z.y;
      if (has_b) out[(int64_t)cb * m + k - 1] = h * z.x;
    }
  }
}

// Thomas solves along b for every mode k on the layout the transform leaves, W[k + q*ma] (k contiguous): one thread
// per (mode, segment of q), consecutive threads = consecutive modes (coalesced), the SEG warps of a block = the SEG
// segments of 32 modes.  Both sweeps are first-order recurrences y_q = A_q*y_{q-1} + B_q: every segment first composes
// its affine map (A, B), the segment carries follow from SEG compositions, then each segment replays its part with
// the right carry -- 2 x mb/SEG dependent steps per sweep instead of mb, and no transposes around the solve (the
// warp-per-mode scan this replaces needed q-contiguous data: transpose 12 us + solve 72 us + transpose 12 us at 2049^2).
//   forward   y_q = (r_q - y_{q-1}) * m_q          A = -m_q, B = r_q*m_q      (m precomputed, [k + q*ma])
//   backward  x_q = y_q - m_q * x_{q+1}            A = -m_q, B = y_q
//   cyclic    x -= q*f, f = (x_0 + x_{mb-1}/gamma)*qden  (Sherman-Morrison): f is left in fout, the inverse transform applies it
constexpr int TSEG = 64;   // segments per system: the sweeps are latency bound, they need the loads in flight
constexpr int TKL = 16;    // systems (modes) per block: 16 lanes x 8 B = one 128-byte line per row, and 2047 modes give 128 blocks
                           // (32 modes per block were 64 blocks on 148 SMs: 70 us; this layout: see profiles/r2_launches_c5.txt)
__global__ void __launch_bounds__(TKL * TSEG) k_thomas_seg(int ma, int mb, const double *__restrict__ mt, const double *__restrict__ qden,
                                                         const double *__restrict__ gam, double *W, double *fout) {
  __shared__ double sA[TSEG][TKL], sB[TSEG][TKL];
  const int lane = threadIdx.x % TKL, seg = threadIdx.x / TKL;
  const int k = blockIdx.x * TKL + lane;
  const bool act = k < ma;
  const int per = (mb + TSEG - 1) / TSEG;
  const int q0 = seg * per, q1 = min(mb, q0 + per);
  const double *m = mt + k;
  double *w = W + k;
  // ---- forward ----
  double A = 1.0, B = 0.0;
  if (act) {
#pragma unroll 8
    for (int q = q0; q < q1; ++q) {
      const double mq = m[(int64_t)q * ma], r = w[(int64_t)q * ma];
      B = fma(-mq, B, r * mq);
      A = -mq * A;
    }
  }
  sA[seg][lane] = A; sB[seg][lane] = B;
  __syncthreads();
  double carry = 0.0;
  for (int s2 = 0; s2 < seg; ++s2) carry = fma(sA[s2][lane], carry, sB[s2][lane]);
  __syncthreads();
  // replay with the right carry; the same pass composes the segment's map of the BACKWARD sweep
  // (x_q = -m_q x_{q+1} + y_q, applied for descending q: G = F_q0 o ... o F_{q1-1}, built as G <- G o F_q while q ascends;
  // the last row has no super-diagonal: m = 0 there)
  A = 1.0; B = 0.0;
  if (act) {
    double y = carry;
#pragma unroll 8
    for (int q = q0; q < q1; ++q) {
      const double mq = m[(int64_t)q * ma];
      y = (w[(int64_t)q * ma] - y) * mq;
      w[(int64_t)q * ma] = y;
      B = fma(A, y, B);
      A = q == mb - 1 ? 0.0 : -mq * A;
    }
  }
  __syncthreads();   // (the carries above have been read by every segment)
  sA[seg][lane] = A; sB[seg][lane] = B;
  __syncthreads();
  carry = 0.0;
  for (int s2 = TSEG - 1; s2 > seg; --s2) carry = fma(sA[s2][lane], carry, sB[s2][lane]);
  __syncthreads();
  double x = carry;
  if (act) {
#pragma unroll 8
    for (int q = q1 - 1; q >= q0; --q) {
      const double mq = q == mb - 1 ? 0.0 : m[(int64_t)q * ma];
      x = fma(-mq, x, w[(int64_t)q * ma]);
      w[(int64_t)q * ma] = x;
    }
  }
  if (fout) {   // cyclic: x_0 lives in segment 0, x_{mb-1} in whichever segment holds the last row
    __syncthreads();
    if (seg == TSEG - 1) sA[0][lane] = act ? w[(int64_t)(mb - 1) * ma] : 0.0;
    __syncthreads();
    if (seg == 0 && act) fout[k] = (x + sA[0][lane] / gam[k]) * qden[k];
  }
}

// dense fallback: b = isdir ? dval : (-rho)/eps0 ; phi = Ainv * b
__global__ void k_dense_rhs(const double *__restrict__ rho, const uint8_t *__restrict__ isdir,
                            const double *__restrict__ dval, double eps0, double d2, int64_t nn, double *b) {
  for (int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; n < nn;
       n += (int64_t)gridDim.x * blockDim.x)
    b[n] = isdir[n] ? dval[n] : ((-rho[n]) / eps0) * d2;   // rows of the inverse are equilibrated by dh^2
}
// Same with Neumann rows: the sigma unknowns are identity rows (x[sigma] = b[sigma], :227-228), so they
// move to the right-hand side of the phi block:  b[r] = -rho/eps0 - A[r,sigma_k] * sigma_k.
__global__ void k_dense_rhs_sigma(const double *__restrict__ rho, const uint8_t *__restrict__ isdir,
                                  const double *__restrict__ dval, double eps0,
                                  const double *__restrict__ rowscale, const double *__restrict__ neu_coef,
                                  const int32_t *__restrict__ neu_dof, const double *__restrict__ sigma,
                                  int64_t nn, double *b) {
  for (int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; n < nn;
       n += (int64_t)gridDim.x * blockDim.x) {
    double r;
    if (isdir[n]) r = dval[n];
    else {
      r = (-rho[n]) / eps0;
      const double cf = neu_coef[n];
      if (cf != 0.0) r -= cf * sigma[neu_dof[n]];
      r *= rowscale[n];
    }
    b[n] = r;
  }
}
__global__ void k_sigma_add(double *sigma, int dof, double delta, int set) {
  if (set) sigma[dof] = delta;
  else atomicAdd(&sigma[dof], delta);   // electrode hits on the main stream add to the same word
}
__global__ void k_dense_gemv(const double *__restrict__ Ainv, const double *__restrict__ b, int64_t nn,
                             double *phi) {
  const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r >= nn) return;
  double s = 0.0;
  for (int64_t cidx = 0; cidx < nn; ++cidx) s = fma(Ainv[r + cidx * nn], b[cidx], s);
  phi[r] = s;
}

// calculate_electric_field!  generalized_poisson.jl:398-410 ; E2 = (Ex, Ey) per node
__global__ void k_efield(int nx, int ny, double dx, double dy, const double *__restrict__ phi, double2 *E2) {
  // one block row per grid row (no 64-bit div / mod per node); the quotients stay IEEE divisions like the reference's
  const double dx2 = __dmul_rn(2.0, dx), dy2 = __dmul_rn(2.0, dy);
  for (int j = blockIdx.y; j < ny; j += gridDim.y) {
    const double *row = phi + (int64_t)j * nx;
    const double *lo = j == 0 ? row : row - nx, *hi = j == ny - 1 ? row : row + nx;
    const double dyj = (j == 0 || j == ny - 1) ? dy : dy2;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nx; i += gridDim.x * blockDim.x) {
      double ex;
      if (i == 0) ex = __ddiv_rn(__dsub_rn(row[0], row[1]), dx);                          // :404
      else if (i == nx - 1) ex = __ddiv_rn(__dsub_rn(row[i - 1], row[i]), dx);            // :405
      else ex = __ddiv_rn(__dsub_rn(row[i - 1], row[i + 1]), dx2);                        // :402
      const double ey = __ddiv_rn(__dsub_rn(lo[i], hi[i]), dyj);                          // :403, :406, :407
      E2[(int64_t)j * nx + i] = make_double2(ex, ey);
    }
  }
}

// Dirichlet nodes of the separable solver sit on whole edges: phi = prescribed value there (the transform wrote the rest)
__global__ void k_phi_edges(int nx, int ny, const uint8_t *__restrict__ isdir, const double *__restrict__ dval, double *phi) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  int64_t n;
  if (t < ny) n = (int64_t)t * nx;                                   // left
  else if (t < 2 * ny) n = (nx - 1) + (int64_t)(t - ny) * nx;        // right
  else if (t < 2 * ny + nx) n = t - 2 * ny;                          // bottom
  else if (t < 2 * ny + 2 * nx) n = (t - 2 * ny - nx) + (int64_t)(ny - 1) * nx;   // top
  else return;
  if (isdir[n]) phi[n] = dval[n];
}

// per-step update of one Dirichlet edge (the RF drive, 11_rf_discharge.jl:95) without re-uploading
__global__ void k_set_edge(int nx, int ny, int edge, double v, double *dval) {
  const int len = edge < 2 ? ny : nx;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < len; t += gridDim.x * blockDim.x) {
    int64_t n;
    if (edge == ISKB_EDGE_LEFT) n = (int64_t)t * nx;
    else if (edge == ISKB_EDGE_RIGHT) n = (nx - 1) + (int64_t)t * nx;
    else if (edge == ISKB_EDGE_BOTTOM) n = t;
    else n = t + (int64_t)(ny - 1) * nx;
    dval[n] = v;
  }
}

int blocks_for(const iskb_ctx *c, int64_t n) {
  int64_t b = (n + TPB - 1) / TPB;
  if (b > (int64_t)c->n_sm * 8) b = (int64_t)c->n_sm * 8;
  return b < 1 ? 1 : (int)b;
}

template <typename T>
int32_t upload(T **dptr, const std::vector<T> &h, cudaStream_t st) {
  if (*dptr) { cudaFree(*dptr); *dptr = nullptr; }
  if (h.empty()) return ISKB_OK;
  CU_TRY(cudaMalloc(dptr, h.size() * sizeof(T)));
  CU_TRY(cudaMemcpyAsync(*dptr, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, st));
  CU_TRY(cudaStreamSynchronize(st));
  return ISKB_OK;
}

}  // namespace

int32_t poisson_free(iskb_ctx *c) {
  PoissonState &ps = c->ps;
  double **ptrs[] = {&ps.d_dval, &ps.d_V, &ps.d_lam, &ps.d_cp, &ps.d_q, &ps.d_qden, &ps.d_w1, &ps.d_w2,
                     &ps.d_Ainv, &ps.d_Vt, &ps.d_gam, &ps.d_msing, &ps.d_rowscale, &ps.d_sigma, &ps.d_neu_coef, &ps.d_cpT, &ps.d_qT,
                     &ps.d_fvec};
  if (ps.d_neu_dof) { cudaFree(ps.d_neu_dof); ps.d_neu_dof = nullptr; }
  for (auto p : ptrs) if (*p) { cudaFree(*p); *p = nullptr; }
  if (ps.d_isdir) { cudaFree(ps.d_isdir); ps.d_isdir = nullptr; }
  if (ps.d_tw) { cudaFree(ps.d_tw); ps.d_tw = nullptr; }
  return ISKB_OK;
}

static long double PIl() { return 3.141592653589793238462643383279502884L; }

static bool classify_axis(bool periodic, bool dir_lo, bool dir_hi, int n, AxisInfo &ax) {
  ax.n = n;
  if (periodic) {
    if (dir_lo || dir_hi) return false;   // mixing on one axis: not separable in closed form
    ax.kind = AX_RING; ax.lo = 0; ax.m = n;
  } else if (dir_lo && dir_hi) { ax.kind = AX_DD; ax.lo = 1; ax.m = n - 2; }
  else if (dir_lo) { ax.kind = AX_DL; ax.lo = 1; ax.m = n - 1; }
  else if (dir_hi) { ax.kind = AX_DR; ax.lo = 0; ax.m = n - 1; }
  else { ax.kind = AX_PATH; ax.lo = 0; ax.m = n; }
  return ax.m >= 1;
}

static int32_t prepare_dense(iskb_ctx *c) {
  PoissonState &ps = c->ps;
  const int64_t nn = (int64_t)c->g.nx * c->g.ny;
  if (nn > 16384) return iskb_fail(ISKB_E_UNSUPPORTED,
                                  "Dirichlet nodes are not whole edges and the grid (%lld nodes) is too large "
                                  "for the dense fallback", (long long)nn);
  std::vector<double> A, b;
  assemble_dense(c, A, b);
  const double d2 = c->g.dx * c->g.dx;
  std::vector<double> rowscale((size_t)nn, 1.0);
  if (ps.n_sigma > 0 || !ps.neu_kind.empty() || ps.has_custom) {
    // keep the phi block; the sigma columns go to the right-hand side (k_dense_rhs_sigma)
    const int64_t nt = nn + ps.n_sigma;
    std::vector<double> P((size_t)(nn * nn)), coef((size_t)nn, 0.0);
    std::vector<int32_t> dof((size_t)nn, 0);
    for (int64_t col = 0; col < nn; ++col)
      for (int64_t r = 0; r < nn; ++r) P[(size_t)(r + col * nn)] = A[(size_t)(r + col * nt)];
    for (int64_t r = 0; r < nn; ++r)
      if (!ps.neu_kind.empty() && ps.neu_kind[(size_t)r] && !ps.isdir[(size_t)r]) {
        dof[(size_t)r] = ps.neu_dof[(size_t)r];
        coef[(size_t)r] = A[(size_t)(r + (nn + dof[(size_t)r]) * nt)];
      }
    A.swap(P);
    ISKB_TRY(upload(&ps.d_neu_coef, coef, c->stream));
    if (ps.d_neu_dof) { cudaFree(ps.d_neu_dof); ps.d_neu_dof = nullptr; }
    CU_TRY(cudaMalloc(&ps.d_neu_dof, nn * sizeof(int32_t)));
    CU_TRY(cudaMemcpyAsync(ps.d_neu_dof, dof.data(), nn * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
    CU_TRY(cudaStreamSynchronize(c->stream));
  }
  // row equilibration: stencil rows carry 1/dh^2, Dirichlet rows 1 (generalized_poisson.jl:65,210-211),
  // Neumann rows 2/dx or 4/(3 dh^2) (:252-253,:262-265); inverting D*A (and scaling the rhs in
  // k_dense_rhs) keeps the inverse accurate to ~cond*eps
  for (int64_t r = 0; r < nn; ++r) {
    if (ps.isdir[(size_t)r]) continue;
    double sc = d2;
    if (ps.has_custom || (!ps.neu_kind.empty() && ps.neu_kind[(size_t)r])) {
      double mx = 0.0;
      for (int64_t col = 0; col < nn; ++col) mx = std::fmax(mx, std::fabs(A[(size_t)(r + col * nn)]));
      sc = mx > 0.0 ? 1.0 / mx : 1.0;
    }
    rowscale[(size_t)r] = sc;
    for (int64_t col = 0; col < nn; ++col) A[(size_t)(r + col * nn)] *= sc;
  }
  if (nn >= 512) {   // same pivots, same arithmetic, on the device (below that the launches cost more than the host loop)
    if (ps.d_Ainv) { cudaFree(ps.d_Ainv); ps.d_Ainv = nullptr; }
    ISKB_TRY(invert_dense_device(c, A, nn, &ps.d_Ainv));
  } else {
    if (!invert_dense(A, nn)) return iskb_fail(ISKB_E_SINGULAR, "dense Poisson operator is singular");
    ISKB_TRY(upload(&ps.d_Ainv, A, c->stream));
  }
  ISKB_TRY(upload(&ps.d_rowscale, rowscale, c->stream));
  ps.nn_dense = nn;
  if (!ps.d_w1) CU_TRY(cudaMalloc(&ps.d_w1, nn * sizeof(double)));
  ps.mode = 2;
  return ISKB_OK;
}

int32_t poisson_prepare(iskb_ctx *c) {
  PoissonState &ps = c->ps;
  if (!ps.created) return iskb_fail(ISKB_E_INVALID, "iskb_poisson_create must be called first");
  const int nx = c->g.nx, ny = c->g.ny;
  const int64_t nn = (int64_t)nx * ny;
  if (ps.structure_dirty || ps.values_dirty) CU_TRY(cudaStreamSynchronize(c->fstream));   // a solve in flight reads these buffers
  if (ps.structure_dirty) {
    {
      std::vector<uint8_t> &m = ps.isdir;
      if (ps.d_isdir) { cudaFree(ps.d_isdir); ps.d_isdir = nullptr; }
      CU_TRY(cudaMalloc(&ps.d_isdir, nn));
      CU_TRY(cudaMemcpyAsync(ps.d_isdir, m.data(), nn, cudaMemcpyHostToDevice, c->stream));
    }
    // which whole edges are Dirichlet, and is every Dirichlet node on such an edge?
    bool eL = true, eR = true, eB = true, eT = true;
    for (int j = 0; j < ny; ++j) { eL &= ps.isdir[(size_t)(0 + (int64_t)j * nx)] != 0; eR &= ps.isdir[(size_t)((nx - 1) + (int64_t)j * nx)] != 0; }
    for (int i = 0; i < nx; ++i) { eB &= ps.isdir[(size_t)i] != 0; eT &= ps.isdir[(size_t)(i + (int64_t)(ny - 1) * nx)] != 0; }
    bool whole = true;
    for (int j = 0; j < ny && whole; ++j)
      for (int i = 0; i < nx; ++i)
        if (ps.isdir[(size_t)(i + (int64_t)j * nx)]) {
          const bool cov = (eL && i == 0) || (eR && i == nx - 1) || (eB && j == 0) || (eT && j == ny - 1);
          if (!cov) { whole = false; break; }
        }
    AxisInfo axi{}, axj{};
    bool has_neumann = ps.n_sigma > 0;
    for (uint8_t k : ps.neu_kind) has_neumann |= k != 0;
    bool sep = whole && c->g.dx == c->g.dy && !has_neumann && !ps.has_custom;
    sep = sep && classify_axis(ps.periodic_i, eL, eR, nx, axi) && classify_axis(ps.periodic_j, eB, eT, ny, axj);
    // choose the transform axis: the solve axis needs >= 3 unknowns when cyclic, >= 1 otherwise
    auto ok_solve = [](const AxisInfo &b) { return b.kind == AX_RING ? b.m >= 3 : b.m >= 1; };
    int choice = -1;   // 0: transform along i, 1: transform along j
    if (sep) {
      const bool can0 = ok_solve(axj), can1 = ok_solve(axi);
      if (can0 && can1) choice = axi.m <= axj.m ? 0 : 1;   // transform along the shorter axis
      else if (can0) choice = 0;
      else if (can1) choice = 1;
    }
    if (choice < 0) {
      ISKB_TRY(prepare_dense(c));
    } else {
      const AxisInfo &A = choice == 0 ? axi : axj, &B = choice == 0 ? axj : axi;
      ps.transposed = choice == 1;
      ps.na = A.n; ps.nb = B.n; ps.a0 = A.lo; ps.ma = A.m; ps.b0 = B.lo; ps.mb = B.m;
      ps.b_cyclic = B.kind == AX_RING;
      const bool a_null = A.kind == AX_RING || A.kind == AX_PATH;
      const bool b_null = B.kind == AX_RING || B.kind == AX_PATH;
      ps.singular = a_null && b_null;
      std::vector<double> V, lam, db;
      axis_eigen(A, V, lam);
      axis_diag(B, db);
      std::vector<double> Vt((size_t)A.m * A.m);
      for (int p = 0; p < A.m; ++p)
        for (int k = 0; k < A.m; ++k) Vt[(size_t)k + (size_t)p * A.m] = V[(size_t)p + (size_t)k * A.m];
      ps.singular_mode = -1;
      if (ps.singular)
        for (int k = 0; k < A.m; ++k) if (std::fabs(lam[(size_t)k]) < 1e-14) ps.singular_mode = k;
      // Thomas factors per mode (and Sherman-Morrison vectors when cyclic)
      const int ma = A.m, mb = B.m;
      std::vector<double> mt((size_t)ma * mb), qt, qden, gam, msing;
      if (ps.b_cyclic) { qt.assign((size_t)ma * mb, 0.0); qden.assign((size_t)ma, 0.0); gam.assign((size_t)ma, 1.0); }
      std::vector<double> d((size_t)mb), qv((size_t)mb);
      for (int k = 0; k < ma; ++k) {
        if (k == ps.singular_mode) {
          for (int q = 0; q < mb; ++q) mt[(size_t)q + (size_t)k * mb] = 0.0;
          continue;
        }
        for (int q = 0; q < mb; ++q) d[(size_t)q] = db[(size_t)q] + lam[(size_t)k];
        double g = 1.0;
        if (ps.b_cyclic) {
          g = -d[0];
          gam[(size_t)k] = g;
          d[0] -= g;
          d[(size_t)mb - 1] -= 1.0 / g;
        }
        double cprev = 0.0;
        for (int q = 0; q < mb; ++q) {
          const double mq = 1.0 / (d[(size_t)q] - cprev);
          mt[(size_t)q + (size_t)k * mb] = mq;
          cprev = mq;
        }
        if (ps.b_cyclic) {
          // solve A' q = u, u = (g, 0, ..., 0, 1)
          double y = 0.0;
          for (int q = 0; q < mb; ++q) {
            const double u = q == 0 ? g : (q == mb - 1 ? 1.0 : 0.0);
            y = (u - y) * mt[(size_t)q + (size_t)k * mb];
            qv[(size_t)q] = y;
          }
          double x = 0.0;
          for (int q = mb - 1; q >= 0; --q) {
            x = qv[(size_t)q] - (q == mb - 1 ? 0.0 : mt[(size_t)q + (size_t)k * mb] * x);
            qv[(size_t)q] = x;
          }
          for (int q = 0; q < mb; ++q) qt[(size_t)q + (size_t)k * mb] = qv[(size_t)q];
          qden[(size_t)k] = 1.0 / (1.0 + qv[0] + qv[(size_t)mb - 1] / g);
        }
      }
      if (ps.singular_mode >= 0) {
        // pinned system on q = 1..mb-1: diag = db[q] (lambda = 0), no cyclic corner
        msing.assign((size_t)mb, 0.0);
        double cprev = 0.0;
        for (int q = 1; q < mb; ++q) {
          const double mq = 1.0 / (db[(size_t)q] - cprev);
          msing[(size_t)q] = mq;
          cprev = mq;
        }
      }
      ISKB_TRY(upload(&ps.d_V, V, c->stream));
      ISKB_TRY(upload(&ps.d_Vt, Vt, c->stream));
      ISKB_TRY(upload(&ps.d_lam, lam, c->stream));
      ISKB_TRY(upload(&ps.d_cp, mt, c->stream));
      ISKB_TRY(upload(&ps.d_q, qt, c->stream));
      ISKB_TRY(upload(&ps.d_qden, qden, c->stream));
      ISKB_TRY(upload(&ps.d_gam, gam, c->stream));
      ISKB_TRY(upload(&ps.d_msing, msing, c->stream));
      // DST-I by FFT when the Dirichlet-Dirichlet interior length + 1 is a power of two
      ps.use_fft = false;
      ps.fft_log2M = 0;
      if (A.kind == AX_DD && ma + 1 >= 16 && ((ma + 1) & ma) == 0 && 2 * (ma + 1) <= 8192) {
        int lg = 0;
        while ((1 << lg) < 2 * (ma + 1)) ++lg;
        const int M = 1 << lg;
        std::vector<double2> tw((size_t)M / 2);
        for (int j = 0; j < M / 2; ++j) {
          const long double th = -2.0L * PIl() * j / M;
          tw[(size_t)j] = make_double2((double)cosl(th), (double)sinl(th));
        }
        if (ps.d_tw) { cudaFree(ps.d_tw); ps.d_tw = nullptr; }
        CU_TRY(cudaMalloc(&ps.d_tw, tw.size() * sizeof(double2)));
        CU_TRY(cudaMemcpyAsync(ps.d_tw, tw.data(), tw.size() * sizeof(double2), cudaMemcpyHostToDevice, c->stream));
        CU_TRY(cudaStreamSynchronize(c->stream));
        CU_TRY(cudaFuncSetAttribute(k_dst_fft<1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, fpad_host(M) * (int)sizeof(double2)));
        CU_TRY(cudaFuncSetAttribute(k_dst_fft<2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, fpad_host(M) * (int)sizeof(double2)));
        ps.use_fft = true;
        ps.fft_log2M = lg;
        // Thomas factors in the layout the transform leaves ([k + q*ma], see k_thomas_seg)
        std::vector<double> mtT((size_t)ma * mb), qtT;
        for (int k = 0; k < ma; ++k)
          for (int q = 0; q < mb; ++q) mtT[(size_t)k + (size_t)q * ma] = mt[(size_t)q + (size_t)k * mb];
        if (ps.b_cyclic) {
          qtT.assign((size_t)ma * mb, 0.0);
          for (int k = 0; k < ma; ++k)
            for (int q = 0; q < mb; ++q) qtT[(size_t)k + (size_t)q * ma] = qt[(size_t)q + (size_t)k * mb];
        }
        ISKB_TRY(upload(&ps.d_cpT, mtT, c->stream));
        ISKB_TRY(upload(&ps.d_qT, qtT, c->stream));
        if (ps.d_fvec) { cudaFree(ps.d_fvec); ps.d_fvec = nullptr; }
        CU_TRY(cudaMalloc(&ps.d_fvec, (size_t)ma * sizeof(double)));
        CU_TRY(cudaStreamSynchronize(c->stream));
      }
      if (ps.d_w1) { cudaFree(ps.d_w1); ps.d_w1 = nullptr; }
      if (ps.d_w2) { cudaFree(ps.d_w2); ps.d_w2 = nullptr; }
      CU_TRY(cudaMalloc(&ps.d_w1, (size_t)ma * mb * sizeof(double)));
      CU_TRY(cudaMalloc(&ps.d_w2, (size_t)ma * mb * sizeof(double)));
      ps.mode = 1;
    }
    ps.structure_dirty = false;
    ps.values_dirty = true;
  }
  if (ps.sigma_host_newer && ps.n_sigma > 0) {
    if (!ps.d_sigma) CU_TRY(cudaMalloc(&ps.d_sigma, ISKB_MAX_SIGMA * sizeof(double)));
    CU_TRY(cudaStreamSynchronize(c->fstream));
    CU_TRY(cudaMemcpyAsync(ps.d_sigma, ps.sigma.data(), ps.n_sigma * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CU_TRY(cudaStreamSynchronize(c->stream));
    ps.sigma_host_newer = false;
  }
  if (ps.values_dirty) {
    if (!ps.d_dval) CU_TRY(cudaMalloc(&ps.d_dval, nn * sizeof(double)));
    CU_TRY(cudaMemcpyAsync(ps.d_dval, ps.dval.data(), nn * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CU_TRY(cudaStreamSynchronize(c->stream));   // host vector may change right after
    ps.values_dirty = false;
  }
  return ISKB_OK;
}

// Asynchronous: the solve is ordered after everything issued so far on the main stream (rho) and runs
// on the field stream; main-stream users of phi / E call fields_join() first.
int32_t poisson_solve(iskb_ctx *c) {
  ISKB_TRY(poisson_prepare(c));
  // only the FFT path forms rho from the fixed-point sums on its way in
  if (c->rho_lazy && !(c->ps.mode != 2 && c->ps.use_fft)) ISKB_TRY(rho_materialize(c));
  CU_TRY(cudaEventRecord(c->ev_rho, c->stream));
  CU_TRY(cudaStreamWaitEvent(c->fstream, c->ev_rho, 0));
  PoissonState &ps = c->ps;
  const int nx = c->g.nx, ny = c->g.ny;
  const int64_t nn = (int64_t)nx * ny;
  if (ps.mode == 2) {
    if (ps.d_neu_coef && (ps.n_sigma > 0 || ps.has_custom))
      k_dense_rhs_sigma<<<blocks_for(c, nn), TPB, 0, c->fstream>>>(c->d_rho, ps.d_isdir, ps.d_dval, ps.eps0, ps.d_rowscale,
                                                                  ps.d_neu_coef, ps.d_neu_dof, ps.d_sigma, nn, ps.d_w1);
    else
      k_dense_rhs<<<blocks_for(c, nn), TPB, 0, c->fstream>>>(c->d_rho, ps.d_isdir, ps.d_dval, ps.eps0, c->g.dx * c->g.dx, nn,
                                                            ps.d_w1);
    LAUNCH_CHECK(c);
    k_dense_gemv<<<(int)((nn + 127) / 128), 128, 0, c->fstream>>>(ps.d_Ainv, ps.d_w1, nn, c->d_phi);
    LAUNCH_CHECK(c);
  } else {
    SolveDims s{nx, ny, ps.transposed ? 1 : 0, ps.a0, ps.ma, ps.b0, ps.mb, c->g.dx * c->g.dx / ps.eps0};
    const int64_t tot = (int64_t)ps.ma * ps.mb;
    const double dst_scale = sqrt(2.0 / (ps.ma + 1));
    const int M = 1 << ps.fft_log2M;
    if (ps.use_fft) {
      // rho --DST (rhs formed on the way in)--> w2[k + q*ma] --Thomas per mode, in place--> w2 --DST--> phi
      FftIo io{s, c->d_rho, ps.d_isdir, ps.d_dval, ps.d_qT, ps.b_cyclic ? ps.d_fvec : nullptr, c->d_phi};
      io.ns = 0;
      if (c->rho_lazy) {   // (every other path asked for rho_materialize above)
        io.ns = c->rho_ns;
        for (int q = 0; q < c->rho_ns; ++q) { io.u[q] = c->rho_u[q]; io.z[q] = c->rho_z[q]; }
        io.V = c->d_V;
        io.q0s = c->q0 / c->fscale;
      }
      k_dst_fft<1, 0><<<(ps.mb + 1) / 2, 512, fpad_host(M) * sizeof(double2), c->fstream>>>(ps.ma, ps.mb, ps.fft_log2M, ps.d_tw, nullptr,
                                                                                ps.d_w2, dst_scale, io);
      LAUNCH_CHECK(c);
      k_thomas_seg<<<(ps.ma + TKL - 1) / TKL, TKL * TSEG, 0, c->fstream>>>(ps.ma, ps.mb, ps.d_cpT, ps.d_qden, ps.d_gam, ps.d_w2,
                                                                   ps.b_cyclic ? ps.d_fvec : nullptr);
      LAUNCH_CHECK(c);
      k_dst_fft<2, 1><<<(ps.mb + 1) / 2, 512, fpad_host(M) * sizeof(double2), c->fstream>>>(ps.ma, ps.mb, ps.fft_log2M, ps.d_tw, ps.d_w2,
                                                                                nullptr, dst_scale, io);
      LAUNCH_CHECK(c);
      k_phi_edges<<<(2 * (nx + ny) + 255) / 256, 256, 0, c->fstream>>>(nx, ny, ps.d_isdir, ps.d_dval, c->d_phi);
      LAUNCH_CHECK(c);
    } else {
      k_build_rhs<<<blocks_for(c, tot), TPB, 0, c->fstream>>>(s, c->d_rho, ps.d_isdir, ps.d_dval, ps.d_w1);
      LAUNCH_CHECK(c);
      dim3 gg((ps.ma + 63) / 64, (ps.mb + 63) / 64);
      // forward transform  w1 -> w2
      k_dgemm<<<gg, 256, 0, c->fstream>>>(ps.ma, ps.mb, ps.ma, ps.d_Vt, ps.ma, ps.d_w1, ps.ma, ps.d_w2, ps.ma);
      LAUNCH_CHECK(c);
      // tridiagonal solves along b on mode-major data:  w2 --T--> w1, in place, w1 --T--> w2
      {
        dim3 tg((ps.ma + 31) / 32, (ps.mb + 31) / 32), tb(32, 8);
        k_transpose<<<tg, tb, 0, c->fstream>>>(ps.ma, ps.mb, ps.d_w2, ps.d_w1);
        LAUNCH_CHECK(c);
        k_thomas_warp<<<(ps.ma + 3) / 4, 128, 0, c->fstream>>>(ps.ma, ps.mb, ps.b_cyclic ? 1 : 0, ps.singular_mode, ps.d_cp,
                                                              ps.d_msing, ps.d_q, ps.d_qden, ps.d_gam, ps.d_w1);
        LAUNCH_CHECK(c);
        dim3 tg2((ps.mb + 31) / 32, (ps.ma + 31) / 32);
        k_transpose<<<tg2, tb, 0, c->fstream>>>(ps.mb, ps.ma, ps.d_w1, ps.d_w2);
        LAUNCH_CHECK(c);
      }
      // inverse transform  w2 -> w1
      k_dgemm<<<gg, 256, 0, c->fstream>>>(ps.ma, ps.mb, ps.ma, ps.d_V, ps.ma, ps.d_w2, ps.ma, ps.d_w1, ps.ma);
      LAUNCH_CHECK(c);
      ps.d_w3 = ps.d_w1;   // (alias, not owned) result of the inverse transform
      k_store_phi<<<blocks_for(c, nn), TPB, 0, c->fstream>>>(s, ps.d_w3, ps.d_isdir, ps.d_dval, c->d_phi);
      LAUNCH_CHECK(c);
    }
  }
  {
    const dim3 eg((unsigned)std::min((nx + TPB - 1) / TPB, 8), (unsigned)std::min(ny, 65535));
    k_efield<<<eg, TPB, 0, c->fstream>>>(nx, ny, c->g.dx, c->g.dy, c->d_phi, c->d_E2);
  }
  LAUNCH_CHECK(c);
  CU_TRY(cudaEventRecord(c->ev_E, c->fstream));
  c->fields_pending = true;
  return ISKB_OK;
}

// ---- C ABI ------------------------------------------------------------------------------------
// test hook: the host-side dense inversion on its own (no GPU needed); A is n*n column-major, overwritten by its inverse
extern "C" int32_t iskb_debug_invert_dense(double *A, int64_t n) {
  std::vector<double> M(A, A + n * n);
  if (!invert_dense(M, n)) return ISKB_E_SINGULAR;
  memcpy(A, M.data(), (size_t)(n * n) * sizeof(double));
  return ISKB_OK;
}

// test hook: the device-side inversion (bit-identical to the host one); A is n*n column-major, overwritten by its inverse
extern "C" int32_t iskb_debug_invert_dense_device(iskb_ctx *c, double *A, int64_t n) {
  if (!c || !A || n < 1) return iskb_fail(ISKB_E_INVALID, "bad arguments");
  CU_TRY(cudaSetDevice(c->device));
  std::vector<double> M(A, A + n * n);
  double *d = nullptr;
  ISKB_TRY(invert_dense_device(c, M, n, &d));
  CU_TRY(cudaMemcpy(A, d, (size_t)(n * n) * sizeof(double), cudaMemcpyDeviceToHost));
  cudaFree(d);
  return ISKB_OK;
}

static int32_t sigma_pull(iskb_ctx *c);
extern "C" int32_t iskb_poisson_create(iskb_ctx *c, double eps0) {
  if (!c || !c->has_grid) return iskb_fail(ISKB_E_INVALID, "iskb_grid_set must be called first");
  PoissonState &ps = c->ps;
  const int64_t nn = (int64_t)c->g.nx * c->g.ny;
  ps.created = true;
  ps.eps0 = eps0;
  ps.periodic_i = ps.periodic_j = false;
  ps.isdir.assign((size_t)nn, 0);
  ps.dval.assign((size_t)nn, 0.0);
  ps.n_sigma = 0;
  ps.sigma.clear();
  ps.neu_kind.clear(); ps.neu_i2.clear(); ps.neu_j2.clear(); ps.neu_dof.clear();
  ps.has_custom = false;
  ps.custom_A.clear();
  ps.structure_dirty = ps.values_dirty = true;
  return ISKB_OK;
}

extern "C" int32_t iskb_poisson_apply_periodic(iskb_ctx *c, int32_t axis) {
  if (!c || !c->ps.created) return iskb_fail(ISKB_E_INVALID, "no Poisson solver");
  if (axis == 1) c->ps.periodic_j = true;        // :291-306 couples j = 1 <-> ny
  else if (axis == 2) c->ps.periodic_i = true;   // :308-323 couples i = 1 <-> nx
  else return iskb_fail(ISKB_E_INVALID, "axis must be 1 or 2");
  c->ps.structure_dirty = true;
  return ISKB_OK;
}

extern "C" int32_t iskb_poisson_apply_dirichlet(iskb_ctx *c, const uint8_t *mask, double phi0) {
  if (!c || !c->ps.created || !mask) return iskb_fail(ISKB_E_INVALID, "no Poisson solver / mask");
  PoissonState &ps = c->ps;
  const int64_t nn = (int64_t)c->g.nx * c->g.ny;
  for (int64_t n = 0; n < nn; ++n)
    if (mask[n]) {
      if (!ps.neu_kind.empty() && ps.neu_kind[(size_t)n])
        return iskb_fail(ISKB_E_UNSUPPORTED, "node %lld carries a Neumann row; overlapping electrodes are not supported", (long long)n);
      if (!ps.isdir[(size_t)n]) { ps.isdir[(size_t)n] = 1; ps.structure_dirty = true; }
      ps.dval[(size_t)n] = phi0;
      ps.values_dirty = true;
    }
  return ISKB_OK;
}

extern "C" int32_t iskb_poisson_apply_dirichlet_edge(iskb_ctx *c, int32_t edge, double phi0) {
  if (!c || !c->ps.created) return iskb_fail(ISKB_E_INVALID, "no Poisson solver");
  PoissonState &ps = c->ps;
  const int nx = c->g.nx, ny = c->g.ny;
  auto setn = [&](int64_t n) {
    if (!ps.isdir[(size_t)n]) { ps.isdir[(size_t)n] = 1; ps.structure_dirty = true; }
    ps.dval[(size_t)n] = phi0;
  };
  switch (edge) {
    case ISKB_EDGE_LEFT: for (int j = 0; j < ny; ++j) setn((int64_t)j * nx); break;
    case ISKB_EDGE_RIGHT: for (int j = 0; j < ny; ++j) setn((nx - 1) + (int64_t)j * nx); break;
    case ISKB_EDGE_BOTTOM: for (int i = 0; i < nx; ++i) setn(i); break;
    case ISKB_EDGE_TOP: for (int i = 0; i < nx; ++i) setn(i + (int64_t)(ny - 1) * nx); break;
    default: return iskb_fail(ISKB_E_INVALID, "bad edge");
  }
  if (!ps.structure_dirty && !ps.values_dirty && ps.d_dval) {
    // solver already built and only this edge's value changed: patch the device copy in place
    const int len = edge < 2 ? ny : nx;
    // on the field stream: ordered after a solve still in flight (it reads d_dval) and before the next one
    k_set_edge<<<(len + 255) / 256, 256, 0, c->fstream>>>(nx, ny, edge, phi0, ps.d_dval);
    LAUNCH_CHECK(c);
    return ISKB_OK;
  }
  ps.values_dirty = true;
  return ISKB_OK;
}

extern "C" int32_t iskb_poisson_get_dense(iskb_ctx *c, double *A_out, double *b_out) {
  if (!c || !c->ps.created) return iskb_fail(ISKB_E_INVALID, "no Poisson solver");
  ISKB_TRY(sigma_pull(c));
  std::vector<double> A, b;
  assemble_dense(c, A, b);
  if (A_out) memcpy(A_out, A.data(), A.size() * sizeof(double));
  if (b_out) memcpy(b_out, b.data(), b.size() * sizeof(double));
  return ISKB_OK;
}

// ---- sigma dofs / Neumann rows / electrodes -----------------------------------
// sigma lives on the device once the solver is built (electrode hits may add to it); before that, and
// after host-side edits, the host vector is authoritative.
static int32_t sigma_pull(iskb_ctx *c) {
  PoissonState &ps = c->ps;
  if (ps.sigma_host_newer || !ps.d_sigma || ps.n_sigma == 0) return ISKB_OK;
  ISKB_TRY(fields_join(c));
  CU_TRY(cudaStreamSynchronize(c->fstream));
  CU_TRY(cudaMemcpyAsync(ps.sigma.data(), ps.d_sigma, ps.n_sigma * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CU_TRY(cudaStreamSynchronize(c->stream));
  return ISKB_OK;
}

int32_t poisson_sigma_device(iskb_ctx *c, double **out) {
  ISKB_TRY(poisson_prepare(c));
  *out = c->ps.d_sigma;
  return ISKB_OK;
}

extern "C" int32_t iskb_poisson_add_dof(iskb_ctx *c, int32_t *dof_out) {
  if (!c || !c->ps.created) return iskb_fail(ISKB_E_INVALID, "no Poisson solver");
  PoissonState &ps = c->ps;
  if (ps.n_sigma >= ISKB_MAX_SIGMA) return iskb_fail(ISKB_E_UNSUPPORTED, "more than %d sigma dofs", ISKB_MAX_SIGMA);
  ISKB_TRY(sigma_pull(c));
  ps.sigma.push_back(0.0);                       // :228  b[s] = 0
  ps.n_sigma++;
  ps.sigma_host_newer = true;
  ps.structure_dirty = true;
  if (dof_out) *dof_out = ps.n_sigma;            // :229 length(s)
  return ISKB_OK;
}

extern "C" int32_t iskb_poisson_apply_neumann(iskb_ctx *c, const uint8_t *mask, int32_t dof) {
  if (!c || !c->ps.created || !mask) return iskb_fail(ISKB_E_INVALID, "no Poisson solver / mask");
  PoissonState &ps = c->ps;
  if (dof < 1 || dof > ps.n_sigma) return iskb_fail(ISKB_E_INVALID, "sigma dof %d does not exist", dof);
  const int nx = c->g.nx, ny = c->g.ny;
  const int64_t nn = (int64_t)nx * ny;
  if (ps.neu_kind.empty()) {
    ps.neu_kind.assign((size_t)nn, 0);
    ps.neu_i2.assign((size_t)nn, 0);
    ps.neu_j2.assign((size_t)nn, 0);
    ps.neu_dof.assign((size_t)nn, 0);
  }
  auto nd = [&](int i, int j) { return mask[(size_t)(i + (int64_t)j * nx)] != 0; };
  for (int j = 0; j < ny; ++j)
    for (int i = 0; i < nx; ++i) {
      if (!nd(i, j)) continue;
      const int64_t r = i + (int64_t)j * nx;
      const bool a = i > 0 ? nd(i - 1, j) : false;          // :241-244
      const bool b = i < nx - 1 ? nd(i + 1, j) : false;
      const bool cc = j > 0 ? nd(i, j - 1) : true;
      const bool d = j < ny - 1 ? nd(i, j + 1) : true;
      int kind = 0;
      if (cc && d && !a && !b) kind = 1;                    // :246-247
      if (cc != d) kind = 2;                                // :256
      if (!kind) continue;
      if (ps.isdir[(size_t)r])
        return iskb_fail(ISKB_E_UNSUPPORTED, "node (%d,%d) is already a Dirichlet node; overlapping electrodes are not supported",
                         i + 1, j + 1);
      ps.neu_kind[(size_t)r] = (uint8_t)kind;
      ps.neu_i2[(size_t)r] = i == 0 ? i + 1 : i - 1;        // :249,:257
      ps.neu_j2[(size_t)r] = cc ? j + 1 : j - 1;            // :258
      ps.neu_dof[(size_t)r] = dof - 1;
      if (kind == 2 && (ps.neu_j2[(size_t)r] < 0 || ps.neu_j2[(size_t)r] >= ny))
        return iskb_fail(ISKB_E_INVALID, "Neumann strip end at (%d,%d) points outside the grid (reference: BoundsError)", i + 1, j + 1);
    }
  ps.structure_dirty = true;
  return ISKB_OK;
}

static int32_t sigma_edit(iskb_ctx *c, int32_t dof, double v, int set) {
  if (!c || !c->ps.created) return iskb_fail(ISKB_E_INVALID, "no Poisson solver");
  PoissonState &ps = c->ps;
  if (dof < 1 || dof > ps.n_sigma) return iskb_fail(ISKB_E_INVALID, "sigma dof %d does not exist", dof);
  if (ps.sigma_host_newer || !ps.d_sigma) {
    ps.sigma[(size_t)(dof - 1)] = set ? v : ps.sigma[(size_t)(dof - 1)] + v;
    ps.sigma_host_newer = true;
    return ISKB_OK;
  }
  // device copy is authoritative: edit it in stream order after a solve still in flight
  k_sigma_add<<<1, 1, 0, c->fstream>>>(ps.d_sigma, dof - 1, v, set);
  LAUNCH_CHECK(c);
  return ISKB_OK;
}
extern "C" int32_t iskb_poisson_sigma_set(iskb_ctx *c, int32_t dof, double value) { return sigma_edit(c, dof, value, 1); }
extern "C" int32_t iskb_poisson_sigma_add(iskb_ctx *c, int32_t dof, double delta) { return sigma_edit(c, dof, delta, 0); }
extern "C" int32_t iskb_poisson_sigma_get(iskb_ctx *c, int32_t dof, double *out) {
  if (!c || !c->ps.created || !out) return iskb_fail(ISKB_E_INVALID, "no Poisson solver");
  if (dof < 1 || dof > c->ps.n_sigma) return iskb_fail(ISKB_E_INVALID, "sigma dof %d does not exist", dof);
  ISKB_TRY(sigma_pull(c));
  *out = c->ps.sigma[(size_t)(dof - 1)];
  return ISKB_OK;
}

extern "C" int32_t iskb_poisson_dense_size(iskb_ctx *c, int64_t *n_out) {
  if (!c || !c->ps.created || !n_out) return iskb_fail(ISKB_E_INVALID, "no Poisson solver");
  *n_out = (int64_t)c->g.nx * c->g.ny + c->ps.n_sigma;
  return ISKB_OK;
}

extern "C" int32_t iskb_phi_at(iskb_ctx *c, int32_t i, int32_t j, double *out) {
  if (!c || !c->has_grid || !out) return iskb_fail(ISKB_E_INVALID, "no grid");
  if (i < 1 || i > c->g.nx || j < 1 || j > c->g.ny) return iskb_fail(ISKB_E_INVALID, "node (%d,%d) outside the grid", i, j);
  ISKB_TRY(fields_join(c));
  CU_TRY(cudaMemcpyAsync(c->h_scratch, c->d_phi + (i - 1) + (int64_t)(j - 1) * c->g.nx, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CU_TRY(cudaStreamSynchronize(c->stream));
  memcpy(out, c->h_scratch, sizeof(double));
  return ISKB_OK;
}

extern "C" int32_t iskb_poisson_set_dense(iskb_ctx *c, const double *A, int64_t nn) {
  if (!c || !c->ps.created || !A) return iskb_fail(ISKB_E_INVALID, "no Poisson solver / operator");
  if (nn != (int64_t)c->g.nx * c->g.ny) return iskb_fail(ISKB_E_INVALID, "operator must be (nx*ny)^2");
  if (nn > 16384) return iskb_fail(ISKB_E_UNSUPPORTED, "caller-assembled operators use the dense path (at most 16384 nodes)");
  c->ps.custom_A.assign(A, A + nn * nn);
  c->ps.has_custom = true;
  c->ps.structure_dirty = true;
  return ISKB_OK;
}

extern "C" int32_t iskb_poisson_mode(iskb_ctx *c, int32_t *mode_out) {
  if (!c || !mode_out) return iskb_fail(ISKB_E_INVALID, "null");
  ISKB_TRY(poisson_prepare(c));
  *mode_out = c->ps.mode;
  return ISKB_OK;
}

extern "C" int32_t iskb_field_solve(iskb_ctx *c) {
  if (!c || !c->has_grid) return iskb_fail(ISKB_E_INVALID, "no grid");
  ISKB_TRY(poisson_solve(c));
  return fields_join(c);
}
