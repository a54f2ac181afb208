// K2 (FP32 SIMT variant) and K4: dense kernels on the z-scaled HVG matrix Z (row-major float32).
// The SIMT Gram is the mode-0 path of scf_gram_accumulate: it is the accuracy anchor for the
// tcgen05 TF32 / 3xTF32 kernels in gram_tc.cu and the fallback when n_cols is tiny.
#include "common.cuh"

namespace {

// ---------------------------------------------------------------------------------------------
// Gram: one CTA = one 128x128 output tile (upper triangle incl. diagonal) x one slab of rows.
// As[kk][m] = Z[row0+kk][mi*128+m], Bs[kk][n] = Z[row0+kk][ni*128+n]; 256 threads, 8x8 micro tile.
constexpr int GT = 128, GK = 16;

__global__ void __launch_bounds__(256) gram_simt_kernel(const float* __restrict__ z, int64_t ldz, int64_t n_rows,
                                                        int n_cols, long long* __restrict__ g_fx, int64_t ldg,
                                                        int n_tiles) {
  __shared__ __align__(16) float As[GK][GT];
  __shared__ __align__(16) float Bs[GK][GT];
  // decode upper-triangular tile pair
  int t = blockIdx.x, mi = 0;
  while (t >= n_tiles - mi) {
    t -= n_tiles - mi;
    ++mi;
  }
  const int ni = mi + t;
  const int64_t row0 = (int64_t)blockIdx.y * SCF_GRAM_SLAB;
  const int64_t row1 = min(row0 + (int64_t)SCF_GRAM_SLAB, n_rows);
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  // loader mapping: 16 rows x 32 float4 per operand = 512 float4 -> 2 per thread
  const int lr = tid >> 5, lc = (tid & 31) * 4;
  const bool a_ok = mi * GT + lc < ldz, b_ok = ni * GT + lc < ldz;
  for (int64_t k0 = row0; k0 < row1; k0 += GK) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int rr = lr + 8 * h;
      const int64_t row = k0 + rr;
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
      if (row < row1) {
        if (a_ok) a = __ldg(reinterpret_cast<const float4*>(z + row * ldz + mi * GT + lc));
        if (b_ok) b = __ldg(reinterpret_cast<const float4*>(z + row * ldz + ni * GT + lc));
      }
      *reinterpret_cast<float4*>(&As[rr][lc]) = a;
      *reinterpret_cast<float4*>(&Bs[rr][lc]) = b;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < GK; ++kk) {
      float a[8], b[8];
      *reinterpret_cast<float4*>(a) = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      *reinterpret_cast<float4*>(a + 4) = *reinterpret_cast<const float4*>(&As[kk][64 + ty * 4]);
      *reinterpret_cast<float4*>(b) = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      *reinterpret_cast<float4*>(b + 4) = *reinterpret_cast<const float4*>(&Bs[kk][64 + tx * 4]);
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = mi * GT + (i < 4 ? ty * 4 + i : 64 + ty * 4 + i - 4);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int n = ni * GT + (j < 4 ? tx * 4 + j : 64 + tx * 4 + j - 4);
      if (m < n_cols && n < n_cols) {
        const unsigned long long v = (unsigned long long)to_fx((double)acc[i][j], SCF_GRAM_SHIFT);
        atomicAdd((unsigned long long*)(g_fx + (int64_t)m * ldg + n), v);
        if (mi != ni) atomicAdd((unsigned long long*)(g_fx + (int64_t)n * ldg + m), v);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Projection Y = Z V: CTA = 64 rows x 64 output columns, BK = 16, 256 threads, 4x4 micro tile.
constexpr int PM = 64, PN = 64, PK = 16;

__global__ void __launch_bounds__(256) project_kernel(const float* __restrict__ z, int64_t ldz, int64_t n_rows,
                                                      int n_cols, const float* __restrict__ v, int64_t ldv, int dims,
                                                      float* __restrict__ y, int64_t ldy) {
  __shared__ __align__(16) float As[PK][PM + 4];
  __shared__ __align__(16) float Bs[PK][PN];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int64_t row0 = (int64_t)blockIdx.x * PM;
  const int col0 = blockIdx.y * PN;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const int ar = tid >> 2, ak = (tid & 3) * 4;   // A: 64 rows x 4 float4 along k
  const int bk = tid >> 4, bc = (tid & 15) * 4;  // B: 16 k x 16 float4 along columns
  for (int k0 = 0; k0 < n_cols; k0 += PK) {
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
    if (row0 + ar < n_rows && k0 + ak < ldz) a = __ldg(reinterpret_cast<const float4*>(z + (row0 + ar) * ldz + k0 + ak));
    if (k0 + ak + 0 >= n_cols) a.x = 0.f;
    if (k0 + ak + 1 >= n_cols) a.y = 0.f;
    if (k0 + ak + 2 >= n_cols) a.z = 0.f;
    if (k0 + ak + 3 >= n_cols) a.w = 0.f;
    if (k0 + bk < n_cols && col0 + bc < ldv) b = __ldg(reinterpret_cast<const float4*>(v + (int64_t)(k0 + bk) * ldv + col0 + bc));
    As[ak + 0][ar] = a.x;
    As[ak + 1][ar] = a.y;
    As[ak + 2][ar] = a.z;
    As[ak + 3][ar] = a.w;
    *reinterpret_cast<float4*>(&Bs[bk][bc]) = b;
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < PK; ++kk) {
      float av[4], bv[4];
      *reinterpret_cast<float4*>(av) = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      *reinterpret_cast<float4*>(bv) = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t r = row0 + ty * 4 + i;
    if (r >= n_rows) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = col0 + tx * 4 + j;
      if (c < ldy) y[r * ldy + c] = c < dims ? acc[i][j] : 0.f;
    }
  }
}

}  // namespace

int32_t scf_gram_tc(const float* z, const float* z_lo, int64_t ldz, int64_t n_rows, int32_t n_cols, int64_t* g_fx,
                    int64_t ldg, int32_t mode, cudaStream_t stream);
int32_t scf_gram_mirror(int64_t* g_fx, int32_t n_cols, int64_t ldg, cudaStream_t stream);

extern "C" int32_t scf_gram_symmetrize(int64_t* g_fx, int32_t n_cols, int64_t ldg, void* stream) {
  SCF_ARG(g_fx, "null pointer");
  SCF_ARG(n_cols > 0 && ldg >= n_cols, "bad sizes");
  return scf_gram_mirror(g_fx, n_cols, ldg, (cudaStream_t)stream);
}

extern "C" int32_t scf_gram_accumulate(const float* z, const float* z_lo, int64_t ldz, int64_t n_rows,
                                       int32_t n_cols, int64_t* g_fx, int64_t ldg, int32_t mode, void* stream) {
  SCF_ARG(z && g_fx, "null pointer");
  SCF_ARG(n_rows >= 0 && n_cols > 0 && ldz >= n_cols && (ldz & 3) == 0 && ldg >= n_cols, "bad sizes");
  SCF_ARG(mode == 0 || mode == 1 || mode == 3, "mode must be 0 (fp32), 1 (tf32) or 3 (3xtf32)");
  if (n_rows == 0) return 0;
  if (mode != 0) return scf_gram_tc(z, z_lo, ldz, n_rows, n_cols, g_fx, ldg, mode, (cudaStream_t)stream);
  const int n_tiles = (n_cols + GT - 1) / GT;
  const int64_t n_slabs = (n_rows + SCF_GRAM_SLAB - 1) / SCF_GRAM_SLAB;
  SCF_ARG(n_slabs <= 65535, "too many slabs for one launch (split the rows)");
  dim3 grid(n_tiles * (n_tiles + 1) / 2, (unsigned)n_slabs);
  gram_simt_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(z, ldz, n_rows, n_cols, (long long*)g_fx, ldg, n_tiles);
  return scf_check_launch("scf_gram_accumulate");
}

extern "C" int32_t scf_project(const float* z, int64_t ldz, int64_t n_rows, int32_t n_cols, const float* v,
                               int64_t ldv, int32_t dims, float* y, int64_t ldy, void* stream) {
  SCF_ARG(z && v && y, "null pointer");
  SCF_ARG(n_rows >= 0 && n_cols > 0 && dims > 0 && ldz >= n_cols && ldv >= dims && ldy >= dims, "bad sizes");
  SCF_ARG((ldz & 3) == 0 && (ldv & 3) == 0, "ldz and ldv must be multiples of 4");
  if (n_rows == 0) return 0;
  dim3 grid((unsigned)((n_rows + PM - 1) / PM), (unsigned)((ldy + PN - 1) / PN));
  project_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(z, ldz, n_rows, n_cols, v, ldv, dims, y, ldy);
  return scf_check_launch("scf_project");
}
