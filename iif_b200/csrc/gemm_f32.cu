// fp32 (FFMA) GEMMs of the head: the 1e-5 parity mode of fc_cls forward / dX / dW.
// One generic kernel  OUT[M,N] = alpha * sum_k A(m,k) * B(n,k) (+ bias[n]) with stride-described
// operands, so the three products of AddmmBackward (a1, a10) share the code:
//   fwd: A = X[B,D]  (k contiguous)   B = W[C,D]      (k contiguous)
//   dX : A = dZ[B,C] (k contiguous)   B = W[C,D] read as B(n=d, k=c)   (n contiguous)
//   dW : A = dZ[B,C] read as A(m=c,k=b) (m contiguous)   B = X[B,D] read as B(n=d,k=b) (n contiguous)
// 64x64x16 tiles, 256 threads, 4x4 register micro-tile, fixed k order -> deterministic.
#include "common.cuh"

namespace iif {

struct SgemmArgs {
  const float* A; int64_t sam, sak;
  const float* B; int64_t sbn, sbk;
  int M, N, K;
  const float* alpha; const float* bias; const float* col_scale;
  float* out; int64_t ldo; float* out2; int64_t ldo2;
};

constexpr int BM = 64, BN = 64, BK = 16, PAD = 4;

template <bool MN_CONTIG>
__device__ __forceinline__ void load_tile(const float* __restrict__ P, int64_t s_mn, int64_t s_k, int mn0, int k0, int MN,
                                          int K, float (&r)[4]) {
  const int tid = threadIdx.x;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    int mn, k;
    if (MN_CONTIG) { mn = tid & 63; k = (tid >> 6) + 4 * j; }
    else { k = tid & 15; mn = (tid >> 4) + 16 * j; }
    const int gm = mn0 + mn, gk = k0 + k;
    r[j] = (gm < MN && gk < K) ? __ldg(P + gm * s_mn + gk * s_k) : 0.f;
  }
}
template <bool MN_CONTIG>
__device__ __forceinline__ void store_tile(float (*S)[BM + PAD], const float (&r)[4]) {
  const int tid = threadIdx.x;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    int mn, k;
    if (MN_CONTIG) { mn = tid & 63; k = (tid >> 6) + 4 * j; }
    else { k = tid & 15; mn = (tid >> 4) + 16 * j; }
    S[k][mn] = r[j];
  }
}

template <bool A_MN, bool B_MN>
__global__ void __launch_bounds__(256) sgemm_kernel(const SgemmArgs a) {
  __shared__ __align__(16) float As[BK][BM + PAD];
  __shared__ __align__(16) float Bs[BK][BN + PAD];
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[4][4] = {};
  float ra[4], rb[4];
  load_tile<A_MN>(a.A, a.sam, a.sak, m0, 0, a.M, a.K, ra);
  load_tile<B_MN>(a.B, a.sbn, a.sbk, n0, 0, a.N, a.K, rb);
  for (int k0 = 0; k0 < a.K; k0 += BK) {
    store_tile<A_MN>(As, ra);
    store_tile<B_MN>(Bs, rb);
    __syncthreads();
    if (k0 + BK < a.K) {
      load_tile<A_MN>(a.A, a.sam, a.sak, m0, k0 + BK, a.M, a.K, ra);
      load_tile<B_MN>(a.B, a.sbn, a.sbk, n0, k0 + BK, a.N, a.K, rb);
    }
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 av = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 bv = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float am[4] = {av.x, av.y, av.z, av.w}, bn[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(am[i], bn[j], acc[i][j]);
    }
    __syncthreads();
  }
  const float alpha = a.alpha ? __ldg(a.alpha) : 1.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= a.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= a.N) continue;
      float v = acc[i][j] * alpha;
      if (a.bias) v += __ldg(a.bias + n);
      if (a.out) a.out[(int64_t)m * a.ldo + n] = v;
      if (a.out2) a.out2[(int64_t)m * a.ldo2 + n] = v * __ldg(a.col_scale + n);
    }
  }
}

static int launch_sgemm(const SgemmArgs& a, cudaStream_t st) {
  if (a.M == 0 || a.N == 0) return IIF_OK;
  dim3 grid((a.N + BN - 1) / BN, (a.M + BM - 1) / BM);
  const bool amn = a.sam == 1 && a.sak != 1, bmn = a.sbn == 1 && a.sbk != 1;
  if (!amn && !bmn) sgemm_kernel<false, false><<<grid, 256, 0, st>>>(a);
  else if (!amn && bmn) sgemm_kernel<false, true><<<grid, 256, 0, st>>>(a);
  else if (amn && !bmn) sgemm_kernel<true, false><<<grid, 256, 0, st>>>(a);
  else sgemm_kernel<true, true><<<grid, 256, 0, st>>>(a);
  return launch_status();
}

static bool bad_dims(int64_t B, int64_t D, int64_t C) {
  return B < 0 || D <= 0 || C <= 0 || B > INT32_MAX || D > INT32_MAX || C > INT32_MAX;
}

}  // namespace iif

using namespace iif;

extern "C" int iif_linear_fwd_f32(const float* x, int64_t ldx, const float* w, int64_t ldw, const float* bias,
                                  const float* col_scale, float* z, int64_t ldz, float* zs, int64_t ldzs, int64_t B,
                                  int64_t D, int64_t C, void* stream) {
  if (bad_dims(B, D, C) || !w || (B > 0 && !x) || (!z && !zs) || ldx < D || ldw < D) return IIF_EINVAL;
  if ((z && ldz < C) || (zs && (ldzs < C || !col_scale))) return IIF_EINVAL;
  SgemmArgs a{x, ldx, 1, w, ldw, 1, (int)B, (int)C, (int)D, nullptr, bias, col_scale, z, ldz, zs, ldzs};
  return launch_sgemm(a, (cudaStream_t)stream);
}

extern "C" int iif_linear_bwd_dx_f32(const float* dz, int64_t lddz, const float* w, int64_t ldw, const float* alpha_dev,
                                     float* dx, int64_t lddx, int64_t B, int64_t D, int64_t C, void* stream) {
  if (bad_dims(B, D, C) || !w || !dx || (B > 0 && !dz) || lddz < C || ldw < D || lddx < D) return IIF_EINVAL;
  SgemmArgs a{dz, lddz, 1, w, 1, ldw, (int)B, (int)D, (int)C, alpha_dev, nullptr, nullptr, dx, lddx, nullptr, 0};
  return launch_sgemm(a, (cudaStream_t)stream);
}

extern "C" int iif_linear_bwd_dw_f32(const float* dz, int64_t lddz, const float* x, int64_t ldx, const float* alpha_dev,
                                     float* dw, int64_t lddw, int64_t B, int64_t D, int64_t C, void* stream) {
  if (bad_dims(B, D, C) || !dw || (B > 0 && (!dz || !x)) || lddz < C || ldx < D || lddw < D) return IIF_EINVAL;
  SgemmArgs a{dz, 1, lddz, x, 1, ldx, (int)C, (int)D, (int)B, alpha_dev, nullptr, nullptr, dw, lddw, nullptr, 0};
  return launch_sgemm(a, (cudaStream_t)stream);
}
