// fp32 parity mode ON THE TENSOR CORES: every fp32 operand is split into three bf16 terms
//     v = v_h + v_m + v_l,   v_h = bf16(v), v_m = bf16(v - v_h), v_l = bf16(v - v_h - v_m)
// (8 + 8 + 8 = 24 significand bits: the split is exact up to the last bit of fp32) and a product A B^T is formed as
// the six partial products whose terms carry at least 2^-24 of the result,
//     A B^T  ~=  A_l B_h + A_m B_m + A_h B_l + A_m B_h + A_h B_m + A_h B_h,
// accumulated in fp32 by the SAME tcgen05 kernel as the bf16 mode -- as ONE GEMM whose contraction dimension is six
// copies long: A'' = [A_l A_m A_h A_m A_h A_h], B'' = [B_h B_m B_l B_h B_m B_h] along K -- SMALLEST terms first: an
// unsplit accumulation chain then collects the 2^-16 and 2^-8 sized corrections before the leading products arrive
// (largest first measured 2.1e-5 max relative error at 16384 x 2048 x 1000, where K runs unsplit, against 2.9e-6 at
// 256 rows).  This file holds the two
// operand-expansion kernels (K along the columns of a row-major matrix / K along its rows).  Replaces the FFMA sgemm
// (gemm_f32.cu) as the 1e-5 parity path of nn.Linear (cls/resnet_pytorch.py:219,293; bbox_head.py:118) when selected.
#include "common.cuh"

namespace iif {

__device__ __forceinline__ void split3(float v, uint16_t& h, uint16_t& m, uint16_t& l) {
  const __nv_bfloat16 bh = __float2bfloat16_rn(v);
  const float r1 = v - __bfloat162float(bh);
  const __nv_bfloat16 bm = __float2bfloat16_rn(r1);
  const float r2 = r1 - __bfloat162float(bm);
  const __nv_bfloat16 bl = __float2bfloat16_rn(r2);
  h = *reinterpret_cast<const uint16_t*>(&bh);
  m = *reinterpret_cast<const uint16_t*>(&bm);
  l = *reinterpret_cast<const uint16_t*>(&bl);
}

// which term (0 = h, 1 = m, 2 = l) the k-th of the six copies holds: the A side and the B side of the product
__constant__ int kTermA[6] = {2, 1, 0, 1, 0, 0};
__constant__ int kTermB[6] = {0, 1, 2, 0, 1, 0};

// K along the COLUMNS: in [rows, cols] fp32 -> out [rows, 6 * cols_pad] bf16, copy k at columns [k * cols_pad, ..)
// (cols_pad = cols rounded up to 8 so that every copy starts 16-byte aligned; the padding is zero)
__global__ void __launch_bounds__(256) split3_cols_kernel(const float* __restrict__ in, int64_t ldi, int64_t rows, int cols,
                                                          int cols_pad, int side, uint16_t* __restrict__ out, int64_t ldo) {
  const int64_t total = rows * (int64_t)cols_pad;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / cols_pad;
    const int c = (int)(i - r * cols_pad);
    uint16_t t[3] = {0, 0, 0};
    if (c < cols) split3(__ldg(in + r * ldi + c), t[0], t[1], t[2]);
    uint16_t* o = out + r * ldo + c;
#pragma unroll
    for (int k = 0; k < 6; ++k) o[(int64_t)k * cols_pad] = t[side ? kTermB[k] : kTermA[k]];
  }
}

// K along the ROWS: in [rows, cols] fp32 -> out [6 * rows_pad, cols] bf16, copy k at rows [k * rows_pad, ..)
// (rows_pad = rows rounded up to 8, like cols_pad above: both operands of a product index K the same way)
__global__ void __launch_bounds__(256) split3_rows_kernel(const float* __restrict__ in, int64_t ldi, int64_t rows,
                                                          int64_t rows_pad, int cols, int side, uint16_t* __restrict__ out,
                                                          int64_t ldo) {
  const int64_t total = rows_pad * (int64_t)cols;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / cols;
    const int c = (int)(i - r * cols);
    uint16_t t[3] = {0, 0, 0};
    if (r < rows) split3(__ldg(in + r * ldi + c), t[0], t[1], t[2]);
#pragma unroll
    for (int k = 0; k < 6; ++k) out[((int64_t)k * rows_pad + r) * ldo + c] = t[side ? kTermB[k] : kTermA[k]];
  }
}

}  // namespace iif

using namespace iif;

extern "C" int iif_split3_bf16(const float* in, int64_t ldi, int64_t rows, int64_t cols, int k_along_rows, int side_b,
                               void* out, int64_t ldo, void* stream) {
  if (rows < 0 || cols < 0 || ldi < cols || cols > (1 << 30)) return IIF_EINVAL;
  if (rows == 0 || cols == 0) return IIF_OK;
  if (!in || !out) return IIF_EINVAL;
  const int cols_pad = (int)((cols + 7) / 8 * 8);
  if (k_along_rows ? ldo < cols : ldo < 6 * (int64_t)cols_pad) return IIF_EINVAL;
  const int64_t rows_pad = (rows + 7) / 8 * 8;
  const int64_t items = k_along_rows ? rows_pad * cols : rows * cols_pad;
  const unsigned grid = (unsigned)((items + 255) / 256 < 16 * kNumSMs ? (items + 255) / 256 : 16 * kNumSMs);
  uint16_t* o = reinterpret_cast<uint16_t*>(out);
  if (k_along_rows) split3_rows_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(in, ldi, rows, rows_pad, (int)cols, side_b ? 1 : 0, o, ldo);
  else split3_cols_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(in, ldi, rows, (int)cols, cols_pad, side_b ? 1 : 0, o, ldo);
  return launch_status();
}
