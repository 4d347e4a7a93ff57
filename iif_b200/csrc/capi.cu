// C-ABI glue: version / error strings / launch counter and the one-call head step.
#include "common.cuh"

namespace iif {
std::atomic<uint64_t> g_launches{0};
}

extern "C" int iif_abi_version(void) { return IIF_B200_ABI_VERSION; }

extern "C" uint64_t iif_launch_count(void) { return iif::g_launches.load(std::memory_order_relaxed); }

extern "C" const char* iif_error_string(int code) {
  switch (code) {
    case IIF_OK: return "ok";
    case IIF_EINVAL: return "invalid argument (null pointer, negative size, bad enum or leading dimension)";
    case IIF_EALIGN: return "pointer / leading dimension violates the documented alignment";
    case IIF_EUNSUPPORTED: return "shape outside the supported range";
    case IIF_EWORKSPACE: return "workspace missing or too small (see iif_gemm_ws_bytes)";
    case IIF_EDRIVER: return "cuTensorMapEncodeTiled unavailable or failed";
    default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "unknown iif_b200 error";
  }
}

// fc_cls -> IIF softmax-CE fwd+bwd -> {dX, dW, db} on one stream: ONE persistent launch when the shape qualifies
// (head_fused.cu), else 2 launches when the loss rows can ride in the backward launch (iif_loss_linear_bwd_bf16),
// else 3.
extern "C" int iif_head_fwd_bwd_bf16(const iif_head_args* h, void* stream) {
  if (!h || !h->x || !h->w || !h->label || !h->z || !h->dz_bf16 || !h->dw) return IIF_EINVAL;
  if (h->lddz % 8 != 0) return IIF_EALIGN;
  int rc;
  if (!(h->flags & IIF_HEAD_NO_PERSISTENT)) {          // the whole step in ONE persistent launch (head_fused.cu)
    rc = iif::head_fused_launch(h, stream, false);
    if (rc != IIF_EUNSUPPORTED) return rc;
  }
  rc = iif_linear_fwd_bf16(h->x, h->ldx, h->w, h->ldw, h->bias, nullptr, h->z, h->ldz, nullptr, 0, h->B, h->D, h->C,
                               h->ws, h->ws_bytes, stream);
  if (rc) return rc;
  if (!(h->flags & IIF_HEAD_NO_FUSED_LOSS)) {
    rc = iif_loss_linear_bwd_bf16(h, stream);
    if (rc != IIF_EUNSUPPORTED) return rc;
  }
  rc = iif_softmax_ce_fwd_bwd(h->z, h->ldz, h->iif, h->label, h->class_weight, h->sample_weight, h->ignore_index,
                              h->scale, h->B, h->C, h->loss_i, h->loss_sum, nullptr, 0, h->dz_bf16, h->lddz, nullptr,
                              h->argmax, h->rank, h->acc_counts, h->scratch, stream);
  if (rc) return rc;
  return iif_linear_bwd_bf16(h->dz_bf16, h->lddz, h->x, h->ldx, h->w, h->ldw, nullptr, h->dx, h->dx_dtype, h->lddx, h->dw,
                             h->lddw, h->db, h->B, h->D, h->C, h->ws, h->ws_bytes, stream);
}
