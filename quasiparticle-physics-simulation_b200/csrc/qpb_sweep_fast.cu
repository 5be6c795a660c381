// Chunked, table-driven tridiagonal sweeps for uniform D (fast path).  Placeholder: not enabled yet.
#include "qpb_internal.h"

int qpbk_prepare_fast(qpb_ctx *c, DiffSlot &s) {
    (void)c;
    s.fast = false;
    return QPB_OK;
}

int qpbk_sweep_fast(qpb_ctx *c, DiffSlot &s, int dir, int iter, int mode) {
    (void)c; (void)s; (void)dir; (void)iter; (void)mode;
    qpb_set_error("fast sweep path is not built");
    return QPB_E_INVALID;
}
