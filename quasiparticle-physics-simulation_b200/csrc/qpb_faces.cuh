// Face couplings of one cell of the masked 5-point operator (shared by the generic sweep kernels and the Krylov path).
#pragma once
#include "qpb_internal.h"

namespace {

struct Faces {
    double eL, eR, eU, eD;  // face couplings (already multiplied by dt/2/dx^2 and D)
    double gbx, gby;        // boundary diagonal terms in x / y
};

template <bool VARD>
__device__ __forceinline__ Faces load_faces(int c, int nx, unsigned fl, double a, const double *__restrict__ bcx,
                                            const double *__restrict__ bcy, const double *__restrict__ ex,
                                            const double *__restrict__ ey, const double *__restrict__ gbx,
                                            const double *__restrict__ gby) {
    Faces f;
    if (VARD) {
        f.eL = (fl & QPB_LK_L) ? ex[c] : 0.0;
        f.eR = (fl & QPB_LK_R) ? ex[c + 1] : 0.0;
        f.eU = (fl & QPB_LK_U) ? ey[c] : 0.0;
        f.eD = (fl & QPB_LK_D) ? ey[c + nx] : 0.0;
        f.gbx = gbx[c];
        f.gby = gby[c];
    } else {
        f.eL = (fl & QPB_LK_L) ? a : 0.0;
        f.eR = (fl & QPB_LK_R) ? a : 0.0;
        f.eU = (fl & QPB_LK_U) ? a : 0.0;
        f.eD = (fl & QPB_LK_D) ? a : 0.0;
        f.gbx = a * bcx[c];
        f.gby = a * bcy[c];
    }
    return f;
}

}  // namespace
