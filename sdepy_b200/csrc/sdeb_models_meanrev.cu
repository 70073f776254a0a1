// Ornstein-Uhlenbeck / Hull-White / Cox-Ingersoll-Ross kernels
#include "sdeb_internal.h"
using namespace sdeb;

bool sdeb_lookup_meanrev(int64_t model, int64_t n, ModelInfo& mi) {
    switch (model) {
    case SDEB_MODEL_MEANREV:
        if (n == 1) { mi = info_of<MeanRevertingSDE<1, false>>(); return true; }
        if (n == 2) { mi = info_of<MeanRevertingSDE<2, false>>(); return true; }
        if (n == 3) { mi = info_of<MeanRevertingSDE<3, false>>(); return true; }
        if (n == 4) { mi = info_of<MeanRevertingSDE<4, false>>(); return true; }
        return false;
    case SDEB_MODEL_HULL_WHITE:
        if (n == 1) { mi = info_of<MeanRevertingSDE<1, true>>(); return true; }
        if (n == 2) { mi = info_of<MeanRevertingSDE<2, true>>(); return true; }
        if (n == 3) { mi = info_of<MeanRevertingSDE<3, true>>(); return true; }
        if (n == 4) { mi = info_of<MeanRevertingSDE<4, true>>(); return true; }
        return false;
    case SDEB_MODEL_CIR:
        if (n == 1) { mi = info_of<CoxIngersollRossSDE<1>>(); return true; }
        if (n == 2) { mi = info_of<CoxIngersollRossSDE<2>>(); return true; }
        return false;
    default:
        return false;
    }
}
