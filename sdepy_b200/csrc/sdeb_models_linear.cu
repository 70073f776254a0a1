// wiener / lognorm / jump-diffusion kernels (LinearSDE instantiations)
#include "sdeb_internal.h"
using namespace sdeb;

bool sdeb_lookup_linear(int64_t model, int64_t n, ModelInfo& mi) {
    switch (model) {
    case SDEB_MODEL_LINEAR:
        if (n == 1) { mi = info_of<LinearSDE<1, false, false>>(); return true; }
        if (n == 2) { mi = info_of<LinearSDE<2, false, false>>(); return true; }
        if (n == 3) { mi = info_of<LinearSDE<3, false, false>>(); return true; }
        if (n == 4) { mi = info_of<LinearSDE<4, false, false>>(); return true; }
        return false;
    case SDEB_MODEL_LINEAR_LOG:
        if (n == 1) { mi = info_of<LinearSDE<1, true, false>>(); return true; }
        if (n == 2) { mi = info_of<LinearSDE<2, true, false>>(); return true; }
        if (n == 3) { mi = info_of<LinearSDE<3, true, false>>(); return true; }
        if (n == 4) { mi = info_of<LinearSDE<4, true, false>>(); return true; }
        return false;
    case SDEB_MODEL_JUMPDIFF:
        if (n == 1) { mi = info_of<LinearSDE<1, true, true>>(); return true; }
        if (n == 2) { mi = info_of<LinearSDE<2, true, true>>(); return true; }
        return false;
    default:
        return false;
    }
}
