// Heston kernels (the north-star hot path)
#include "sdeb_internal.h"
using namespace sdeb;

bool sdeb_lookup_heston(int64_t model, int64_t n, ModelInfo& mi) {
    switch (model) {
    case SDEB_MODEL_HESTON:
        if (n == 1) { mi = info_of<HestonSDE<1, false>>(); return true; }
        if (n == 2) { mi = info_of<HestonSDE<2, false>>(); return true; }
        return false;
    case SDEB_MODEL_HESTON_FULL:
        if (n == 1) { mi = info_of<HestonSDE<1, true>>(); return true; }
        if (n == 2) { mi = info_of<HestonSDE<2, true>>(); return true; }
        return false;
    default:
        return false;
    }
}
