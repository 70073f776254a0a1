// Preset kernel instantiations.  This file is compiled SDEB_N_UNITS times with
// -DSDEB_UNIT=k: each unit instantiates (and registers) a few model functors, so
// the units build in parallel (one integrate_body instantiation is ~20 s of
// nvcc; see _build.py).
#include "sdeb_internal.h"
using namespace sdeb;

#define SDEB_CAT2(a, b) a##b
#define SDEB_CAT(a, b) SDEB_CAT2(a, b)
#define REG(MODEL, N, ...) \
    if (model == MODEL && n == N) { mi = info_of<__VA_ARGS__>(); return true; }

bool SDEB_CAT(sdeb_lookup_unit_, SDEB_UNIT)(int64_t model, int64_t n, ModelInfo& mi) {
#if SDEB_UNIT == 0          // Heston: the north-star hot path
    REG(SDEB_MODEL_HESTON, 1, HestonSDE<1, false>)
    REG(SDEB_MODEL_HESTON, 2, HestonSDE<2, false>)
#elif SDEB_UNIT == 1
    REG(SDEB_MODEL_HESTON_FULL, 1, HestonSDE<1, true>)
    REG(SDEB_MODEL_HESTON_FULL, 2, HestonSDE<2, true>)
#elif SDEB_UNIT == 2        // wiener / lognorm / jump-diffusion (LinearSDE)
    REG(SDEB_MODEL_LINEAR, 1, LinearSDE<1, false, false>)
    REG(SDEB_MODEL_LINEAR, 2, LinearSDE<2, false, false>)
#elif SDEB_UNIT == 3
    REG(SDEB_MODEL_LINEAR, 3, LinearSDE<3, false, false>)
    REG(SDEB_MODEL_LINEAR, 4, LinearSDE<4, false, false>)
#elif SDEB_UNIT == 4
    REG(SDEB_MODEL_LINEAR_LOG, 1, LinearSDE<1, true, false>)
    REG(SDEB_MODEL_LINEAR_LOG, 2, LinearSDE<2, true, false>)
#elif SDEB_UNIT == 5
    REG(SDEB_MODEL_LINEAR_LOG, 3, LinearSDE<3, true, false>)
    REG(SDEB_MODEL_LINEAR_LOG, 4, LinearSDE<4, true, false>)
#elif SDEB_UNIT == 6
    REG(SDEB_MODEL_JUMPDIFF, 1, LinearSDE<1, true, true>)
    REG(SDEB_MODEL_JUMPDIFF, 2, LinearSDE<2, true, true>)
#elif SDEB_UNIT == 7        // Ornstein-Uhlenbeck / Hull-White / Cox-Ingersoll-Ross
    REG(SDEB_MODEL_MEANREV, 1, MeanRevertingSDE<1, false>)
    REG(SDEB_MODEL_MEANREV, 2, MeanRevertingSDE<2, false>)
    REG(SDEB_MODEL_CIR, 1, CoxIngersollRossSDE<1>)
#elif SDEB_UNIT == 8
    REG(SDEB_MODEL_MEANREV, 3, MeanRevertingSDE<3, false>)
    REG(SDEB_MODEL_MEANREV, 4, MeanRevertingSDE<4, false>)
    REG(SDEB_MODEL_CIR, 2, CoxIngersollRossSDE<2>)
#elif SDEB_UNIT == 9
    REG(SDEB_MODEL_HULL_WHITE, 1, MeanRevertingSDE<1, true>)
    REG(SDEB_MODEL_HULL_WHITE, 2, MeanRevertingSDE<2, true>)
#elif SDEB_UNIT == 10
    REG(SDEB_MODEL_HULL_WHITE, 3, MeanRevertingSDE<3, true>)
    REG(SDEB_MODEL_HULL_WHITE, 4, MeanRevertingSDE<4, true>)
#else
#error "SDEB_UNIT out of range"
#endif
    return false;
}
