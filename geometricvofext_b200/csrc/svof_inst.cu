// svof_inst.cu -- explicit instantiation of the capacity-variant kernels.
// Compiled once per variant: nvcc -DSV_VARIANT=0|1|2|3|4 (see geometricvofext_b200/build.py).
#include "svof_geom_kernels.cuh"

namespace svof {
#if SV_VARIANT == 0
template struct GeoLaunch<CapsHex>;
#elif SV_VARIANT == 1
template struct GeoLaunch<CapsSmall>;
#elif SV_VARIANT == 2
template struct GeoLaunch<CapsPoly>;
#elif SV_VARIANT == 3
template struct GeoLaunch<CapsSplit>;
#else
template struct GeoLaunch<CapsHexSplit>;
#endif
}  // namespace svof
