// explicit instantiations of the deferred-jump sweep, packed s16x2 policy (one translation unit so that it builds in parallel)
#include "sweep_lat_kernel.cuh"
namespace sdb {
SD_INSTANTIATE_LAT(sweep_lat_lookup_p16, Packed16)
}
