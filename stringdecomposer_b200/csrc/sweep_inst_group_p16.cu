// instantiations of sweep_group_kernel<Packed16, C, T>
#include "sweep_kernel.cuh"
namespace sdb {
SD_INSTANTIATE_GROUP(sweep_group_lookup_p16, Packed16)
}
