// instantiations of sweep_group_kernel<Scalar32, C, T>
#include "sweep_kernel.cuh"
namespace sdb {
SD_INSTANTIATE_GROUP(sweep_group_lookup_s32, Scalar32)
}
