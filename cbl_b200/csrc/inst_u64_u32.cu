// Index<uint64_t, uint32_t>: see cbl_index_impl.cuh
#include "cbl_index_impl.cuh"
namespace cbl {
CBL_INSTANTIATE_INDEX(make_index_u64_u32, uint64_t, uint32_t)
}
