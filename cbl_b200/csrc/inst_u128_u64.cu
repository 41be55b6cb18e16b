// Index<u128, uint64_t>: see cbl_index_impl.cuh
#include "cbl_index_impl.cuh"
namespace cbl {
CBL_INSTANTIATE_INDEX(make_index_u128_u64, u128, uint64_t)
}
