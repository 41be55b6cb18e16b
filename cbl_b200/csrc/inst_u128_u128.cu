// Index<u128, u128>: see cbl_index_impl.cuh
#include "cbl_index_impl.cuh"
namespace cbl {
CBL_INSTANTIATE_INDEX(make_index_u128_u128, u128, u128)
}
