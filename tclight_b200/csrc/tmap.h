// Host-side TMA tensor-map construction (cuTensorMapEncodeTiled resolved at run time through
// cudaGetDriverEntryPoint so the library links against cudart only).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace tcl {

// Builds a rank-`rank` tiled tensor map over 16-bit elements.
//   dims[i]      : extent of dimension i (dimension 0 is the contiguous one)
//   strides_b[i] : byte stride of dimension i+1 (rank-1 entries, multiples of 16)
//   box[i]       : box extent in tensor space
//   estr[i]      : traversal stride (1 = dense)
//   swizzle_bytes: 0, 32, 64 or 128
// Returns 0 on success, else sets the last-error string and returns a negative code.
int make_tmap(CUtensorMap* out, const void* base, bool bf16, int rank, const uint64_t* dims,
              const uint64_t* strides_b, const uint32_t* box, const uint32_t* estr,
              int swizzle_bytes);

}  // namespace tcl
