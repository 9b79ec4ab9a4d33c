// Host-side TMA tensor-map cache shared by the tcgen05 GEMM and attention kernels.
#pragma once
#include <cuda.h>

namespace fvqa {

// 2-D h16 row-major tensor [rows, cols] with leading dimension `ld` (elements), tiled in boxes of
// 64 columns (one 128-byte swizzle row) x box_rows rows, SWIZZLE_128B. Cached per (ptr, shape, box).
// Requires fvqa_init(). Returns FVQA_OK or an error code (message via fvqa_last_error()).
int get_tmap(const void* ptr, int rows, int cols, int ld, int box_rows, CUtensorMap* out);
// Same boxes over a 3-D (column, position, sequence) view: rows of a box beyond seq_len are out of bounds.
int get_tmap_seq(const void* ptr, int n_seq, int seq_len, int cols, int ld, int box_rows, CUtensorMap* out);
int num_sms();

}  // namespace fvqa
