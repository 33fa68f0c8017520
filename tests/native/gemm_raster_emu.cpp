// host build of the GEMM tile-order function (csrc/b2_gemm_raster.h) for tests/test_gemm_raster.py
#include "b2_gemm_raster.h"

extern "C" void tile_coords(int tile, int num_m, int num_n, int* m_blk, int* n_blk) {
  int m, n;
  b2_gemm_tile_coords(tile, num_m, num_n, m, n);
  *m_blk = m;
  *n_blk = n;
}
extern "C" int raster_group(void) { return B2_GEMM_RASTER_GROUP; }
