// mpi4jax_b200 -- tile order of the persistent GEMM kernels (csrc/b2_gemm.cu).
//
// tile index -> (m_blk, n_blk).  B2_GEMM_RASTER_GROUP = 0 (default, the order that was measured):
// row-major over the tile grid.  B2_GEMM_RASTER_GROUP = G > 0: tiles are visited in bands of G
// M-blocks, column by column inside a band, so that the ~148 tiles in flight cover G rows x ~148/G
// columns of the tile grid: A row-panels are reused ~148/G times and B column-panels G times from
// L2 instead of (148 / num_n) rows x all columns (B larger than L2 at 8192^3).  Every rank uses
// the same order (the fused all-reduce pairs tile t of all ranks).  Not yet measured: default off.
#pragma once

#ifndef B2_GEMM_RASTER_GROUP
#define B2_GEMM_RASTER_GROUP 0
#endif

#ifdef __CUDACC__
#define B2_HD __host__ __device__
#else
#define B2_HD
#endif

B2_HD inline void b2_gemm_tile_coords(int tile, int num_m, int num_n, int& m_blk, int& n_blk) {
#if B2_GEMM_RASTER_GROUP > 0
  const int G = B2_GEMM_RASTER_GROUP;
  const int per_band = G * num_n;
  const int band = tile / per_band;
  const int first_m = band * G;
  const int rows = (num_m - first_m) < G ? (num_m - first_m) : G;   // the last band may be thinner
  const int r = tile - band * per_band;
  m_blk = first_m + r % rows;
  n_blk = r / rows;
#else
  (void)num_m;
  m_blk = tile / num_n;
  n_blk = tile % num_n;
#endif
}
