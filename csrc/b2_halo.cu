// mpi4jax_b200 -- fused multi-field 2-D halo exchange.
//
// The reference's flagship workload (examples/shallow_water.py:172-264,
// `enforce_boundaries`) exchanges a 1-cell halo with up to four neighbours by
// issuing, per field, 2 sendrecv + 1 send + 1 recv custom calls, each of them a
// blocking MPI call behind a full stream synchronisation, with XLA gather /
// dynamic-update-slice kernels around them for the strided column pack and the
// halo unpack (48 custom calls per model step at 8 ranks).
//
// Here ONE kernel launch exchanges the halos of up to B2_HALO_MAX_FIELDS fields
// in all four directions: strided column pack, NVLink put straight into the
// neighbour's symmetric receive buffer, release/acquire flag handshake, halo
// unpack and the physical wall conditions, all fused.  Two phases (columns,
// then rows carrying the fresh corner cells) reproduce exactly the corner
// values of the reference's west/north/east/south message order.  Receive
// buffers are double-buffered on a device-side ticket, so the kernel needs no
// extra "ready" handshake and is CUDA-graph replayable.
#include <cstdio>
#include <cstring>

#include "b2_device.cuh"
#include "b2_runtime.h"

extern "C" void b2_set_error(const char* fmt, ...);
extern "C" void b2_count_launch(B2Comm* c);

#define HALO_THREADS 1024
// buffer / flag index = side of the RECEIVER the data lands on
#define SIDE_W 0
#define SIDE_E 1
#define SIDE_S 2
#define SIDE_N 3

__device__ __forceinline__ float* halo_buf(const B2DevComm& c, int rank, unsigned parity, int side) {
  return (float*)(c.heap[rank] + c.lay.halo_buf_off + ((size_t)parity * 4 + side) * c.lay.halo_cap);
}
__device__ __forceinline__ unsigned* halo_flag(const B2DevComm& c, int rank, int side) {
  return (unsigned*)(c.heap[rank] + c.lay.halo_flag_off) + side * 16;   // 64 B apart
}

__global__ void __launch_bounds__(HALO_THREADS) b2_k_halo(const B2DevComm c, const B2HaloDesc d) {
  const unsigned ticket = b2_ticket_read(c.ticket + 2);
  const unsigned epoch = ticket + 1u;
  const unsigned par = ticket & 1u;
  const int t = threadIdx.x, nt = blockDim.x;
  const int ny = d.ny, nx = d.nx, F = d.nfields;

  // ---------------- phase 1: columns (full height) ----------------
  if (d.west >= 0) {   // my column 1 becomes the west neighbour's east halo
    float* dst = halo_buf(c, d.west, par, SIDE_E);
    for (int k = t; k < F * ny; k += nt) {
      const int f = k / ny, j = k - f * ny;
      dst[k] = d.field[f][(size_t)j * nx + 1];
    }
  }
  if (d.east >= 0) {   // my column nx-2 becomes the east neighbour's west halo
    float* dst = halo_buf(c, d.east, par, SIDE_W);
    for (int k = t; k < F * ny; k += nt) {
      const int f = k / ny, j = k - f * ny;
      dst[k] = d.field[f][(size_t)j * nx + (nx - 2)];
    }
  }
  __syncthreads();
  if (t == 0 && d.west >= 0) b2_st_release_sys(halo_flag(c, d.west, SIDE_E), epoch);
  if (t == 1 && d.east >= 0) b2_st_release_sys(halo_flag(c, d.east, SIDE_W), epoch);
  if (t == 0 && d.east >= 0) b2_wait_ge(c, halo_flag(c, c.rank, SIDE_E), epoch, B2_OPC_HALO, d.east);
  if (t == 1 && d.west >= 0) b2_wait_ge(c, halo_flag(c, c.rank, SIDE_W), epoch, B2_OPC_HALO, d.west);
  __syncthreads();
  if (d.east >= 0) {
    const float* src = halo_buf(c, c.rank, par, SIDE_E);
    for (int k = t; k < F * ny; k += nt) {
      const int f = k / ny, j = k - f * ny;
      d.field[f][(size_t)j * nx + (nx - 1)] = __ldcv(src + k);
    }
  }
  if (d.west >= 0) {
    const float* src = halo_buf(c, c.rank, par, SIDE_W);
    for (int k = t; k < F * ny; k += nt) {
      const int f = k / ny, j = k - f * ny;
      d.field[f][(size_t)j * nx] = __ldcv(src + k);
    }
  }
  __syncthreads();

  // ---------------- phase 2: rows (full width, fresh corners) ----------------
  if (d.north >= 0) {  // my row ny-2 becomes the north neighbour's south halo
    float* dst = halo_buf(c, d.north, par, SIDE_S);
    for (int k = t; k < F * nx; k += nt) {
      const int f = k / nx, i = k - f * nx;
      dst[k] = d.field[f][(size_t)(ny - 2) * nx + i];
    }
  }
  if (d.south >= 0) {  // my row 1 becomes the south neighbour's north halo
    float* dst = halo_buf(c, d.south, par, SIDE_N);
    for (int k = t; k < F * nx; k += nt) {
      const int f = k / nx, i = k - f * nx;
      dst[k] = d.field[f][(size_t)nx + i];
    }
  }
  __syncthreads();
  if (t == 0 && d.north >= 0) b2_st_release_sys(halo_flag(c, d.north, SIDE_S), epoch);
  if (t == 1 && d.south >= 0) b2_st_release_sys(halo_flag(c, d.south, SIDE_N), epoch);
  if (t == 0 && d.south >= 0) b2_wait_ge(c, halo_flag(c, c.rank, SIDE_S), epoch, B2_OPC_HALO, d.south);
  if (t == 1 && d.north >= 0) b2_wait_ge(c, halo_flag(c, c.rank, SIDE_N), epoch, B2_OPC_HALO, d.north);
  __syncthreads();
  if (d.south >= 0) {
    const float* src = halo_buf(c, c.rank, par, SIDE_S);
    for (int k = t; k < F * nx; k += nt) {
      const int f = k / nx, i = k - f * nx;
      d.field[f][i] = __ldcv(src + k);
    }
  }
  if (d.north >= 0) {
    const float* src = halo_buf(c, c.rank, par, SIDE_N);
    for (int k = t; k < F * nx; k += nt) {
      const int f = k / nx, i = k - f * nx;
      d.field[f][(size_t)(ny - 1) * nx + i] = __ldcv(src + k);
    }
  }
  __syncthreads();

  // ---------------- wall conditions (examples/shallow_water.py:258-262) ----------------
  for (int f = 0; f < F; ++f) {
    if (d.kind[f] == 1 && !d.periodic_x && d.at_east_wall)
      for (int j = t; j < ny; j += nt) d.field[f][(size_t)j * nx + (nx - 2)] = 0.f;
    if (d.kind[f] == 2 && d.at_north_wall)
      for (int i = t; i < nx; i += nt) d.field[f][(size_t)(ny - 2) * nx + i] = 0.f;
  }
  __syncthreads();
  if (t == 0) b2_st_volatile(c.ticket + 2, ticket + 1u);
}

extern "C" int b2_halo_exchange(B2Comm* c, const B2HaloDesc* d, cudaStream_t stream) {
  if (d->nfields < 1 || d->nfields > B2_HALO_MAX_FIELDS) {
    b2_set_error("halo_exchange: nfields must be in [1, %d]", B2_HALO_MAX_FIELDS);
    return B2_ERR_BAD_ARG;
  }
  const size_t need = (size_t)d->nfields * (size_t)(d->ny > d->nx ? d->ny : d->nx) * sizeof(float);
  if (need > c->dev.lay.halo_cap) {
    b2_set_error("halo_exchange: halo buffers too small (%zu > %zu); raise MPI4JAX_B200_HALO_BYTES",
                 need, c->dev.lay.halo_cap);
    return B2_ERR_BAD_ARG;
  }
  const int P = c->dev.size;
  const int nb[4] = {d->west, d->east, d->south, d->north};
  for (int k = 0; k < 4; ++k)
    if (nb[k] < -1 || nb[k] >= P) {
      b2_set_error("halo_exchange: invalid neighbour rank %d", nb[k]);
      return B2_ERR_BAD_ARG;
    }
  b2_k_halo<<<1, HALO_THREADS, 0, stream>>>(c->dev, *d);
  b2_count_launch(c);
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) {
    b2_set_error("halo_exchange: kernel launch failed: %s", cudaGetErrorString(err));
    return 1000 + (int)err;
  }
  return 0;
}
