// mpi4jax_b200 -- shallow-water step with the halo exchange FUSED INTO the stencil kernels.
//
// "Compute step followed by a collective -> one kernel over peer memory": the stand-alone
// path (b2_swe.cu + b2_halo.cu) runs 5 stencil kernels and 4 exchange kernels per model step.
// Here a step is 5 kernels and NO exchange kernel:
//
//   K1' fluxes            push edges of fe,fn,q,ke ─┐ (NVLink stores straight from registers)
//   K2' tendencies   wait/unpack fe,fn,q,ke  <──────┘ ; push edges of h',u,v ─┐
//   K3' friction-u flux   wait/unpack h',u,v <────────────────────────────────┘ ; computes the
//                         west/south halo of (fe,fn) LOCALLY  -> the reference's exchange of the
//                         friction-u fluxes (shallow_water.py:372-376) is not needed at all
//   K4' friction-u apply + friction-v flux ; push edges of fe2,fn2 ─┐
//   K5' friction-v apply  wait/unpack fe2,fn2 <─────────────────────┘
//
// Producer side: the thread that computed an edge cell (column 1 / nx-2, row 1 / ny-2, the four
// interior corners) writes it, as an 8-byte {value, flag} pair, directly into the neighbouring
// GPU's receive buffer over NVLink -- no pack pass, no fence, no separate signal (flag-in-data:
// the 8-byte store is atomic, so a reader that sees the flag sees the value).
// Consumer side: the first SWE_UNPACKERS CTAs poll the receive buffers and write the halo cells
// of the field arrays while ALL OTHER CTAs already compute every cell that does not touch a halo
// (pass 1); cells next to the boundary are computed after a device-local "halos ready" counter
// trips (pass 2).  The NVLink latency of the exchange is hidden behind the interior compute.
//
// Buffers are double-buffered on per-side message counters kept in device memory and advanced
// by the kernels themselves, so the whole step sequence is CUDA-graph replayable and needs no
// "ready" handshake: a neighbour can only be two exchanges ahead of me after it received my
// contribution to the exchange in between, which I push only after my unpack of the previous
// one has completed (pass 2 starts after the ready counter).
// Numerics: identical to the stand-alone path (same bodies, b2_swe_body.cuh); corner halo cells
// come straight from the diagonal neighbours (see b2_halo.cu / tests/test_halo_equivalence.py).
#include <cstdio>
#include <cstring>

#include "b2_device.cuh"
#include "b2_runtime.h"

extern "C" void b2_set_error(const char* fmt, ...);
extern "C" void b2_count_launch(B2Comm* c);

#include "b2_halo_ll.cuh"
#include "b2_swe_body.cuh"

#define SWE_UNPACKERS 16

struct FusedArgs {
  B2SweParams p;
  float *h, *hn, *u, *v, *dh, *du, *dv, *fe, *fn, *q, *ke, *fe2, *fn2;
  int nb[FS_NSIDES];     // rank behind each side (-1 = none)
  int fstride;           // elements per field in a receive buffer
};

// ---- producer: push the edge cells this thread just computed --------------------------------
template <int NF>
__device__ __forceinline__ void fz_push(const B2DevComm& c, const FusedArgs& a, const unsigned* s_tx,
                                        const SweOut4& o, int j, int i0) {
  const int nx = a.p.nx, ny = a.p.ny, fs = a.fstride;
  if (!(i0 == 0 || (i0 <= nx - 2 && nx - 2 <= i0 + 3) || j == 1 || j == ny - 2)) return;   // no edge cell here
#pragma unroll
  for (int f = 0; f < NF; ++f) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int i = i0 + k;
      const float val = o.a[f][k];
      if (i == 1 && a.nb[FS_W] >= 0)
        fz_put(fz_buf(c, a.nb[FS_W], s_tx[FS_W] & 1u, FS_E) + f * fs + j, val, s_tx[FS_W] + 1u);
      if (i == nx - 2 && a.nb[FS_E] >= 0)
        fz_put(fz_buf(c, a.nb[FS_E], s_tx[FS_E] & 1u, FS_W) + f * fs + j, val, s_tx[FS_E] + 1u);
      if (j == 1) {
        if (i == 1 && a.nb[FS_SW] >= 0)
          fz_put(fz_buf(c, a.nb[FS_SW], s_tx[FS_SW] & 1u, FS_NE) + f, val, s_tx[FS_SW] + 1u);
        if (i == nx - 2 && a.nb[FS_SE] >= 0)
          fz_put(fz_buf(c, a.nb[FS_SE], s_tx[FS_SE] & 1u, FS_NW) + f, val, s_tx[FS_SE] + 1u);
      }
      if (j == ny - 2) {
        if (i == 1 && a.nb[FS_NW] >= 0)
          fz_put(fz_buf(c, a.nb[FS_NW], s_tx[FS_NW] & 1u, FS_SE) + f, val, s_tx[FS_NW] + 1u);
        if (i == nx - 2 && a.nb[FS_NE] >= 0)
          fz_put(fz_buf(c, a.nb[FS_NE], s_tx[FS_NE] & 1u, FS_SW) + f, val, s_tx[FS_NE] + 1u);
      }
    }
    // rows: four contiguous lanes -> two 16-byte stores when the whole group is interior
    if (j == 1 && a.nb[FS_S] >= 0) {
      uint2* dst = fz_buf(c, a.nb[FS_S], s_tx[FS_S] & 1u, FS_N) + f * fs + i0;
      const unsigned fl = s_tx[FS_S] + 1u;
      if (i0 >= 1 && i0 + 3 <= nx - 2) {
        fz_put2(dst, o.a[f][0], o.a[f][1], fl);
        fz_put2(dst + 2, o.a[f][2], o.a[f][3], fl);
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (i0 + k >= 1 && i0 + k <= nx - 2) fz_put(dst + k, o.a[f][k], fl);
      }
    }
    if (j == ny - 2 && a.nb[FS_N] >= 0) {
      uint2* dst = fz_buf(c, a.nb[FS_N], s_tx[FS_N] & 1u, FS_S) + f * fs + i0;
      const unsigned fl = s_tx[FS_N] + 1u;
      if (i0 >= 1 && i0 + 3 <= nx - 2) {
        fz_put2(dst, o.a[f][0], o.a[f][1], fl);
        fz_put2(dst + 2, o.a[f][2], o.a[f][3], fl);
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (i0 + k >= 1 && i0 + k <= nx - 2) fz_put(dst + k, o.a[f][k], fl);
      }
    }
  }
}

// ---- consumer: the first SWE_UNPACKERS CTAs poll the receive buffers and fill the halo cells ----
// FLD(f) must be a compile-time selectable expression -> macro keeps the field pointers in registers
template <int NF, typename GetField>
__device__ __forceinline__ void fz_unpack(const B2DevComm& c, const FusedArgs& a, const unsigned* s_rx,
                                          GetField fld, int nunp) {
  const int nx = a.p.nx, ny = a.p.ny, fs = a.fstride;
  const size_t P = (size_t)a.p.pitch;
  const int t = blockIdx.x * blockDim.x + threadIdx.x, nt = nunp * blockDim.x;
#pragma unroll
  for (int f = 0; f < NF; ++f) {
    float* const F = fld(f);
    if (a.nb[FS_W] >= 0) {
      const uint2* src = fz_buf(c, c.rank, s_rx[FS_W] & 1u, FS_W) + f * fs;
      const unsigned fl = s_rx[FS_W] + 1u;
      for (int j = 1 + t; j <= ny - 2; j += nt) F[(size_t)j * P] = fz_get(c, src + j, fl, FS_W);
    }
    if (a.nb[FS_E] >= 0) {
      const uint2* src = fz_buf(c, c.rank, s_rx[FS_E] & 1u, FS_E) + f * fs;
      const unsigned fl = s_rx[FS_E] + 1u;
      for (int j = 1 + t; j <= ny - 2; j += nt) F[(size_t)j * P + (nx - 1)] = fz_get(c, src + j, fl, FS_E);
    }
    if (a.nb[FS_S] >= 0) {
      const uint2* src = fz_buf(c, c.rank, s_rx[FS_S] & 1u, FS_S) + f * fs;
      const unsigned fl = s_rx[FS_S] + 1u;
      for (int i = 1 + t; i <= nx - 2; i += nt) F[i] = fz_get(c, src + i, fl, FS_S);
    }
    if (a.nb[FS_N] >= 0) {
      const uint2* src = fz_buf(c, c.rank, s_rx[FS_N] & 1u, FS_N) + f * fs;
      const unsigned fl = s_rx[FS_N] + 1u;
      for (int i = 1 + t; i <= nx - 2; i += nt) F[(size_t)(ny - 1) * P + i] = fz_get(c, src + i, fl, FS_N);
    }
    if (t == f) {     // corners of field f: one thread each
      if (a.nb[FS_SW] >= 0)
        F[0] = fz_get(c, fz_buf(c, c.rank, s_rx[FS_SW] & 1u, FS_SW) + f, s_rx[FS_SW] + 1u, FS_SW);
      if (a.nb[FS_SE] >= 0)
        F[nx - 1] = fz_get(c, fz_buf(c, c.rank, s_rx[FS_SE] & 1u, FS_SE) + f, s_rx[FS_SE] + 1u, FS_SE);
      if (a.nb[FS_NW] >= 0)
        F[(size_t)(ny - 1) * P] =
            fz_get(c, fz_buf(c, c.rank, s_rx[FS_NW] & 1u, FS_NW) + f, s_rx[FS_NW] + 1u, FS_NW);
      if (a.nb[FS_NE] >= 0)
        F[(size_t)(ny - 1) * P + (nx - 1)] =
            fz_get(c, fz_buf(c, c.rank, s_rx[FS_NE] & 1u, FS_NE) + f, s_rx[FS_NE] + 1u, FS_NE);
    }
  }
}

// one aligned group of 4 cells: stencil body (+ NVLink push of the edge cells it produced)
template <int KID>
__device__ __forceinline__ void fz_compute(const B2DevComm& c, const FusedArgs& a, const unsigned* s_tx,
                                           int j, int g) {
  const int i0 = g << 2, nx = a.p.nx;
  bool m[4];
  bool any = false;
#pragma unroll
  for (int k = 0; k < 4; ++k) { m[k] = (i0 + k >= 1) && (i0 + k <= nx - 2); any |= m[k]; }
  if (!any) return;
  [[maybe_unused]] SweOut4 o;
  if constexpr (KID == 1) {
    swe_k1_body(a.p, a.h, a.u, a.v, a.fe, a.fn, a.q, a.ke, j, i0, m, o);
    fz_push<4>(c, a, s_tx, o, j, i0);
  } else if constexpr (KID == 2) {
    swe_k2_body(a.p, a.h, a.hn, a.u, a.v, a.dh, a.du, a.dv, a.fe, a.fn, a.q, a.ke, j, i0, m, o);
    fz_push<3>(c, a, s_tx, o, j, i0);
  } else if constexpr (KID == 3) {
    swe_k3_body(a.p, a.u, a.fe, a.fn, j, i0, m, true, a.nb[FS_S] >= 0);
  } else if constexpr (KID == 4) {
    swe_k4_body(a.p, a.u, a.v, a.fe, a.fn, a.fe2, a.fn2, j, i0, m, o);
    fz_push<2>(c, a, s_tx, o, j, i0);
  } else {
    swe_k5_body(a.p, a.v, a.fe2, a.fn2, j, i0, m);
  }
}

// KID: 1 fluxes | 2 tendencies | 3 friction-u flux | 4 friction-u apply + friction-v flux | 5 friction-v
template <int KID>
__global__ void __launch_bounds__(SWE_THREADS) swe_fused(const B2DevComm c, const FusedArgs a) {
  constexpr bool HAS_IN = (KID == 2 || KID == 3 || KID == 5);
  constexpr bool HAS_OUT = (KID == 1 || KID == 2 || KID == 4);
  constexpr int NF_IN = KID == 2 ? 4 : KID == 3 ? 3 : 2;
  __shared__ unsigned s_rx[FS_NSIDES], s_tx[FS_NSIDES];
  if (threadIdx.x < FS_NSIDES) {
    s_rx[threadIdx.x] = b2_ld_volatile(c.ticket + TK_RX + threadIdx.x);
    s_tx[threadIdx.x] = b2_ld_volatile(c.ticket + TK_TX + threadIdx.x);
  }
  __syncthreads();
  const int nunp = (int)gridDim.x < SWE_UNPACKERS ? (int)gridDim.x : SWE_UNPACKERS;

  if (HAS_IN && (int)blockIdx.x < nunp) {
    if constexpr (KID == 2)
      fz_unpack<NF_IN>(c, a, s_rx, [&](int f) { return f == 0 ? a.fe : f == 1 ? a.fn : f == 2 ? a.q : a.ke; }, nunp);
    if constexpr (KID == 3)
      fz_unpack<NF_IN>(c, a, s_rx, [&](int f) { return f == 0 ? a.hn : f == 1 ? a.u : a.v; }, nunp);
    if constexpr (KID == 5)
      fz_unpack<NF_IN>(c, a, s_rx, [&](int f) { return f == 0 ? a.fe2 : a.fn2; }, nunp);
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) atomicAdd(c.ticket + TK_READY, 1u);
  }

  // Persistent CTAs: grid = min(#tiles, SMs x 8); every CTA strides over 256-group tiles so the
  // per-CTA protocol cost (counter loads, ready wait, finish atomic) is paid once, not per tile.
  const int nx = a.p.nx, ny = a.p.ny;
  const int ngroups = a.p.pitch >> 2;
  const long long total = (long long)(ny - 2) * ngroups;
  const int gA = (nx - 2) >> 2;                 // group that owns column nx-2 (and sees halo nx-1)
  // pass 1 (all cells when there is nothing to wait for; else the cells that touch no halo).
  // Tiles of 256 groups are handed out dynamically: CTAs that spent time unpacking simply take
  // fewer tiles, so the unpack never becomes the tail of the kernel.
  // The next tile is fetched (one atomic) while the current one is being computed, so the
  // scheduler latency is off the critical path.
  __shared__ unsigned s_next;
  constexpr int TILE_R = 4;                                   // 4 x 256 groups per tile
  const long long tile_groups = (long long)TILE_R * SWE_THREADS;
  const unsigned ntiles = (unsigned)((total + tile_groups - 1) / tile_groups);
  if (threadIdx.x == 0) s_next = atomicAdd(c.ticket + TK_TILE, 1u);
  __syncthreads();
  unsigned tile = s_next;
  while (tile < ntiles) {
    __syncthreads();                                          // everyone has read s_next
    unsigned prefetched = 0;
    if (threadIdx.x == 0) prefetched = atomicAdd(c.ticket + TK_TILE, 1u);
#pragma unroll
    for (int r = 0; r < TILE_R; ++r) {
      const unsigned idx = tile * (unsigned)tile_groups + (unsigned)r * SWE_THREADS + threadIdx.x;
      if (idx < (unsigned)total) {
        const unsigned rr = idx / (unsigned)ngroups;
        const int row = (int)rr + 1, g = (int)(idx - rr * (unsigned)ngroups);
        const bool isB = (g == 0 || g == gA || row == 1 || row == ny - 2);
        if (!(HAS_IN && isB)) fz_compute<KID>(c, a, s_tx, row, g);
      }
    }
    if (threadIdx.x == 0) s_next = prefetched;
    __syncthreads();
    tile = s_next;
  }
  if (HAS_IN) {
    if (threadIdx.x == 0) {                              // halos in place?
      const unsigned target = (unsigned)nunp;
      unsigned long long t0 = 0;
      unsigned spins = 0;
      while ((int)(b2_ld_volatile(c.ticket + TK_READY) - target) < 0) {
        if ((++spins & 0xfffu) == 0) {
          unsigned long long now = b2_gtime();
          if (t0 == 0) t0 = now;
          else if (now - t0 > c.timeout_ns) b2_fatal(c, B2_ERR_TIMEOUT, B2_OPC_HALO, -1, target, 0, 6);
        }
      }
      __threadfence();
    }
    __syncthreads();
    // pass 2: dense sweep over the groups next to the boundary --
    //   [0, 2(ny-2))            : rows 1..ny-2, groups {0, gA}
    //   [2(ny-2), +2(ngroups-2)): rows 1 and ny-2, every other group
    const int ncol = 2 * (ny - 2), nrow_items = ngroups - 2;
    const int nB = ncol + 2 * nrow_items;
    for (int b = blockIdx.x * SWE_THREADS + threadIdx.x; b < nB; b += gridDim.x * SWE_THREADS) {
      int row, g;
      if (b < ncol) {
        row = 1 + (b >> 1);
        g = (b & 1) ? gA : 0;
      } else {
        const int r = b - ncol;
        row = (r < nrow_items) ? 1 : ny - 2;
        g = (r < nrow_items ? r : r - nrow_items) + 1;
        if (g >= gA) ++g;
      }
      fz_compute<KID>(c, a, s_tx, row, g);
    }
  }

  // last CTA to finish advances the message counters
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned old = atomicAdd(c.ticket + TK_FIN, 1u);
    if (old == gridDim.x - 1) {
      b2_st_volatile(c.ticket + TK_FIN, 0u);
      b2_st_volatile(c.ticket + TK_TILE, 0u);
      for (int k = 0; k < FS_NSIDES; ++k)
        if (a.nb[k] >= 0) {
          if (HAS_IN) b2_st_volatile(c.ticket + TK_RX + k, s_rx[k] + 1u);
          if (HAS_OUT) b2_st_volatile(c.ticket + TK_TX + k, s_tx[k] + 1u);
        }
      if (HAS_IN) b2_st_volatile(c.ticket + TK_READY, 0u);
      __threadfence();
    }
  }
}

template <int KID>
static unsigned fused_blocks(const B2Comm* c, const B2SweParams& p) {
  static int occ = 0;                                     // resident CTAs per SM for this kernel
  if (occ == 0) {
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, swe_fused<KID>, SWE_THREADS, 0) != cudaSuccess ||
        occ < 1)
      occ = 1;
  }
  const long long n = (long long)(p.ny - 2) * (p.pitch / 4);
  const long long tiles = (n + SWE_THREADS - 1) / SWE_THREADS;
  const long long cap = (long long)c->sm_count * occ;    // persistent: exactly one resident wave
  return (unsigned)(tiles < cap ? tiles : cap);
}

extern "C" int b2_swe_multistep(B2Comm* c, const B2SweParams* p0, const B2SweState* st,
                                const B2HaloDesc* topo, int nsteps, int first_step, cudaStream_t s);

// Same contract as b2_swe_multistep (b2_swe.cu), fused kernels.  Falls back to the stand-alone
// path when there is no lateral friction (the K3'..K5' chain that carries two of the exchanges).
extern "C" int b2_swe_multistep_fused(B2Comm* c, const B2SweParams* p0, const B2SweState* st,
                                      const B2HaloDesc* topo, int nsteps, int first_step,
                                      cudaStream_t s) {
  if (!(p0->viscosity > 0.f)) return b2_swe_multistep(c, p0, st, topo, nsteps, first_step, s);
  if ((long long)p0->ny * (p0->pitch / 4) >= (1ll << 31)) {
    b2_set_error("swe(fused): grid too large for 32-bit group indices");
    return B2_ERR_BAD_ARG;
  }
  if (p0->pitch % 4 != 0 || p0->pitch < p0->nx || p0->ny < 4 || p0->nx < 10) {
    b2_set_error("swe(fused): need ny >= 4, nx >= 10 and a row pitch that is a multiple of 4 and >= nx");
    return B2_ERR_BAD_ARG;
  }
  FusedArgs a;
  memset(&a, 0, sizeof a);
  a.p = *p0;
  const int mx = p0->ny > p0->nx ? p0->ny : p0->nx;
  a.fstride = (mx + 3) / 4 * 4;
  if ((size_t)4 * a.fstride * sizeof(uint2) > c->dev.lay.halo_ll_cap) {
    b2_set_error("swe(fused): halo buffers too small (%zu > %zu); raise MPI4JAX_B200_HALO_BYTES",
                 (size_t)4 * a.fstride * sizeof(uint2), c->dev.lay.halo_ll_cap);
    return B2_ERR_BAD_ARG;
  }
  const int nb[FS_NSIDES] = {topo->west, topo->east, topo->south, topo->north,
                             topo->sw, topo->se, topo->nw, topo->ne};
  for (int k = 0; k < FS_NSIDES; ++k) {
    if (nb[k] < -1 || nb[k] >= c->dev.size) {
      b2_set_error("swe(fused): invalid neighbour rank %d", nb[k]);
      return B2_ERR_BAD_ARG;
    }
    a.nb[k] = nb[k];
  }
  a.u = st->u; a.v = st->v; a.dh = st->dh; a.du = st->du; a.dv = st->dv;
  a.fe = st->fe; a.fn = st->fn; a.q = st->q; a.ke = st->ke; a.fe2 = st->fe2; a.fn2 = st->fn2;
  float* h = st->h0;
  float* hn = st->h1;
  const unsigned g1 = fused_blocks<1>(c, a.p), g2 = fused_blocks<2>(c, a.p), g3 = fused_blocks<3>(c, a.p),
                 g4 = fused_blocks<4>(c, a.p), g5 = fused_blocks<5>(c, a.p);
  cudaError_t err = cudaSuccess;
  for (int it = 0; it < nsteps; ++it) {
    a.p.first_step = (first_step && it == 0) ? 1 : 0;
    a.h = h;
    a.hn = hn;
    swe_fused<1><<<g1, SWE_THREADS, 0, s>>>(c->dev, a);
    swe_fused<2><<<g2, SWE_THREADS, 0, s>>>(c->dev, a);
    swe_fused<3><<<g3, SWE_THREADS, 0, s>>>(c->dev, a);
    swe_fused<4><<<g4, SWE_THREADS, 0, s>>>(c->dev, a);
    swe_fused<5><<<g5, SWE_THREADS, 0, s>>>(c->dev, a);
    for (int k = 0; k < 5; ++k) b2_count_launch(c);
    if ((err = cudaGetLastError()) != cudaSuccess) break;
    float* t = h; h = hn; hn = t;
  }
  if (err == cudaSuccess && h != st->h0)
    err = cudaMemcpyAsync(st->h0, h, (size_t)a.p.ny * a.p.pitch * sizeof(float),
                          cudaMemcpyDeviceToDevice, s);
  if (err != cudaSuccess) {
    b2_set_error("swe_multistep_fused: launch failed: %s", cudaGetErrorString(err));
    return 1000 + (int)err;
  }
  return 0;
}
