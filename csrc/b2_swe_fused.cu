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

#include "b2_swe_body.cuh"

#define SWE_UNPACKERS 8
enum { FS_W = 0, FS_E, FS_S, FS_N, FS_SW, FS_SE, FS_NW, FS_NE, FS_NSIDES };

// ticket words used by the fused path (local device memory)
#define TK_READY 5     // += 1 per unpacker CTA, reset to 0 by the last CTA of the kernel
#define TK_FIN 6       // finish counter
#define TK_RX 32       // [8] messages received per side
#define TK_TX 40       // [8] messages sent per side

struct B2SweState {
  float *h0, *h1, *u, *v, *dh, *du, *dv, *fe, *fn, *q, *ke, *fe2, *fn2;
};

struct FusedArgs {
  B2SweParams p;
  float *h, *hn, *u, *v, *dh, *du, *dv, *fe, *fn, *q, *ke, *fe2, *fn2;
  int nb[FS_NSIDES];     // rank behind each side (-1 = none)
  int fstride;           // elements per field in a receive buffer
};

__device__ __forceinline__ uint2* fz_buf(const B2DevComm& c, int rank, unsigned parity, int side) {
  return (uint2*)(c.heap[rank] + c.lay.halo_ll_off +
                  ((size_t)parity * FS_NSIDES + side) * c.lay.halo_ll_cap);
}
__device__ __forceinline__ void fz_put(uint2* p, float v, unsigned flag) {
  asm volatile("st.relaxed.sys.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(__float_as_uint(v)),
               "r"(flag) : "memory");
}
__device__ __forceinline__ void fz_put2(uint2* p, float v0, float v1, unsigned flag) {
  asm volatile("st.relaxed.sys.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p),
               "r"(__float_as_uint(v0)), "r"(flag), "r"(__float_as_uint(v1)), "r"(flag) : "memory");
}
__device__ __forceinline__ float fz_get(const B2DevComm& c, const uint2* p, unsigned flag, int side) {
  unsigned v, f;
  unsigned long long t0 = 0;
  unsigned spins = 0;
  while (true) {
    asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(v), "=r"(f) : "l"(p) : "memory");
    if (f == flag) break;
    if ((++spins & 0xfffu) == 0) {
      unsigned long long now = b2_gtime();
      if (t0 == 0) t0 = now;
      else if (now - t0 > c.timeout_ns) b2_fatal(c, B2_ERR_TIMEOUT, B2_OPC_HALO, side, flag, f, 5);
    }
  }
  return __uint_as_float(v);
}

// my message towards side k lands on the opposite side of the neighbour
__device__ __constant__ int kOpp[FS_NSIDES] = {FS_E, FS_W, FS_N, FS_S, FS_NE, FS_NW, FS_SE, FS_SW};

// ---- producer: push the edge cells this thread just computed --------------------------------
template <int NF>
__device__ __forceinline__ void fz_push(const B2DevComm& c, const FusedArgs& a, const unsigned* s_tx,
                                        const SweOut4& o, int j, int i0) {
  const int nx = a.p.nx, ny = a.p.ny, fs = a.fstride;
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
template <int NF>
__device__ __forceinline__ void fz_unpack(const B2DevComm& c, const FusedArgs& a, const unsigned* s_rx,
                                          float* const (&fld)[4], int nunp) {
  const int nx = a.p.nx, ny = a.p.ny, fs = a.fstride;
  const size_t P = (size_t)a.p.pitch;
  const int t = blockIdx.x * blockDim.x + threadIdx.x, nt = nunp * blockDim.x;
  if (a.nb[FS_W] >= 0) {
    const uint2* src = fz_buf(c, c.rank, s_rx[FS_W] & 1u, FS_W);
    const unsigned fl = s_rx[FS_W] + 1u;
    for (int k = t; k < NF * (ny - 2); k += nt) {
      const int f = k / (ny - 2), j = 1 + k - f * (ny - 2);
      fld[f][(size_t)j * P] = fz_get(c, src + f * fs + j, fl, FS_W);
    }
  }
  if (a.nb[FS_E] >= 0) {
    const uint2* src = fz_buf(c, c.rank, s_rx[FS_E] & 1u, FS_E);
    const unsigned fl = s_rx[FS_E] + 1u;
    for (int k = t; k < NF * (ny - 2); k += nt) {
      const int f = k / (ny - 2), j = 1 + k - f * (ny - 2);
      fld[f][(size_t)j * P + (nx - 1)] = fz_get(c, src + f * fs + j, fl, FS_E);
    }
  }
  if (a.nb[FS_S] >= 0) {
    const uint2* src = fz_buf(c, c.rank, s_rx[FS_S] & 1u, FS_S);
    const unsigned fl = s_rx[FS_S] + 1u;
    for (int k = t; k < NF * (nx - 2); k += nt) {
      const int f = k / (nx - 2), i = 1 + k - f * (nx - 2);
      fld[f][i] = fz_get(c, src + f * fs + i, fl, FS_S);
    }
  }
  if (a.nb[FS_N] >= 0) {
    const uint2* src = fz_buf(c, c.rank, s_rx[FS_N] & 1u, FS_N);
    const unsigned fl = s_rx[FS_N] + 1u;
    for (int k = t; k < NF * (nx - 2); k += nt) {
      const int f = k / (nx - 2), i = 1 + k - f * (nx - 2);
      fld[f][(size_t)(ny - 1) * P + i] = fz_get(c, src + f * fs + i, fl, FS_N);
    }
  }
  if (t < NF) {
    const int f = t;
    if (a.nb[FS_SW] >= 0)
      fld[f][0] = fz_get(c, fz_buf(c, c.rank, s_rx[FS_SW] & 1u, FS_SW) + f, s_rx[FS_SW] + 1u, FS_SW);
    if (a.nb[FS_SE] >= 0)
      fld[f][nx - 1] = fz_get(c, fz_buf(c, c.rank, s_rx[FS_SE] & 1u, FS_SE) + f, s_rx[FS_SE] + 1u, FS_SE);
    if (a.nb[FS_NW] >= 0)
      fld[f][(size_t)(ny - 1) * P] =
          fz_get(c, fz_buf(c, c.rank, s_rx[FS_NW] & 1u, FS_NW) + f, s_rx[FS_NW] + 1u, FS_NW);
    if (a.nb[FS_NE] >= 0)
      fld[f][(size_t)(ny - 1) * P + (nx - 1)] =
          fz_get(c, fz_buf(c, c.rank, s_rx[FS_NE] & 1u, FS_NE) + f, s_rx[FS_NE] + 1u, FS_NE);
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
    if constexpr (KID == 2) { float* const fld[4] = {a.fe, a.fn, a.q, a.ke}; fz_unpack<NF_IN>(c, a, s_rx, fld, nunp); }
    if constexpr (KID == 3) { float* const fld[4] = {a.hn, a.u, a.v, nullptr}; fz_unpack<NF_IN>(c, a, s_rx, fld, nunp); }
    if constexpr (KID == 5) { float* const fld[4] = {a.fe2, a.fn2, nullptr, nullptr}; fz_unpack<NF_IN>(c, a, s_rx, fld, nunp); }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) atomicAdd(c.ticket + TK_READY, 1u);
  }

  int j = 0, i0 = 0;
  bool m[4];
  const bool active = swe_map(a.p, j, i0, m);
  // groups whose stencil touches a halo cell (or that own an edge cell)
  const int nx = a.p.nx, ny = a.p.ny;
  const bool isB = active && (i0 == 0 || (i0 <= nx - 1 && i0 + 4 >= nx - 2) || j == 1 || j == ny - 2);

  auto compute = [&]() {
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
  };

  if (!HAS_IN) {
    if (active) compute();
  } else {
    if (active && !isB) compute();                       // pass 1: overlaps the exchange
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
    if (isB) compute();                                  // pass 2: cells next to the boundary
  }

  // last CTA to finish advances the message counters
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned old = atomicAdd(c.ticket + TK_FIN, 1u);
    if (old == gridDim.x - 1) {
      b2_st_volatile(c.ticket + TK_FIN, 0u);
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

static unsigned fused_blocks(const B2SweParams& p) {
  const long long n = (long long)(p.ny - 2) * (p.pitch / 4);
  return (unsigned)((n + SWE_THREADS - 1) / SWE_THREADS);
}

extern "C" int b2_swe_multistep(B2Comm* c, const B2SweParams* p0, const B2SweState* st,
                                const B2HaloDesc* topo, int nsteps, int first_step, cudaStream_t s);

// Same contract as b2_swe_multistep (b2_swe.cu), fused kernels.  Falls back to the stand-alone
// path when there is no lateral friction (the K3'..K5' chain that carries two of the exchanges).
extern "C" int b2_swe_multistep_fused(B2Comm* c, const B2SweParams* p0, const B2SweState* st,
                                      const B2HaloDesc* topo, int nsteps, int first_step,
                                      cudaStream_t s) {
  if (!(p0->viscosity > 0.f)) return b2_swe_multistep(c, p0, st, topo, nsteps, first_step, s);
  if (p0->pitch % 4 != 0 || p0->pitch < p0->nx || p0->ny < 4 || p0->nx < 4) {
    b2_set_error("swe(fused): need ny,nx >= 4 and a row pitch that is a multiple of 4 and >= nx");
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
  const unsigned grid = fused_blocks(a.p);
  cudaError_t err = cudaSuccess;
  for (int it = 0; it < nsteps; ++it) {
    a.p.first_step = (first_step && it == 0) ? 1 : 0;
    a.h = h;
    a.hn = hn;
    swe_fused<1><<<grid, SWE_THREADS, 0, s>>>(c->dev, a);
    swe_fused<2><<<grid, SWE_THREADS, 0, s>>>(c->dev, a);
    swe_fused<3><<<grid, SWE_THREADS, 0, s>>>(c->dev, a);
    swe_fused<4><<<grid, SWE_THREADS, 0, s>>>(c->dev, a);
    swe_fused<5><<<grid, SWE_THREADS, 0, s>>>(c->dev, a);
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
