// mpi4jax_b200 -- shallow-water stencil kernels (fp32), the compute side of the
// reference's demo application (examples/shallow_water.py:270-403).  The
// reference leaves this arithmetic to XLA (dozens of fused elementwise kernels
// plus pad / dynamic-update-slice copies per step, separated by 48 blocking MPI
// custom calls).  Here one model step is five stencil launches and four fused
// halo exchanges (b2_halo.cu):
//
//   K1 fluxes      (h,u,v)                 -> fe, fn, q, ke        [exchange fe,fn,q,ke]
//   K2 tendencies  (h,fe,fn,q,ke,d*_old)   -> dh,du,dv, h',u,v     [exchange h',u,v]
//   K3 friction-u flux   (u)               -> fe, fn               [exchange fe,fn]
//   K4 friction-u apply + friction-v flux  -> u, fe2, fn2          [exchange fe2,fn2]
//   K5 friction-v apply                    -> v
//
// The physics, including the order of boundary updates, the edge-padded `hc`
// rows at the y walls and the `v - u` term of the second friction flux
// (shallow_water.py:387-392), follows the reference line by line so that the
// two programs integrate the same discrete system.  The redundant `hc` halo
// exchange of the reference (:278-279) is elided: after `enforce_boundaries(h)`
// at the end of the previous step, h's halo already holds exactly the values
// that exchange would deliver ("shallow water").
#include <cstdio>

#include "b2_runtime.h"

extern "C" void b2_set_error(const char* fmt, ...);
extern "C" void b2_count_launch(B2Comm* c);

struct B2SweParams {
  int ny, nx;
  float dx, dy, dt, gravity, viscosity;
  float ab_a, ab_b;          // Adams-Bashforth weights
  int first_step;
  int south_wall, north_wall; // this rank touches the y walls (hc edge padding)
  const float* coriolis;      // [ny]
};

#define SWE_BX 128
#define SWE_BY 2

__device__ __forceinline__ float hc_at(const float* __restrict__ h, const B2SweParams& p, int j, int i) {
  // hc = h with the physical wall rows replaced by their neighbouring interior row
  if (p.south_wall && j == 0) j = 1;
  if (p.north_wall && j == p.ny - 1) j = p.ny - 2;
  return h[(size_t)j * p.nx + i];
}

__global__ void __launch_bounds__(SWE_BX * SWE_BY)
swe_k1_fluxes(B2SweParams p, const float* __restrict__ h, const float* __restrict__ u,
              const float* __restrict__ v, float* __restrict__ fe, float* __restrict__ fn,
              float* __restrict__ q, float* __restrict__ ke) {
  const int i = blockIdx.x * SWE_BX + threadIdx.x + 1;
  const int j = blockIdx.y * SWE_BY + threadIdx.y + 1;
  if (i >= p.nx - 1 || j >= p.ny - 1) return;
  const size_t c = (size_t)j * p.nx + i;
  const float hc = hc_at(h, p, j, i), hce = hc_at(h, p, j, i + 1), hcn = hc_at(h, p, j + 1, i),
              hcne = hc_at(h, p, j + 1, i + 1);
  const float uc = u[c], vc = v[c];
  fe[c] = 0.5f * (hc + hce) * uc;
  fn[c] = 0.5f * (hc + hcn) * vc;
  const float rel = (v[c + 1] - vc) / p.dx - (u[c + p.nx] - uc) / p.dy;
  q[c] = (p.coriolis[j] + rel) * (1.0f / (0.25f * (hc + hce + hcn + hcne)));
  const float uw = u[c - 1], vs = v[c - p.nx];
  ke[c] = 0.5f * (0.5f * (uc * uc + uw * uw) + 0.5f * (vc * vc + vs * vs));
}

__global__ void __launch_bounds__(SWE_BX * SWE_BY)
swe_k2_tendencies(B2SweParams p, const float* __restrict__ h, float* __restrict__ h_new,
                  float* __restrict__ u, float* __restrict__ v, float* __restrict__ dh,
                  float* __restrict__ du, float* __restrict__ dv, const float* __restrict__ fe,
                  const float* __restrict__ fn, const float* __restrict__ q,
                  const float* __restrict__ ke) {
  const int i = blockIdx.x * SWE_BX + threadIdx.x;
  const int j = blockIdx.y * SWE_BY + threadIdx.y;
  if (i >= p.nx || j >= p.ny) return;
  const size_t c = (size_t)j * p.nx + i;
  const int nx = p.nx;
  if (i == 0 || j == 0 || i == p.nx - 1 || j == p.ny - 1) {
    h_new[c] = h[c];   // halo cells are carried over; the exchange that follows refreshes them
    return;
  }
  const float fec = fe[c], fnc = fn[c], qc = q[c], kec = ke[c], hcv = h[c];
  const float dh_new = -(fec - fe[c - 1]) / p.dx - (fnc - fn[c - nx]) / p.dy;
  float du_new = -p.gravity * (h[c + 1] - hcv) / p.dx +
                 0.5f * (qc * 0.5f * (fnc + fn[c + 1]) + q[c - nx] * 0.5f * (fn[c - nx] + fn[c - nx + 1]));
  float dv_new = -p.gravity * (h[c + nx] - hcv) / p.dy -
                 0.5f * (qc * 0.5f * (fec + fe[c + nx]) + q[c - 1] * 0.5f * (fe[c - 1] + fe[c + nx - 1]));
  du_new += -(ke[c + 1] - kec) / p.dx;
  dv_new += -(ke[c + nx] - kec) / p.dy;
  if (p.first_step) {
    u[c] += p.dt * du_new;
    v[c] += p.dt * dv_new;
    h_new[c] = hcv + p.dt * dh_new;
  } else {
    u[c] += p.dt * (p.ab_a * du_new + p.ab_b * du[c]);
    v[c] += p.dt * (p.ab_a * dv_new + p.ab_b * dv[c]);
    h_new[c] = hcv + p.dt * (p.ab_a * dh_new + p.ab_b * dh[c]);
  }
  dh[c] = dh_new;
  du[c] = du_new;
  dv[c] = dv_new;
}

__global__ void __launch_bounds__(SWE_BX * SWE_BY)
swe_k3_friction_flux_u(B2SweParams p, const float* __restrict__ u, float* __restrict__ fe,
                       float* __restrict__ fn) {
  const int i = blockIdx.x * SWE_BX + threadIdx.x + 1;
  const int j = blockIdx.y * SWE_BY + threadIdx.y + 1;
  if (i >= p.nx - 1 || j >= p.ny - 1) return;
  const size_t c = (size_t)j * p.nx + i;
  const float uc = u[c];
  fe[c] = p.viscosity * (u[c + 1] - uc) / p.dx;
  fn[c] = p.viscosity * (u[c + p.nx] - uc) / p.dy;
}

__global__ void __launch_bounds__(SWE_BX * SWE_BY)
swe_k4_friction_u_flux_v(B2SweParams p, float* __restrict__ u, const float* __restrict__ v,
                         const float* __restrict__ fe, const float* __restrict__ fn,
                         float* __restrict__ fe2, float* __restrict__ fn2) {
  const int i = blockIdx.x * SWE_BX + threadIdx.x + 1;
  const int j = blockIdx.y * SWE_BY + threadIdx.y + 1;
  if (i >= p.nx - 1 || j >= p.ny - 1) return;
  const size_t c = (size_t)j * p.nx + i;
  const float un = u[c] + p.dt * ((fe[c] - fe[c - 1]) / p.dx + (fn[c] - fn[c - p.nx]) / p.dy);
  u[c] = un;
  // NOTE: `v - u` mirrors the reference (examples/shallow_water.py:387-392)
  fe2[c] = p.viscosity * (v[c + 1] - un) / p.dx;
  fn2[c] = p.viscosity * (v[c + p.nx] - un) / p.dy;
}

__global__ void __launch_bounds__(SWE_BX * SWE_BY)
swe_k5_friction_v(B2SweParams p, float* __restrict__ v, const float* __restrict__ fe2,
                  const float* __restrict__ fn2) {
  const int i = blockIdx.x * SWE_BX + threadIdx.x + 1;
  const int j = blockIdx.y * SWE_BY + threadIdx.y + 1;
  if (i >= p.nx - 1 || j >= p.ny - 1) return;
  const size_t c = (size_t)j * p.nx + i;
  v[c] += p.dt * ((fe2[c] - fe2[c - 1]) / p.dx + (fn2[c] - fn2[c - p.nx]) / p.dy);
}

static dim3 swe_grid(int nx, int ny) {
  return dim3((unsigned)((nx + SWE_BX - 1) / SWE_BX), (unsigned)((ny + SWE_BY - 1) / SWE_BY));
}

static int swe_done(B2Comm* c, const char* name) {
  b2_count_launch(c);
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) {
    b2_set_error("%s: kernel launch failed: %s", name, cudaGetErrorString(err));
    return 1000 + (int)err;
  }
  return 0;
}

extern "C" {

int b2_swe_fluxes(B2Comm* c, const B2SweParams* p, const float* h, const float* u, const float* v,
                  float* fe, float* fn, float* q, float* ke, cudaStream_t s) {
  swe_k1_fluxes<<<swe_grid(p->nx, p->ny), dim3(SWE_BX, SWE_BY), 0, s>>>(*p, h, u, v, fe, fn, q, ke);
  return swe_done(c, "swe_fluxes");
}

int b2_swe_tendencies(B2Comm* c, const B2SweParams* p, const float* h, float* h_new, float* u,
                      float* v, float* dh, float* du, float* dv, const float* fe, const float* fn,
                      const float* q, const float* ke, cudaStream_t s) {
  swe_k2_tendencies<<<swe_grid(p->nx, p->ny), dim3(SWE_BX, SWE_BY), 0, s>>>(*p, h, h_new, u, v, dh,
                                                                           du, dv, fe, fn, q, ke);
  return swe_done(c, "swe_tendencies");
}

int b2_swe_friction_flux_u(B2Comm* c, const B2SweParams* p, const float* u, float* fe, float* fn,
                           cudaStream_t s) {
  swe_k3_friction_flux_u<<<swe_grid(p->nx, p->ny), dim3(SWE_BX, SWE_BY), 0, s>>>(*p, u, fe, fn);
  return swe_done(c, "swe_friction_flux_u");
}

int b2_swe_friction_u_flux_v(B2Comm* c, const B2SweParams* p, float* u, const float* v,
                             const float* fe, const float* fn, float* fe2, float* fn2,
                             cudaStream_t s) {
  swe_k4_friction_u_flux_v<<<swe_grid(p->nx, p->ny), dim3(SWE_BX, SWE_BY), 0, s>>>(*p, u, v, fe, fn,
                                                                                  fe2, fn2);
  return swe_done(c, "swe_friction_u_flux_v");
}

int b2_swe_friction_v(B2Comm* c, const B2SweParams* p, float* v, const float* fe2, const float* fn2,
                      cudaStream_t s) {
  swe_k5_friction_v<<<swe_grid(p->nx, p->ny), dim3(SWE_BX, SWE_BY), 0, s>>>(*p, v, fe2, fn2);
  return swe_done(c, "swe_friction_v");
}

// One call = `nsteps` model steps (5 stencil launches + 4 fused halo exchanges each), all
// enqueued on `s` without touching the host again: the whole time loop of the reference's
// `do_multistep` (examples/shallow_water.py:406-411) becomes a launch sequence that CUDA-graph
// capture turns into a single replayable graph.  `h0`/`h1` ping-pong; the state is returned in
// h0 (an odd step count ends with one device copy).
struct B2SweState {
  float *h0, *h1, *u, *v, *dh, *du, *dv, *fe, *fn, *q, *ke, *fe2, *fn2;
};

int b2_swe_multistep(B2Comm* c, const B2SweParams* p0, const B2SweState* st, const B2HaloDesc* topo,
                     int nsteps, int first_step, cudaStream_t s) {
  B2SweParams p = *p0;
  float* h = st->h0;
  float* hn = st->h1;
  B2HaloDesc d = *topo;
  d.ny = p.ny;
  d.nx = p.nx;
  int rc = 0;
  for (int it = 0; it < nsteps && rc == 0; ++it) {
    p.first_step = (first_step && it == 0) ? 1 : 0;
    if ((rc = b2_swe_fluxes(c, &p, h, st->u, st->v, st->fe, st->fn, st->q, st->ke, s))) break;
    d.nfields = 4;
    d.field[0] = st->fe; d.kind[0] = 1;
    d.field[1] = st->fn; d.kind[1] = 2;
    d.field[2] = st->q;  d.kind[2] = 0;
    d.field[3] = st->ke; d.kind[3] = 0;
    if ((rc = b2_halo_exchange(c, &d, s))) break;
    if ((rc = b2_swe_tendencies(c, &p, h, hn, st->u, st->v, st->dh, st->du, st->dv, st->fe, st->fn,
                                st->q, st->ke, s))) break;
    d.nfields = 3;
    d.field[0] = hn;    d.kind[0] = 0;
    d.field[1] = st->u; d.kind[1] = 1;
    d.field[2] = st->v; d.kind[2] = 2;
    if ((rc = b2_halo_exchange(c, &d, s))) break;
    if (p.viscosity > 0.f) {
      if ((rc = b2_swe_friction_flux_u(c, &p, st->u, st->fe, st->fn, s))) break;
      d.nfields = 2;
      d.field[0] = st->fe; d.kind[0] = 1;
      d.field[1] = st->fn; d.kind[1] = 2;
      if ((rc = b2_halo_exchange(c, &d, s))) break;
      if ((rc = b2_swe_friction_u_flux_v(c, &p, st->u, st->v, st->fe, st->fn, st->fe2, st->fn2, s))) break;
      d.field[0] = st->fe2;
      d.field[1] = st->fn2;
      if ((rc = b2_halo_exchange(c, &d, s))) break;
      if ((rc = b2_swe_friction_v(c, &p, st->v, st->fe2, st->fn2, s))) break;
    }
    float* t = h; h = hn; hn = t;
  }
  if (rc == 0 && h != st->h0) {
    cudaError_t e = cudaMemcpyAsync(st->h0, h, (size_t)p.ny * p.nx * sizeof(float),
                                    cudaMemcpyDeviceToDevice, s);
    if (e != cudaSuccess) { b2_set_error("swe_multistep: copy failed: %s", cudaGetErrorString(e)); rc = 1000 + (int)e; }
  }
  return rc;
}

}  // extern "C"
