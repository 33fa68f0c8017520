// mpi4jax_b200 -- shallow water with the flux and tendency kernels fused (EXPERIMENTAL, opt-in:
// MPI4JAX_B200_SWE_K12=1 / ShallowWaterModel(k12=True)).  See b2_swe_k12_body.cuh for the idea and
// the bulk / frame split; tests/test_swe_host_emulation.py checks the indexing of these bodies on
// the host.  A step is
//
//   K12 (bulk)  |  K1 (frame, width 2) -> exchange(fe, fn, q, ke) -> K2 (ring)
//   exchange(h', u', v')  ->  K34 (u' -> u)  ->  exchange(fe2, fn2)  ->  K5 (v' -> v)
//
// 21 array passes instead of 32; u and v return to their home buffers within the step, h
// ping-pongs across steps as in b2_swe_multistep.
#include <cstdio>

#include "b2_runtime.h"
#include "b2_swe_k12_body.cuh"
#include "b2_launch.cuh"

extern "C" void b2_set_error(const char* fmt, ...);
extern "C" void b2_count_launch(B2Comm* c);
extern "C" int b2_swe_multistep(B2Comm* c, const B2SweParams* p0, const B2SweState* st,
                                const B2HaloDesc* topo, int nsteps, int first_step, cudaStream_t s);
extern "C" int b2_swe_friction_u_fused(B2Comm* c, const B2SweParams* p, const float* u, float* u_new,
                                       const float* v, float* fe2, float* fn2, int has_south,
                                       cudaStream_t s);

__global__ void __launch_bounds__(SWE_THREADS, 2)
swe_k12_bulk(B2SweParams p, const float* __restrict__ h, float* __restrict__ h_new,
             const float* __restrict__ u, float* __restrict__ u_new, const float* __restrict__ v,
             float* __restrict__ v_new, float* __restrict__ dh, float* __restrict__ du,
             float* __restrict__ dv) {
  b2_pdl_enter();
  int j, i0;
  bool m[4];
  if (!swe_map(p, j, i0, m)) return;
  if (j < 2 || j > p.ny - 3) return;           // rows 1 and ny-2 belong to the ring kernel entirely
  swe_k12_body(p, h, h_new, u, u_new, v, v_new, dh, du, dv, j, i0, m);
}

__global__ void __launch_bounds__(SWE_THREADS)
swe_k1_frame(B2SweParams p, SweFrame f, const float* __restrict__ h, const float* __restrict__ u,
             const float* __restrict__ v, float* __restrict__ fe, float* __restrict__ fn,
             float* __restrict__ q, float* __restrict__ ke) {
  b2_pdl_enter();
  int j, i0;
  bool m[4];
  if (!swe_frame_task(p, f, (long long)blockIdx.x * SWE_THREADS + threadIdx.x, j, i0, m)) return;
  if (!(m[0] || m[1] || m[2] || m[3])) return;
  SweOut4 o;
  swe_k1_body(p, h, u, v, fe, fn, q, ke, j, i0, m, o);
}

__global__ void __launch_bounds__(SWE_THREADS)
swe_k2_ring(B2SweParams p, SweFrame f, const float* __restrict__ h, float* __restrict__ h_new,
            const float* __restrict__ u, float* __restrict__ u_new, const float* __restrict__ v,
            float* __restrict__ v_new, float* __restrict__ dh, float* __restrict__ du,
            float* __restrict__ dv, const float* __restrict__ fe, const float* __restrict__ fn,
            const float* __restrict__ q, const float* __restrict__ ke) {
  b2_pdl_enter();
  int j, i0;
  bool m[4];
  if (!swe_frame_task(p, f, (long long)blockIdx.x * SWE_THREADS + threadIdx.x, j, i0, m)) return;
  if (!(m[0] || m[1] || m[2] || m[3])) return;
  swe_k2_ring_body(p, h, h_new, u, u_new, v, v_new, dh, du, dv, fe, fn, q, ke, j, i0, m);
}

__global__ void __launch_bounds__(SWE_THREADS)
swe_k5_pp(B2SweParams p, const float* __restrict__ v, float* __restrict__ v_new,
          const float* __restrict__ fe2, const float* __restrict__ fn2) {
  b2_pdl_enter();
  int j, i0;
  bool m[4];
  if (!swe_map(p, j, i0, m)) return;
  swe_k5_pp_body(p, v, v_new, fe2, fn2, j, i0, m);
}

__global__ void __launch_bounds__(SWE_THREADS, 2)
swe_k345_bulk(B2SweParams p, const float* __restrict__ u, float* __restrict__ u_new,
              const float* __restrict__ v, float* __restrict__ v_new, int has_south) {
  b2_pdl_enter();
  int j, i0;
  bool m[4];
  if (!swe_map(p, j, i0, m)) return;
  if (j < 2 || j > p.ny - 3) return;
  swe_k345_body(p, u, u_new, v, v_new, j, i0, has_south != 0);
}

__global__ void __launch_bounds__(SWE_THREADS)
swe_k34_frame(B2SweParams p, SweFrame f, const float* __restrict__ u, float* __restrict__ u_new,
              const float* __restrict__ v, float* __restrict__ fe2, float* __restrict__ fn2,
              int has_south) {
  b2_pdl_enter();
  int j, i0;
  bool m[4];
  if (!swe_frame_task(p, f, (long long)blockIdx.x * SWE_THREADS + threadIdx.x, j, i0, m)) return;
  if (!(m[0] || m[1] || m[2] || m[3])) return;
  SweOut4 o;
  swe_k34_body(p, u, u_new, v, fe2, fn2, j, i0, m, has_south != 0, o);
}

__global__ void __launch_bounds__(SWE_THREADS)
swe_k5_ring(B2SweParams p, SweFrame f, const float* __restrict__ v, float* __restrict__ v_new,
            const float* __restrict__ fe2, const float* __restrict__ fn2) {
  b2_pdl_enter();
  int j, i0;
  bool m[4];
  if (!swe_frame_task(p, f, (long long)blockIdx.x * SWE_THREADS + threadIdx.x, j, i0, m)) return;
  if (!(m[0] || m[1] || m[2] || m[3])) return;
  swe_k5_ring_body(p, v, v_new, fe2, fn2, j, i0, m);
}

// Two independent roles in one launch (the step is latency-bound at 8 GPUs, every launch counts):
// CTAs [0, bulk_blocks) run the bulk kernel, the rest the frame pass that precedes the exchange.
__global__ void __launch_bounds__(SWE_THREADS, 2)
swe_k12_bulk_k1_frame(B2SweParams p, SweFrame f, unsigned bulk_blocks, const float* __restrict__ h,
                      float* __restrict__ h_new, const float* __restrict__ u, float* __restrict__ u_new,
                      const float* __restrict__ v, float* __restrict__ v_new, float* __restrict__ dh,
                      float* __restrict__ du, float* __restrict__ dv, float* __restrict__ fe,
                      float* __restrict__ fn, float* __restrict__ q, float* __restrict__ ke) {
  b2_pdl_enter();
  int j, i0;
  bool m[4];
  if (blockIdx.x < bulk_blocks) {
    if (!swe_map(p, j, i0, m)) return;
    if (j < 2 || j > p.ny - 3) return;
    swe_k12_body(p, h, h_new, u, u_new, v, v_new, dh, du, dv, j, i0, m);
  } else {
    const long long t = (long long)(blockIdx.x - bulk_blocks) * SWE_THREADS + threadIdx.x;
    if (!swe_frame_task(p, f, t, j, i0, m)) return;
    if (!(m[0] || m[1] || m[2] || m[3])) return;
    SweOut4 o;
    swe_k1_body(p, h, u, v, fe, fn, q, ke, j, i0, m, o);
  }
}

// (The friction phase is NOT merged the same way: its bulk kernel and the K34 frame pass both
// write u_new in the groups that contain ring and bulk lanes, so they must stay stream-ordered.)

static int k12_done(B2Comm* c, const char* name) {
  b2_count_launch(c);
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) {
    b2_set_error("%s: kernel launch failed: %s", name, cudaGetErrorString(err));
    return 1000 + (int)err;
  }
  return 0;
}

static unsigned blocks_for(long long tasks) {
  return (unsigned)((tasks + SWE_THREADS - 1) / SWE_THREADS);
}

// fuse_friction != 0: the friction phase is fused as well (bulk kernel + K34 on the frame ->
// exchange -> K5 on the ring): 16 instead of 21 array passes.
static int multistep_k12(B2Comm* c, const B2SweParams* p0, const B2SweState* st, const B2HaloDesc* topo,
                         int nsteps, int first_step, int fuse_friction, cudaStream_t s) {
  // blocks too small for the bulk / frame split, or no friction step to bring u, v back to their
  // home buffers: the stand-alone path
  if (!swe_k12_supported(*p0) || !(p0->viscosity > 0.f) || !st->u1 || !st->v1)
    return b2_swe_multistep(c, p0, st, topo, nsteps, first_step, s);
  B2SweParams p = *p0;
  if (p.pitch % 4 != 0 || p.pitch < p.nx) {
    b2_set_error("swe_multistep_k12: the row pitch must be a multiple of 4 floats and >= nx");
    return B2_ERR_BAD_ARG;
  }
  float* h = st->h0;
  float* hn = st->h1;
  float* const ua = st->u;
  float* const ub = st->u1;
  float* const va = st->v;
  float* const vb = st->v1;
  B2HaloDesc d = *topo;
  d.ny = p.ny;
  d.nx = p.nx;
  d.pitch = p.pitch;
  const SweFrame f2 = swe_frame(p, 2), f1 = swe_frame(p, 1);
  const unsigned all_blocks = blocks_for((long long)(p.ny - 2) * (p.pitch / 4));
  int rc = 0;
  for (int it = 0; it < nsteps && rc == 0; ++it) {
    p.first_step = (first_step && it == 0) ? 1 : 0;
    // bulk update and the frame's flux pass touch disjoint outputs: one launch, two CTA roles
    b2_launch(swe_k12_bulk_k1_frame, all_blocks + blocks_for(f2.total), SWE_THREADS, 0, s, p, f2, all_blocks,
              h, hn, ua, ub, va, vb, st->dh, st->du, st->dv, st->fe, st->fn, st->q, st->ke);
    if ((rc = k12_done(c, "swe_k12_bulk_k1_frame"))) break;
    d.nfields = 4;
    d.field[0] = st->fe; d.kind[0] = 1;
    d.field[1] = st->fn; d.kind[1] = 2;
    d.field[2] = st->q;  d.kind[2] = 0;
    d.field[3] = st->ke; d.kind[3] = 0;
    if ((rc = b2_halo_exchange(c, &d, s))) break;
    b2_launch(swe_k2_ring, blocks_for(f1.total), SWE_THREADS, 0, s, p, f1, h, hn, ua, ub, va, vb, st->dh,
              st->du, st->dv, st->fe, st->fn, st->q, st->ke);
    if ((rc = k12_done(c, "swe_k2_ring"))) break;
    d.nfields = 3;
    d.field[0] = hn; d.kind[0] = 0;
    d.field[1] = ub; d.kind[1] = 1;
    d.field[2] = vb; d.kind[2] = 2;
    if ((rc = b2_halo_exchange(c, &d, s))) break;
    const int has_south = topo->south >= 0;
    if (fuse_friction) {
      b2_launch(swe_k345_bulk, all_blocks, SWE_THREADS, 0, s, p, ub, ua, vb, va, has_south);
      if ((rc = k12_done(c, "swe_k345_bulk"))) break;
      b2_launch(swe_k34_frame, blocks_for(f2.total), SWE_THREADS, 0, s, p, f2, ub, ua, vb, st->fe2, st->fn2,
                has_south);
      if ((rc = k12_done(c, "swe_k34_frame"))) break;
    } else {
      if ((rc = b2_swe_friction_u_fused(c, &p, ub, ua, vb, st->fe2, st->fn2, has_south, s))) break;
    }
    d.nfields = 2;
    d.field[0] = st->fe2; d.kind[0] = 1;
    d.field[1] = st->fn2; d.kind[1] = 2;
    if ((rc = b2_halo_exchange(c, &d, s))) break;
    if (fuse_friction) {
      b2_launch(swe_k5_ring, blocks_for(f1.total), SWE_THREADS, 0, s, p, f1, vb, va, st->fe2, st->fn2);
      if ((rc = k12_done(c, "swe_k5_ring"))) break;
    } else {
      b2_launch(swe_k5_pp, all_blocks, SWE_THREADS, 0, s, p, vb, va, st->fe2, st->fn2);
      if ((rc = k12_done(c, "swe_k5_pp"))) break;
    }
    float* t = h; h = hn; hn = t;
  }
  if (rc == 0 && h != st->h0) {
    cudaError_t e = cudaMemcpyAsync(st->h0, h, (size_t)p.ny * p.pitch * sizeof(float),
                                    cudaMemcpyDeviceToDevice, s);
    if (e != cudaSuccess) {
      b2_set_error("swe_multistep_k12: copy failed: %s", cudaGetErrorString(e));
      rc = 1000 + (int)e;
    }
  }
  return rc;
}

extern "C" int b2_swe_multistep_k12(B2Comm* c, const B2SweParams* p0, const B2SweState* st,
                                    const B2HaloDesc* topo, int nsteps, int first_step, cudaStream_t s) {
  return multistep_k12(c, p0, st, topo, nsteps, first_step, 0, s);
}

extern "C" int b2_swe_multistep_k12f(B2Comm* c, const B2SweParams* p0, const B2SweState* st,
                                     const B2HaloDesc* topo, int nsteps, int first_step, cudaStream_t s) {
  return multistep_k12(c, p0, st, topo, nsteps, first_step, 1, s);
}
