// Host emulation of the shallow-water stencil bodies (csrc/b2_swe_body.cuh, b2_swe_k12_body.cuh):
// the same source, compiled by g++ with the CUDA qualifiers defined away, driven by plain loops
// instead of a grid.  Lets the CPU test-suite check the INDEXING of kernels it cannot launch:
// the fused flux+tendency path (K12 on the bulk, K1/K2 on the frame) against K1 -> K2 everywhere.
// Build: g++ -O1 -ffp-contract=off -shared -fPIC -I csrc -I $CUDA/include tests/native/swe_host_emu.cpp
#include <cmath>
#include <cstring>
#include <cuda_runtime.h>

static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }

// swe_map is grid-based; the harness enumerates (j, group) itself
struct FakeIdx { unsigned x; };
static FakeIdx blockIdx, threadIdx;

#include "b2_swe_k12_body.cuh"

static void masks(const B2SweParams& p, int i0, bool m[4]) {
  for (int k = 0; k < 4; ++k) m[k] = (i0 + k >= 1) && (i0 + k <= p.nx - 2);
}

extern "C" {

// K1 on every interior row / group (what swe_k1_fluxes does)
void emu_k1_all(const B2SweParams* p, const float* h, const float* u, const float* v, float* fe, float* fn,
                float* q, float* ke) {
  for (int j = 1; j <= p->ny - 2; ++j)
    for (int i0 = 0; i0 < p->pitch; i0 += 4) {
      bool m[4];
      masks(*p, i0, m);
      if (!(m[0] || m[1] || m[2] || m[3])) continue;
      SweOut4 o;
      swe_k1_body(*p, h, u, v, fe, fn, q, ke, j, i0, m, o);
    }
}

// K2 everywhere, in place on u, v like swe_k2_tendencies
void emu_k2_all(const B2SweParams* p, const float* h, float* h_new, float* u, float* v, float* dh, float* du,
                float* dv, const float* fe, const float* fn, const float* q, const float* ke) {
  for (int j = 1; j <= p->ny - 2; ++j)
    for (int i0 = 0; i0 < p->pitch; i0 += 4) {
      bool m[4];
      masks(*p, i0, m);
      if (!(m[0] || m[1] || m[2] || m[3])) continue;
      SweOut4 o;
      swe_k2_body(*p, h, h_new, u, v, dh, du, dv, fe, fn, q, ke, j, i0, m, o);
    }
}

// frame pass of K1 (width 2)
void emu_k1_frame(const B2SweParams* p, const float* h, const float* u, const float* v, float* fe, float* fn,
                  float* q, float* ke) {
  const SweFrame f = swe_frame(*p, 2);
  for (long long t = 0; t < f.total; ++t) {
    int j, i0;
    bool m[4];
    if (!swe_frame_task(*p, f, t, j, i0, m)) continue;
    if (!(m[0] || m[1] || m[2] || m[3])) continue;
    SweOut4 o;
    swe_k1_body(*p, h, u, v, fe, fn, q, ke, j, i0, m, o);
  }
}

void emu_k12_bulk(const B2SweParams* p, const float* h, float* h_new, const float* u, float* u_new,
                  const float* v, float* v_new, float* dh, float* du, float* dv) {
  for (int j = 2; j <= p->ny - 3; ++j)
    for (int i0 = 0; i0 < p->pitch; i0 += 4) {
      bool m[4];
      masks(*p, i0, m);
      if (!(m[0] || m[1] || m[2] || m[3])) continue;
      swe_k12_body(*p, h, h_new, u, u_new, v, v_new, dh, du, dv, j, i0, m);
    }
}

void emu_k2_ring(const B2SweParams* p, const float* h, float* h_new, const float* u, float* u_new,
                 const float* v, float* v_new, float* dh, float* du, float* dv, const float* fe,
                 const float* fn, const float* q, const float* ke) {
  const SweFrame f = swe_frame(*p, 1);
  for (long long t = 0; t < f.total; ++t) {
    int j, i0;
    bool m[4];
    if (!swe_frame_task(*p, f, t, j, i0, m)) continue;
    if (!(m[0] || m[1] || m[2] || m[3])) continue;
    swe_k2_ring_body(*p, h, h_new, u, u_new, v, v_new, dh, du, dv, fe, fn, q, ke, j, i0, m);
  }
}

void emu_k5(const B2SweParams* p, float* v, const float* fe2, const float* fn2) {
  for (int j = 1; j <= p->ny - 2; ++j)
    for (int i0 = 0; i0 < p->pitch; i0 += 4) {
      bool m[4];
      masks(*p, i0, m);
      if (!(m[0] || m[1] || m[2] || m[3])) continue;
      swe_k5_body(*p, v, fe2, fn2, j, i0, m);
    }
}

void emu_k5_pp(const B2SweParams* p, const float* v, float* v_new, const float* fe2, const float* fn2) {
  for (int j = 1; j <= p->ny - 2; ++j)
    for (int i0 = 0; i0 < p->pitch; i0 += 4) {
      bool m[4];
      masks(*p, i0, m);
      if (!(m[0] || m[1] || m[2] || m[3])) continue;
      swe_k5_pp_body(*p, v, v_new, fe2, fn2, j, i0, m);
    }
}

// merged friction kernel: u -> u_new, friction-v fluxes (swe_k34_friction_u)
void emu_k34(const B2SweParams* p, const float* u, float* u_new, const float* v, float* fe2, float* fn2,
             int has_south) {
  for (int j = 1; j <= p->ny - 2; ++j)
    for (int i0 = 0; i0 < p->pitch; i0 += 4) {
      bool m[4];
      masks(*p, i0, m);
      if (!(m[0] || m[1] || m[2] || m[3])) continue;
      SweOut4 o;
      swe_k34_body(*p, u, u_new, v, fe2, fn2, j, i0, m, has_south != 0, o);
    }
}

// the two-kernel formulation it replaced: K3 (fluxes of u, local halo) then K4 (apply + v fluxes)
void emu_k3_k4(const B2SweParams* p, float* u, const float* v, float* fe, float* fn, float* fe2, float* fn2,
               int has_south) {
  for (int j = 1; j <= p->ny - 2; ++j)
    for (int i0 = 0; i0 < p->pitch; i0 += 4) {
      bool m[4];
      masks(*p, i0, m);
      if (!(m[0] || m[1] || m[2] || m[3])) continue;
      swe_k3_body(*p, u, fe, fn, j, i0, m, true, has_south != 0);
    }
  for (int j = 1; j <= p->ny - 2; ++j)
    for (int i0 = 0; i0 < p->pitch; i0 += 4) {
      bool m[4];
      masks(*p, i0, m);
      if (!(m[0] || m[1] || m[2] || m[3])) continue;
      SweOut4 o;
      swe_k4_body(*p, u, v, fe, fn, fe2, fn2, j, i0, m, o);
    }
}

// friction phase of the fused pipeline: bulk kernel, K34 on the frame (width 2), K5 on the ring
void emu_k345_bulk(const B2SweParams* p, const float* u, float* u_new, const float* v, float* v_new,
                   int has_south) {
  for (int j = 2; j <= p->ny - 3; ++j)
    for (int i0 = 0; i0 < p->pitch; i0 += 4) {
      bool m[4];
      masks(*p, i0, m);
      if (!(m[0] || m[1] || m[2] || m[3])) continue;
      swe_k345_body(*p, u, u_new, v, v_new, j, i0, has_south != 0);
    }
}

void emu_k34_frame(const B2SweParams* p, const float* u, float* u_new, const float* v, float* fe2,
                   float* fn2, int has_south) {
  const SweFrame f = swe_frame(*p, 2);
  for (long long t = 0; t < f.total; ++t) {
    int j, i0;
    bool m[4];
    if (!swe_frame_task(*p, f, t, j, i0, m)) continue;
    if (!(m[0] || m[1] || m[2] || m[3])) continue;
    SweOut4 o;
    swe_k34_body(*p, u, u_new, v, fe2, fn2, j, i0, m, has_south != 0, o);
  }
}

void emu_k5_ring(const B2SweParams* p, const float* v, float* v_new, const float* fe2, const float* fn2) {
  const SweFrame f = swe_frame(*p, 1);
  for (long long t = 0; t < f.total; ++t) {
    int j, i0;
    bool m[4];
    if (!swe_frame_task(*p, f, t, j, i0, m)) continue;
    if (!(m[0] || m[1] || m[2] || m[3])) continue;
    swe_k5_ring_body(*p, v, v_new, fe2, fn2, j, i0, m);
  }
}

int emu_k12_supported(const B2SweParams* p) { return swe_k12_supported(*p) ? 1 : 0; }

// which (row, group) tasks does a frame of width w enumerate?  marks[j * ngroups + g] += 1
long long emu_frame_marks(const B2SweParams* p, int w, int* marks) {
  const SweFrame f = swe_frame(*p, w);
  for (long long t = 0; t < f.total; ++t) {
    int j, i0;
    bool m[4];
    if (swe_frame_task(*p, f, t, j, i0, m)) marks[j * f.ngroups + (i0 >> 2)] += 1;
  }
  return f.total;
}

}  // extern "C"
