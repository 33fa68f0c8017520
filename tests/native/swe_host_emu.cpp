// Host emulation of the shallow-water stencil bodies (csrc/b2_swe_body.cuh, b2_swe_k12_body.cuh,
// b2_swe_ca_body.cuh): the same source, compiled by g++ with the CUDA qualifiers defined away, driven
// by plain loops instead of a grid.  Lets the CPU test-suite check the INDEXING of kernels it cannot
// launch -- in particular the communication-avoiding step (frame kernels with owner views, deep halo
// exchange geometry, bulk / frame partition) against the stand-alone K1 -> K2 -> K34 -> K5 pipeline,
// bit for bit (every rounding in the bodies is explicit, so host and device agree on the arithmetic).
// Build: g++ -O1 -ffp-contract=off -shared -fPIC -I csrc -I $CUDA/include tests/native/swe_host_emu.cpp
#include <cmath>
#include <cstring>
#include <cuda_runtime.h>

static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
static inline float __frcp_rn(float a) { return 1.0f / a; }      // correctly rounded, like the intrinsic

// swe_map is grid-based; the harness enumerates (j, group) itself
struct FakeIdx { unsigned x; };
static FakeIdx blockIdx, threadIdx;

#include "b2_swe_ca_body.cuh"
#include "b2_swe_k12_body.cuh"

static void masks(const B2SweParams& p, int i0, bool m[4]) {
  for (int k = 0; k < 4; ++k) m[k] = (i0 + k >= 1) && (i0 + k <= p.nx - 2);
}

struct EmuStep {      // the arrays of one rank for one step (B2SweState roles of b2_swe_multistep_ca)
  const float *h;
  float *h_o, *u, *v, *dh, *du, *dv, *dub, *dvb, *upf, *vpf;
};

extern "C" {

// K1 on every interior row / group (what swe_k1_fluxes does)
void emu_k1_all(const B2SweParams* p, const float* h, const float* u, const float* v, float* fe, float* fn,
                float* q, float* ke) {
  for (int j = 1; j <= p->ny - 2; ++j)
    for (int i0 = 0; i0 < p->pitch; i0 += 4) {
      bool m[4];
      masks(*p, i0, m);
      if (!(m[0] || m[1] || m[2] || m[3])) continue;
      swe_k1_body(*p, h, u, v, fe, fn, q, ke, j, i0, m);
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
      swe_k2_body(*p, h, h_new, u, v, dh, du, dv, fe, fn, q, ke, j, i0, m);
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

// merged friction kernel: u -> u_new, friction-v fluxes (swe_k34_friction_u)
void emu_k34(const B2SweParams* p, const float* u, float* u_new, const float* v, float* fe2, float* fn2,
             int has_south) {
  for (int j = 1; j <= p->ny - 2; ++j)
    for (int i0 = 0; i0 < p->pitch; i0 += 4) {
      bool m[4];
      masks(*p, i0, m);
      if (!(m[0] || m[1] || m[2] || m[3])) continue;
      swe_k34_body(*p, u, u_new, v, fe2, fn2, j, i0, m, has_south != 0);
    }
}

// ---- communication-avoiding step (csrc/b2_swe_ca_body.cuh, b2_swe_k12_body.cuh): the kernels of
// b2_swe_ca.cu as loops.  `reverse` walks the tasks backwards: a kernel whose threads only read what no
// thread of the same launch (or phase) writes gives the same bits in any order.
static CACtx make_ctx(const B2SweParams* p, const B2SweCA* x, const EmuStep* e) {
  CACtx c;
  c.p = *p; c.x = *x; c.x.cb1 = swe_ca_cb1(*p);
  c.h = e->h; c.hn = e->h_o; c.ua = e->u; c.va = e->v; c.dh = e->dh; c.du = e->du; c.dv = e->dv;
  c.dub = e->dub; c.dvb = e->dvb;
  c.upf = e->upf; c.vpf = e->vpf;
  return c;
}

int emu_ca_supported(const B2SweParams* p) { return swe_ca_supported(*p) ? 1 : 0; }
int emu_ca_cb1(const B2SweParams* p) { return swe_ca_cb1(*p); }

void emu_ca_tend_frame(const B2SweParams* p, const B2SweCA* x, const EmuStep* e, int reverse) {
  const CACtx c = make_ctx(p, x, e);
  const CAFrame f = ca_frame(c.p, 5, c.x.cb1 - 2);
  for (long long k = 0; k < f.total; ++k) {
    int j, i;
    if (ca_frame_cell(c.p, f, (int)(reverse ? f.total - 1 - k : k), j, i)) swe_ca_tend_cell(c, j, i);
  }
}

void emu_ca_fric_frame(const B2SweParams* p, const B2SweCA* x, const EmuStep* e, int reverse) {
  const CACtx c = make_ctx(p, x, e);
  const CAFrame f = ca_frame(c.p, 3, c.x.cb1);
  const long long n = f.total + ca_ext_total(c.p);
  for (long long k = 0; k < n; ++k) swe_ca_fric_task(c, f, e->u, e->v, reverse ? n - 1 - k : k);
}

// swe_ca_bulk_step: CTAs one after the other, every phase as a loop over the CTA's threads
// (`reverse`: CTAs and threads backwards -- a phase only reads what earlier phases wrote)
// swe_ca_bulk_k12 + swe_ca_bulk_fric (the friction kernel runs after kernel A: it reads the band's u', v')
void emu_ca_bulk_k12(const B2SweParams* p, const EmuStep* e) {
  const int cb1 = swe_ca_cb1(*p);
  for (long long t = 0; t < ca_bulk_tasks(*p, cb1); ++t) {
    int j, i0;
    ca_bulk_task(*p, cb1, t, j, i0);
    swe_k12_body(*p, e->h, e->h_o, e->u, e->upf, e->v, e->vpf, e->dh, e->du, e->dv, e->dh, e->du, e->dv, j, i0);
  }
}
void emu_ca_bulk_fric(const B2SweParams* p, const EmuStep* e) {
  const int cb1 = swe_ca_cb1(*p);
  for (long long t = 0; t < ca_bulk_tasks(*p, cb1); ++t) {
    int j, i0;
    ca_bulk_task(*p, cb1, t, j, i0);
    swe_k345_body(*p, e->upf, e->u, e->vpf, e->v, j, i0);
  }
}

// which interior cells do the bulk kernel and the frame kernels WRITE?  marks[j * nx + i]: +1 frame
// (kernel D's cells = kernel A's full updates), +16 bulk, +256 kernel A's u' / v' band
void emu_ca_marks(const B2SweParams* p, int* marks) {
  const int cb1 = swe_ca_cb1(*p);
  const CAFrame fd = ca_frame(*p, 3, cb1), fa = ca_frame(*p, 5, cb1 - 2);
  int j, i;
  for (long long t = 0; t < fd.total; ++t)
    if (ca_frame_cell(*p, fd, (int)t, j, i)) marks[j * p->nx + i] += 1;
  for (long long t = 0; t < fa.total; ++t)
    if (ca_frame_cell(*p, fa, (int)t, j, i)) marks[j * p->nx + i] += 256;
  for (long long t = 0; t < ca_bulk_tasks(*p, cb1); ++t) {
    int i0;
    ca_bulk_task(*p, cb1, t, j, i0);
    for (int k = 0; k < 4; ++k) marks[j * p->nx + i0 + k] += 16;
  }
}

// the message that lands on `side` of the receiver: read from the sender's arrays exactly as the push
// loop of b2_k_halo_ca does, scattered exactly as its poll loop does (no transport in between)
void emu_ca_deliver(int ny, int nx, int pitch, int epitch, int side, int recv_has_south, int recv_has_north,
                    float* const* send_field, float* const* recv_field, float* const* recv_ext) {
  const int jlo = recv_has_south ? 1 : 0, jhi = recv_has_north ? ny - 1 : ny;
  for (int e = 0; e < ca_msg_count(ny, nx, side); ++e) {
    int f, js, is, jr, ir, layer;
    ca_msg_elem(ny, nx, side, e, f, js, is, jr, ir, layer);
    ca_scatter(ny, nx, (size_t)pitch, epitch, side, e, jlo, jhi, send_field[f][(size_t)js * pitch + is], recv_field,
               recv_ext);
  }
}

// swe_ca_init_ext
void emu_ca_init_ext(const B2SweParams* p, const B2SweCA* x, const float* u, const float* v) {
  int j, i;
  for (long long t = 0; t < ca_ext_total(*p); ++t)
    if (ca_ext_cell(*p, t, j, i)) {
      const size_t e = ca_e(*x, j, i);
      x->uppx[e] = x->upx[e];
      x->vppx[e] = x->vpx[e];
    }
  for (j = 1; j <= p->ny - 2; ++j)
    for (i = 1; i <= p->nx - 2; ++i)
      if (j == 1 || j == p->ny - 2 || i == 1 || i == p->nx - 2) {
        x->upx[ca_e(*x, j, i)] = u[ca_m(*p, j, i)];
        x->vpx[ca_e(*x, j, i)] = v[ca_m(*p, j, i)];
      }
}

}  // extern "C"
