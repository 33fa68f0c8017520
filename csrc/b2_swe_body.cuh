// mpi4jax_b200 -- shallow-water stencil bodies shared by the stand-alone kernels
// (b2_swe.cu) and the kernels that have the halo exchange fused in (b2_swe_fused.cu).
// Each body handles one aligned group of 4 cells of row j: loads, arithmetic, vector stores;
// it returns the freshly computed lane values so that a fused kernel can push edge cells to
// the neighbouring GPUs straight from registers.
#pragma once

#include "b2_runtime.h"

struct B2SweParams {
  int ny, nx, pitch;
  float dx, dy, dt, gravity, viscosity;
  float rdx, rdy;             // 1/dx, 1/dy: x * (1/dx) instead of x / dx (<= 1 ulp apart, ~8x cheaper)
  float ab_a, ab_b;           // Adams-Bashforth weights
  int first_step;
  int south_wall, north_wall; // this rank touches the y walls (hc edge padding)
  const float* coriolis;      // [ny]
};

// Device pointers of a model's fields (mirrored by ctypes in _src/native/__init__.py; the layout
// is part of the ABI checked at import).  h0/h1 and u/u1 are ping-pong pairs.
struct B2SweState {
  float *h0, *h1, *u, *v, *dh, *du, *dv, *fe, *fn, *q, *ke, *fe2, *fn2;
  float* u1;   // partner of u for the friction update of the stand-alone path (b2_swe.cu)
  float* v1;   // partner of v, used by the fused flux+tendency path only (b2_swe_k12.cu)
};

// "ext" arrays of the communication-avoiding step (b2_swe_ca_body.cuh): [(ny + 4) x epitch] floats,
// cell (j, i) at (j + 2) * epitch + (i + 2); only cells beyond the interior (and the ring mirror) are used
struct B2SweCA {
  float *hx, *upx, *vpx, *uppx, *vppx;
  int epitch;
  int cb1;            // first column of the east frame: a multiple of 4, <= nx - 4 (filled in natively)
};

#define SWE_THREADS 256

// An aligned group of four cells plus its west / east neighbours.
struct Row6 {
  float w, c0, c1, c2, c3, e;
};

__device__ __forceinline__ float4 ld4(const float* __restrict__ a, size_t off) {
  return *reinterpret_cast<const float4*>(a + off);
}
__device__ __forceinline__ void st4(float* __restrict__ a, size_t off, float4 v) {
  *reinterpret_cast<float4*>(a + off) = v;
}
// row `j`, group starting at column i0 (multiple of 4); neighbours are fetched only when asked
template <bool W, bool E>
__device__ __forceinline__ Row6 ld_row(const float* __restrict__ a, int j, int i0, int pitch) {
  const size_t off = (size_t)j * pitch + i0;
  const float4 v = ld4(a, off);
  Row6 r;
  r.c0 = v.x; r.c1 = v.y; r.c2 = v.z; r.c3 = v.w;
  r.w = (W && i0 > 0) ? a[off - 1] : 0.f;
  r.e = (E && i0 + 4 < pitch) ? a[off + 4] : 0.f;
  return r;
}
__device__ __forceinline__ float4 sel4(const bool m[4], float4 a, float4 b) {
  return make_float4(m[0] ? a.x : b.x, m[1] ? a.y : b.y, m[2] ? a.z : b.z, m[3] ? a.w : b.w);
}

// thread -> (row j in [1, ny-2], group g); returns false when out of range
__device__ __forceinline__ bool swe_map(const B2SweParams& p, int& j, int& i0, bool m[4]) {
  const int ngroups = p.pitch >> 2;
  const long long idx = (long long)blockIdx.x * SWE_THREADS + threadIdx.x;
  if (idx >= (long long)(p.ny - 2) * ngroups) return false;
  j = (int)(idx / ngroups) + 1;
  i0 = (int)(idx % ngroups) << 2;
#pragma unroll
  for (int k = 0; k < 4; ++k) m[k] = (i0 + k >= 1) && (i0 + k <= p.nx - 2);
  return m[0] || m[1] || m[2] || m[3];
}

__device__ __forceinline__ int hc_row(const B2SweParams& p, int j) {
  // hc = h with the physical wall rows replaced by their neighbouring interior row
  if (p.south_wall && j == 0) return 1;
  if (p.north_wall && j == p.ny - 1) return p.ny - 2;
  return j;
}

struct SweOut4 { float a[4][4]; };   // [field][lane]

// B2_SWE_EXPLICIT_ROUNDING=1 makes swe_k1_body use the explicit-rounding helpers below instead of
// plain expressions (whose FMA contraction ptxas picks, differently per lane).  With it, a kernel
// that recomputes the flux quantities (b2_swe_k12.cu) agrees with the two-kernel path to the bit.
// Default 0: the validated binaries were built from the plain expressions; flip after a GPU run.
#ifndef B2_SWE_EXPLICIT_ROUNDING
#define B2_SWE_EXPLICIT_ROUNDING 1
#endif

// ---- the four diagnostic quantities of the flux kernel, with every rounding spelled out --------
// (explicit round-to-nearest intrinsics are never contracted by ptxas).  Spelling them out makes
// the result independent of the kernel the expression is inlined into, which is what allows a
// kernel that RECOMPUTES these quantities at its stencil neighbours to be bit-identical to one
// that reads them from memory.  The operation order is the one nvcc chose for the plain
// expressions of swe_k1_body (read off its SASS: which products are rounded before the FMA).
__device__ __forceinline__ float swe_fe(float h_c, float h_e, float u_c) {          // mass flux east
  return __fmul_rn(__fmul_rn(__fadd_rn(h_c, h_e), 0.5f), u_c);
}
__device__ __forceinline__ float swe_fn(float h_c, float h_n, float v_c) {          // mass flux north
  return __fmul_rn(__fmul_rn(__fadd_rn(h_c, h_n), 0.5f), v_c);
}
__device__ __forceinline__ float swe_q(const B2SweParams& p, float cor, float v_e, float v_c, float u_n,
                                       float u_c, float h_c, float h_e, float h_n, float h_ne) {
  const float rel = __fmaf_rn(__fadd_rn(v_e, -v_c), p.rdx, -__fmul_rn(__fadd_rn(u_n, -u_c), p.rdy));
  const float den = __fmul_rn(__fadd_rn(h_ne, __fadd_rn(h_n, __fadd_rn(h_c, h_e))), 0.25f);
  return __fmul_rn(__fadd_rn(cor, rel), 1.0f / den);                                 // potential vorticity
}
__device__ __forceinline__ float swe_ke(float u_c, float u_w, float v_c, float v_s) {   // kinetic energy
  const float uu = __fmaf_rn(u_c, u_c, __fmul_rn(u_w, u_w));
  const float vv = __fmaf_rn(v_c, v_c, __fmul_rn(v_s, v_s));
  return __fmul_rn(__fmaf_rn(vv, 0.5f, __fmul_rn(uu, 0.5f)), 0.5f);
}

__device__ __forceinline__ void swe_k1_body(const B2SweParams& p, const float* __restrict__ h,
                                            const float* __restrict__ u, const float* __restrict__ v,
                                            float* __restrict__ fe, float* __restrict__ fn,
                                            float* __restrict__ q, float* __restrict__ ke, int j,
                                            int i0, const bool m[4], SweOut4& o) {
  const int P = p.pitch;
  const Row6 hc = ld_row<false, true>(h, hc_row(p, j), i0, P);
  const Row6 hn = ld_row<false, true>(h, hc_row(p, j + 1), i0, P);
  const Row6 uc = ld_row<true, false>(u, j, i0, P);
  const Row6 un = ld_row<false, false>(u, j + 1, i0, P);
  const Row6 vc = ld_row<false, true>(v, j, i0, P);
  const Row6 vs = ld_row<false, false>(v, j - 1, i0, P);
  const float cor = p.coriolis[j];
  const float H[5] = {hc.c0, hc.c1, hc.c2, hc.c3, hc.e}, HN[5] = {hn.c0, hn.c1, hn.c2, hn.c3, hn.e};
  const float U[5] = {uc.w, uc.c0, uc.c1, uc.c2, uc.c3}, UN[4] = {un.c0, un.c1, un.c2, un.c3};
  const float V[5] = {vc.c0, vc.c1, vc.c2, vc.c3, vc.e}, VS[4] = {vs.c0, vs.c1, vs.c2, vs.c3};
  float FE[4], FN[4], Q[4], KE[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float uk = U[k + 1], vk = V[k];
#if B2_SWE_EXPLICIT_ROUNDING
    FE[k] = swe_fe(H[k], H[k + 1], uk);
    FN[k] = swe_fn(H[k], HN[k], vk);
    Q[k] = swe_q(p, cor, V[k + 1], vk, UN[k], uk, H[k], H[k + 1], HN[k], HN[k + 1]);
    KE[k] = swe_ke(uk, U[k], vk, VS[k]);
#else
    FE[k] = 0.5f * (H[k] + H[k + 1]) * uk;
    FN[k] = 0.5f * (H[k] + HN[k]) * vk;
    const float rel = (V[k + 1] - vk) * p.rdx - (UN[k] - uk) * p.rdy;
    Q[k] = (cor + rel) * (1.0f / (0.25f * (H[k] + H[k + 1] + HN[k] + HN[k + 1])));
    KE[k] = 0.5f * (0.5f * (uk * uk + U[k] * U[k]) + 0.5f * (vk * vk + VS[k] * VS[k]));
#endif
    if (!m[k]) FE[k] = FN[k] = Q[k] = KE[k] = 0.f;   // halo / pad lanes: refreshed by the exchange
    if (p.north_wall && j == p.ny - 2) FN[k] = 0.f;    // "v" wall rule (shallow_water.py:261-262)
    o.a[0][k] = FE[k]; o.a[1][k] = FN[k]; o.a[2][k] = Q[k]; o.a[3][k] = KE[k];
  }
  const size_t off = (size_t)j * P + i0;
  st4(fe, off, make_float4(FE[0], FE[1], FE[2], FE[3]));
  st4(fn, off, make_float4(FN[0], FN[1], FN[2], FN[3]));
  st4(q, off, make_float4(Q[0], Q[1], Q[2], Q[3]));
  st4(ke, off, make_float4(KE[0], KE[1], KE[2], KE[3]));
}

struct SweK2In {
  float fe_c, fe_w, fen_c, fen_w;      // fe[j][i], fe[j][i-1], fe[j+1][i], fe[j+1][i-1]
  float fn_c, fn_e, fns_c, fns_e;      // fn[j][i], fn[j][i+1], fn[j-1][i], fn[j-1][i+1]
  float q_c, q_w, qs_c;                // q[j][i], q[j][i-1], q[j-1][i]
  float ke_c, ke_e, ken_c;             // ke[j][i], ke[j][i+1], ke[j+1][i]
  float h_c, h_e, h_n;                 // h[j][i], h[j][i+1], h[j+1][i]
  float u_o, v_o, dh_o, du_o, dv_o;    // own cell, old values
};
struct SweK2Out {
  float h, u, v, dh, du, dv;
};

// The per-cell arithmetic of the tendency kernel with every rounding spelled out (cf. swe_fe ..
// swe_ke above): the same bits whether it is inlined into the vectorised kernels (swe_k2_body,
// swe_k12_body) or into the scalar frame kernels of b2_swe_ca_body.cuh.
__device__ __forceinline__ SweK2Out swe_k2_cell(const B2SweParams& p, const SweK2In& x) {
  // dh = -(fe_c - fe_w) / dx - (fn_c - fn_s) / dy
  const float dh_new = __fmaf_rn(__fadd_rn(x.fns_c, -x.fn_c), p.rdy, -__fmul_rn(__fadd_rn(x.fe_c, -x.fe_w), p.rdx));
  // du = -g (h_e - h_c) / dx + 1/2 (q_c (fn_c + fn_e) / 2 + q_s (fn_s + fn_se) / 2) - (ke_e - ke_c) / dx
  const float gu = __fmul_rn(__fmul_rn(-p.gravity, __fadd_rn(x.h_e, -x.h_c)), p.rdx);
  const float su = __fadd_rn(__fmul_rn(__fmul_rn(x.q_c, 0.5f), __fadd_rn(x.fn_c, x.fn_e)),
                             __fmul_rn(__fmul_rn(x.qs_c, 0.5f), __fadd_rn(x.fns_c, x.fns_e)));
  float du_new = __fmaf_rn(0.5f, su, gu);
  du_new = __fmaf_rn(__fadd_rn(x.ke_c, -x.ke_e), p.rdx, du_new);
  // dv = -g (h_n - h_c) / dy - 1/2 (q_c (fe_c + fe_n) / 2 + q_w (fe_w + fe_nw) / 2) - (ke_n - ke_c) / dy
  const float gv = __fmul_rn(__fmul_rn(-p.gravity, __fadd_rn(x.h_n, -x.h_c)), p.rdy);
  const float sv = __fadd_rn(__fmul_rn(__fmul_rn(x.q_c, 0.5f), __fadd_rn(x.fe_c, x.fen_c)),
                             __fmul_rn(__fmul_rn(x.q_w, 0.5f), __fadd_rn(x.fe_w, x.fen_w)));
  float dv_new = __fmaf_rn(-0.5f, sv, gv);
  dv_new = __fmaf_rn(__fadd_rn(x.ke_c, -x.ken_c), p.rdy, dv_new);
  SweK2Out o;
  if (p.first_step) {
    o.u = __fmaf_rn(p.dt, du_new, x.u_o);
    o.v = __fmaf_rn(p.dt, dv_new, x.v_o);
    o.h = __fmaf_rn(p.dt, dh_new, x.h_c);
  } else {
    o.u = __fmaf_rn(p.dt, __fmaf_rn(p.ab_a, du_new, __fmul_rn(p.ab_b, x.du_o)), x.u_o);
    o.v = __fmaf_rn(p.dt, __fmaf_rn(p.ab_a, dv_new, __fmul_rn(p.ab_b, x.dv_o)), x.v_o);
    o.h = __fmaf_rn(p.dt, __fmaf_rn(p.ab_a, dh_new, __fmul_rn(p.ab_b, x.dh_o)), x.h_c);
  }
  o.dh = dh_new; o.du = du_new; o.dv = dv_new;
  return o;
}

__device__ __forceinline__ void swe_k2_body(const B2SweParams& p, const float* __restrict__ h,
                                            float* __restrict__ h_new, float* __restrict__ u,
                                            float* __restrict__ v, float* __restrict__ dh,
                                            float* __restrict__ du, float* __restrict__ dv,
                                            const float* __restrict__ fe, const float* __restrict__ fn,
                                            const float* __restrict__ q, const float* __restrict__ ke,
                                            int j, int i0, const bool m[4], SweOut4& o) {
  const int P = p.pitch;
  const size_t off = (size_t)j * P + i0;
  const Row6 fec = ld_row<true, false>(fe, j, i0, P), fen = ld_row<true, false>(fe, j + 1, i0, P);
  const Row6 fnc = ld_row<false, true>(fn, j, i0, P), fns = ld_row<false, true>(fn, j - 1, i0, P);
  const Row6 qc = ld_row<true, false>(q, j, i0, P), qs = ld_row<false, false>(q, j - 1, i0, P);
  const Row6 kec = ld_row<false, true>(ke, j, i0, P), ken = ld_row<false, false>(ke, j + 1, i0, P);
  const Row6 hc = ld_row<false, true>(h, j, i0, P), hn = ld_row<false, false>(h, j + 1, i0, P);
  const float4 u4 = ld4(u, off), v4 = ld4(v, off);
  float4 dh4 = make_float4(0, 0, 0, 0), du4 = dh4, dv4 = dh4;
  if (!p.first_step) { dh4 = ld4(dh, off); du4 = ld4(du, off); dv4 = ld4(dv, off); }
  const float FE[5] = {fec.w, fec.c0, fec.c1, fec.c2, fec.c3}, FEN[5] = {fen.w, fen.c0, fen.c1, fen.c2, fen.c3};
  const float FN[5] = {fnc.c0, fnc.c1, fnc.c2, fnc.c3, fnc.e}, FNS[5] = {fns.c0, fns.c1, fns.c2, fns.c3, fns.e};
  const float Q[5] = {qc.w, qc.c0, qc.c1, qc.c2, qc.c3}, QS[4] = {qs.c0, qs.c1, qs.c2, qs.c3};
  const float KE[5] = {kec.c0, kec.c1, kec.c2, kec.c3, kec.e}, KEN[4] = {ken.c0, ken.c1, ken.c2, ken.c3};
  const float H[5] = {hc.c0, hc.c1, hc.c2, hc.c3, hc.e}, HN[4] = {hn.c0, hn.c1, hn.c2, hn.c3};
  const float Uo[4] = {u4.x, u4.y, u4.z, u4.w}, Vo[4] = {v4.x, v4.y, v4.z, v4.w};
  const float DHo[4] = {dh4.x, dh4.y, dh4.z, dh4.w}, DUo[4] = {du4.x, du4.y, du4.z, du4.w},
              DVo[4] = {dv4.x, dv4.y, dv4.z, dv4.w};
  float Un[4], Vn[4], Hn[4], DH[4], DU[4], DV[4];
#if B2_SWE_EXPLICIT_ROUNDING
  // the shared per-cell function (also used by the fused flux+tendency kernel)
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    SweK2In in;
    in.fe_c = FE[k + 1]; in.fe_w = FE[k]; in.fen_c = FEN[k + 1]; in.fen_w = FEN[k];
    in.fn_c = FN[k]; in.fn_e = FN[k + 1]; in.fns_c = FNS[k]; in.fns_e = FNS[k + 1];
    in.q_c = Q[k + 1]; in.q_w = Q[k]; in.qs_c = QS[k];
    in.ke_c = KE[k]; in.ke_e = KE[k + 1]; in.ken_c = KEN[k];
    in.h_c = H[k]; in.h_e = H[k + 1]; in.h_n = HN[k];
    in.u_o = Uo[k]; in.v_o = Vo[k]; in.dh_o = DHo[k]; in.du_o = DUo[k]; in.dv_o = DVo[k];
    const SweK2Out r = swe_k2_cell(p, in);
    Un[k] = r.u; Vn[k] = r.v; Hn[k] = r.h; DH[k] = r.dh; DU[k] = r.du; DV[k] = r.dv;
    if (!m[k]) { Un[k] = Uo[k]; Vn[k] = Vo[k]; Hn[k] = H[k]; DH[k] = DU[k] = DV[k] = 0.f; }
    if (p.north_wall && j == p.ny - 2) Vn[k] = 0.f;
    o.a[0][k] = Hn[k]; o.a[1][k] = Un[k]; o.a[2][k] = Vn[k];
  }
#else
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float fe_c = FE[k + 1], fe_w = FE[k], fn_c = FN[k], q_c = Q[k + 1];
    const float dh_new = -(fe_c - fe_w) * p.rdx - (fn_c - FNS[k]) * p.rdy;
    float du_new = -p.gravity * (H[k + 1] - H[k]) * p.rdx +
                   0.5f * (q_c * 0.5f * (fn_c + FN[k + 1]) + QS[k] * 0.5f * (FNS[k] + FNS[k + 1]));
    float dv_new = -p.gravity * (HN[k] - H[k]) * p.rdy -
                   0.5f * (q_c * 0.5f * (fe_c + FEN[k + 1]) + Q[k] * 0.5f * (fe_w + FEN[k]));
    du_new += -(KE[k + 1] - KE[k]) * p.rdx;
    dv_new += -(KEN[k] - KE[k]) * p.rdy;
    if (p.first_step) {
      Un[k] = Uo[k] + p.dt * du_new;
      Vn[k] = Vo[k] + p.dt * dv_new;
      Hn[k] = H[k] + p.dt * dh_new;
    } else {
      Un[k] = Uo[k] + p.dt * (p.ab_a * du_new + p.ab_b * DUo[k]);
      Vn[k] = Vo[k] + p.dt * (p.ab_a * dv_new + p.ab_b * DVo[k]);
      Hn[k] = H[k] + p.dt * (p.ab_a * dh_new + p.ab_b * DHo[k]);
    }
    DH[k] = dh_new; DU[k] = du_new; DV[k] = dv_new;
    if (!m[k]) { Un[k] = Uo[k]; Vn[k] = Vo[k]; Hn[k] = H[k]; DH[k] = DU[k] = DV[k] = 0.f; }
    if (p.north_wall && j == p.ny - 2) Vn[k] = 0.f;    // "v" wall rule, applied after the update
    o.a[0][k] = Hn[k]; o.a[1][k] = Un[k]; o.a[2][k] = Vn[k];
  }
#endif
  st4(u, off, make_float4(Un[0], Un[1], Un[2], Un[3]));
  st4(v, off, make_float4(Vn[0], Vn[1], Vn[2], Vn[3]));
  st4(h_new, off, make_float4(Hn[0], Hn[1], Hn[2], Hn[3]));
  st4(dh, off, make_float4(DH[0], DH[1], DH[2], DH[3]));
  st4(du, off, make_float4(DU[0], DU[1], DU[2], DU[3]));
  st4(dv, off, make_float4(DV[0], DV[1], DV[2], DV[3]));
}

// `local_halo`: also produce the WEST halo column of fe and the SOUTH halo row of fn from this
// rank's own (freshly exchanged) u halo -- the only halo cells of the friction fluxes that the
// next kernel reads -- which makes the reference's halo exchange of (fe, fn) at this point
// (shallow_water.py:372-376) unnecessary; the values are bit-identical to what the neighbour
// would have sent because both sides evaluate the same expression on the same operands.
__device__ __forceinline__ void swe_k3_body(const B2SweParams& p, const float* __restrict__ u,
                                            float* __restrict__ fe, float* __restrict__ fn, int j,
                                            int i0, const bool m[4], bool local_halo,
                                            bool has_south) {
  const int P = p.pitch;
  const Row6 uc = ld_row<false, true>(u, j, i0, P), un = ld_row<false, false>(u, j + 1, i0, P);
  const float U[5] = {uc.c0, uc.c1, uc.c2, uc.c3, uc.e}, UN[4] = {un.c0, un.c1, un.c2, un.c3};
  float FE[4], FN[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const bool mf = m[k] || (local_halo && i0 + k == 0);
    FE[k] = mf ? p.viscosity * (U[k + 1] - U[k]) * p.rdx : 0.f;
    FN[k] = m[k] ? p.viscosity * (UN[k] - U[k]) * p.rdy : 0.f;
    if (p.north_wall && j == p.ny - 2) FN[k] = 0.f;
  }
  const size_t off = (size_t)j * P + i0;
  st4(fe, off, make_float4(FE[0], FE[1], FE[2], FE[3]));
  st4(fn, off, make_float4(FN[0], FN[1], FN[2], FN[3]));
  if (local_halo && has_south && j == 1) {
    const float4 us = ld4(u, (size_t)i0);          // row 0 = south halo of u
    const float US[4] = {us.x, us.y, us.z, us.w};
    float F0[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) F0[k] = m[k] ? p.viscosity * (U[k] - US[k]) * p.rdy : 0.f;
    st4(fn, (size_t)i0, make_float4(F0[0], F0[1], F0[2], F0[3]));
  }
}

__device__ __forceinline__ void swe_k4_body(const B2SweParams& p, float* __restrict__ u,
                                            const float* __restrict__ v, const float* __restrict__ fe,
                                            const float* __restrict__ fn, float* __restrict__ fe2,
                                            float* __restrict__ fn2, int j, int i0, const bool m[4],
                                            SweOut4& o) {
  const int P = p.pitch;
  const size_t off = (size_t)j * P + i0;
  const Row6 fec = ld_row<true, false>(fe, j, i0, P);
  const float4 fnc = ld4(fn, off), fns = ld4(fn, off - P), u4 = ld4(u, off);
  const Row6 vc = ld_row<false, true>(v, j, i0, P);
  const float4 vn = ld4(v, off + P);
  const float FE[5] = {fec.w, fec.c0, fec.c1, fec.c2, fec.c3};
  const float FN[4] = {fnc.x, fnc.y, fnc.z, fnc.w}, FNS[4] = {fns.x, fns.y, fns.z, fns.w};
  const float Uo[4] = {u4.x, u4.y, u4.z, u4.w};
  const float V[5] = {vc.c0, vc.c1, vc.c2, vc.c3, vc.e}, VN[4] = {vn.x, vn.y, vn.z, vn.w};
  float Un[4], FE2[4], FN2[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float un = Uo[k] + p.dt * ((FE[k + 1] - FE[k]) * p.rdx + (FN[k] - FNS[k]) * p.rdy);
    Un[k] = m[k] ? un : Uo[k];
    // NOTE: `v - u` mirrors the reference (examples/shallow_water.py:387-392)
    FE2[k] = m[k] ? p.viscosity * (V[k + 1] - un) * p.rdx : 0.f;
    FN2[k] = m[k] ? p.viscosity * (VN[k] - un) * p.rdy : 0.f;
    if (p.north_wall && j == p.ny - 2) FN2[k] = 0.f;
    o.a[0][k] = FE2[k]; o.a[1][k] = FN2[k];
  }
  st4(u, off, make_float4(Un[0], Un[1], Un[2], Un[3]));
  st4(fe2, off, make_float4(FE2[0], FE2[1], FE2[2], FE2[3]));
  st4(fn2, off, make_float4(FN2[0], FN2[1], FN2[2], FN2[3]));
}

// ---- friction phase helpers, every rounding spelled out (no FMA anywhere: on a host build they
// equal the plain expressions of swe_k34_body / swe_k5_body bit for bit) ---------------------------
__device__ __forceinline__ float swe_visc_flux(float nu, float a, float b, float rd) {   // nu (a - b) / d
  return __fmul_rn(__fmul_rn(nu, __fadd_rn(a, -b)), rd);
}
// x + dt ((fe_c - fe_w) / dx + (fn_c - fn_s) / dy)
__device__ __forceinline__ float swe_apply_div(const B2SweParams& p, float x, float fe_c, float fe_w,
                                               float fn_c, float fn_s) {
  return __fadd_rn(x, __fmul_rn(p.dt, __fadd_rn(__fmul_rn(__fadd_rn(fe_c, -fe_w), p.rdx),
                                                  __fmul_rn(__fadd_rn(fn_c, -fn_s), p.rdy))));
}
// the friction update of u at one cell (5-point stencil), with the reference's boundary rules
__device__ __forceinline__ float swe_friction_u(const B2SweParams& p, float u_c, float u_e, float u_w,
                                                float u_n, float u_s, bool fn_c_zero, bool fn_s_zero) {
  const float fe_c = swe_visc_flux(p.viscosity, u_e, u_c, p.rdx);
  const float fe_w = swe_visc_flux(p.viscosity, u_c, u_w, p.rdx);
  const float fn_c = fn_c_zero ? 0.f : swe_visc_flux(p.viscosity, u_n, u_c, p.rdy);
  const float fn_s = fn_s_zero ? 0.f : swe_visc_flux(p.viscosity, u_c, u_s, p.rdy);
  return swe_apply_div(p, u_c, fe_c, fe_w, fn_c, fn_s);
}

// K3+K4 in one pass: the friction-u fluxes are re-evaluated from u's 5-point stencil instead of
// being written to and re-read from HBM (saves one launch and 5 of 37 array passes per step).
//   fe[c]    = nu (u[c+1] - u[c]) / dx      fe[c-1]  = nu (u[c] - u[c-1]) / dx
//   fn[c]    = nu (u[c+nx] - u[c]) / dy     fn[c-nx] = nu (u[c] - u[c-nx]) / dy
// with the reference's boundary values: fn = 0 on row ny-2 of the north-wall ranks ("v" rule),
// fn's south halo row = 0 on south-wall ranks (never received), u's halo supplying the west /
// south halo fluxes elsewhere -- the same expressions on the same operands as the separate
// K3 -> exchange -> K4 sequence, hence the same bits.
//
// u is read with its 5-point neighbourhood, so the update must NOT be in place (a neighbour's
// thread may already have stored its new value): u -> u_new ping-pong, like h in K2.  The halo
// rows / columns / pad lanes of u_new are copies of u's, i.e. what the in-place update of the
// reference leaves there (u's halo as exchanged before the friction step).
__device__ __forceinline__ void swe_k34_body(const B2SweParams& p, const float* __restrict__ u,
                                             float* __restrict__ u_new,
                                             const float* __restrict__ v, float* __restrict__ fe2,
                                             float* __restrict__ fn2, int j, int i0, const bool m[4],
                                             bool has_south, SweOut4& o) {
  const int P = p.pitch;
  const size_t off = (size_t)j * P + i0;
  const Row6 uc = ld_row<true, true>(u, j, i0, P);
  const float4 un4 = ld4(u, off + P), us4 = ld4(u, off - P);
  const Row6 vc = ld_row<false, true>(v, j, i0, P);
  const float4 vn = ld4(v, off + P);
  const float U[6] = {uc.w, uc.c0, uc.c1, uc.c2, uc.c3, uc.e};
  const float UN[4] = {un4.x, un4.y, un4.z, un4.w}, US[4] = {us4.x, us4.y, us4.z, us4.w};
  const float V[5] = {vc.c0, vc.c1, vc.c2, vc.c3, vc.e}, VN[4] = {vn.x, vn.y, vn.z, vn.w};
  const bool fn_c_zero = p.north_wall && j == p.ny - 2;
  const bool fn_s_zero = (j == 1) && !has_south;
  float Un[4], FE2[4], FN2[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float uk = U[k + 1];
#if B2_SWE_EXPLICIT_ROUNDING
    const float un = swe_friction_u(p, uk, U[k + 2], U[k], UN[k], US[k], fn_c_zero, fn_s_zero);
    Un[k] = m[k] ? un : uk;
    FE2[k] = m[k] ? swe_visc_flux(p.viscosity, V[k + 1], un, p.rdx) : 0.f;
    FN2[k] = m[k] ? swe_visc_flux(p.viscosity, VN[k], un, p.rdy) : 0.f;
#else
    const float fe_c = p.viscosity * (U[k + 2] - uk) * p.rdx;
    const float fe_w = p.viscosity * (uk - U[k]) * p.rdx;
    const float fn_c = fn_c_zero ? 0.f : p.viscosity * (UN[k] - uk) * p.rdy;
    const float fn_s = fn_s_zero ? 0.f : p.viscosity * (uk - US[k]) * p.rdy;
    const float un = uk + p.dt * ((fe_c - fe_w) * p.rdx + (fn_c - fn_s) * p.rdy);
    Un[k] = m[k] ? un : uk;
    // NOTE: `v - u` mirrors the reference (examples/shallow_water.py:387-392)
    FE2[k] = m[k] ? p.viscosity * (V[k + 1] - un) * p.rdx : 0.f;
    FN2[k] = m[k] ? p.viscosity * (VN[k] - un) * p.rdy : 0.f;
#endif
    if (p.north_wall && j == p.ny - 2) FN2[k] = 0.f;
    o.a[0][k] = FE2[k]; o.a[1][k] = FN2[k];
  }
  st4(u_new, off, make_float4(Un[0], Un[1], Un[2], Un[3]));
  st4(fe2, off, make_float4(FE2[0], FE2[1], FE2[2], FE2[3]));
  st4(fn2, off, make_float4(FN2[0], FN2[1], FN2[2], FN2[3]));
  if (j == 1) st4(u_new, (size_t)i0, us4);                                   // south halo row
  if (j == p.ny - 2) st4(u_new, (size_t)(p.ny - 1) * P + i0, un4);         // north halo row
}

__device__ __forceinline__ void swe_k5_body(const B2SweParams& p, float* __restrict__ v,
                                            const float* __restrict__ fe2,
                                            const float* __restrict__ fn2, int j, int i0,
                                            const bool m[4]) {
  const int P = p.pitch;
  const size_t off = (size_t)j * P + i0;
  const Row6 fec = ld_row<true, false>(fe2, j, i0, P);
  const float4 fnc = ld4(fn2, off), fns = ld4(fn2, off - P), v4 = ld4(v, off);
  const float FE[5] = {fec.w, fec.c0, fec.c1, fec.c2, fec.c3};
  const float FN[4] = {fnc.x, fnc.y, fnc.z, fnc.w}, FNS[4] = {fns.x, fns.y, fns.z, fns.w};
  const float Vo[4] = {v4.x, v4.y, v4.z, v4.w};
  float Vn[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
#if B2_SWE_EXPLICIT_ROUNDING
    const float vn = swe_apply_div(p, Vo[k], FE[k + 1], FE[k], FN[k], FNS[k]);
#else
    const float vn = Vo[k] + p.dt * ((FE[k + 1] - FE[k]) * p.rdx + (FN[k] - FNS[k]) * p.rdy);
#endif
    Vn[k] = m[k] ? vn : Vo[k];
  }
  st4(v, off, make_float4(Vn[0], Vn[1], Vn[2], Vn[3]));
}

