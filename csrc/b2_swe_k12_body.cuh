// mpi4jax_b200 -- shallow water, the vectorised BULK kernels of the communication-avoiding step
// (b2_swe_ca.cu): cells at least four away from the block edge, whole float4 groups, no halo;
// 16 array passes per step instead of the stand-alone kernels' 32.
//
//  * swe_k12_body   flux + tendency kernels fused.  The stand-alone step writes fe, fn, q, ke in K1
//                   (4 array passes) only for K2 to read them back (4 more, next to h, u, v a second
//                   time): 23 passes.  Here the four quantities are recomputed at the stencil
//                   neighbours of a cell from h, u, v (rows j-1..j+1) and the Adams-Bashforth
//                   update is applied directly: 12 passes.  u', v' go to their own arrays (the friction
//                   kernel's input), so only h needs a second buffer.
//  * swe_k345_body  friction phase fused (u' -> u and v' -> v in one kernel).  The stand-alone phase
//                   is K34 (u' -> u, writes the friction-v fluxes fe2, fn2), exchange(fe2, fn2), K5
//                   (v' -> v): 9 passes.  A bulk cell's K5 needs fe2 / fn2 only at its own, its west
//                   and its south neighbour, so they are recomputed from u_new at those three cells
//                   (three evaluations of u's 5-point friction stencil instead of one): 4 passes.
//
// Bit-compatibility with the stand-alone kernels: the recomputed quantities are the shared
// explicit-rounding helpers of b2_swe_body.cuh (swe_fe .. swe_ke, swe_k2_cell, swe_friction_u,
// swe_visc_flux, swe_apply_div), same operands, hence the same bits.
#pragma once

#include "b2_swe_body.cuh"

// Row of six with index d + 1 for column offset d = -1 .. 4 relative to the group start.
struct Row6A {
  float a[6];
};
template <bool W, bool E>
__device__ __forceinline__ Row6A ld_row_a(const float* __restrict__ x, int j, int i0, int pitch) {
  const Row6 r = ld_row<W, E>(x, j, i0, pitch);
  Row6A o;
  o.a[0] = r.w; o.a[1] = r.c0; o.a[2] = r.c1; o.a[3] = r.c2; o.a[4] = r.c3; o.a[5] = r.e;
  return o;
}

// K12 for one aligned group of row j, all four lanes at least two cells from the block edge.
__device__ __forceinline__ void swe_k12_body(const B2SweParams& p, const float* __restrict__ h,
                                             float* __restrict__ h_new, const float* __restrict__ u,
                                             float* __restrict__ u_new, const float* __restrict__ v,
                                             float* __restrict__ v_new, const float* dh, const float* du,
                                             const float* dv, float* dh_o, float* du_o, float* dv_o, int j,
                                             int i0) {      // (dh_o may be dh, ...: each cell reads then writes its own)
  const int P = p.pitch;
  const size_t off = (size_t)j * P + i0;
  // index d + 1 <-> column i0 + d
  const Row6A Hm = ld_row_a<false, true>(h, j - 1, i0, P), Hc = ld_row_a<true, true>(h, j, i0, P),
              Hp = ld_row_a<true, true>(h, j + 1, i0, P);
  const Row6A Um = ld_row_a<false, false>(u, j - 1, i0, P), Uc = ld_row_a<true, true>(u, j, i0, P),
              Up = ld_row_a<true, false>(u, j + 1, i0, P);
  const Row6A Vm = ld_row_a<false, true>(v, j - 1, i0, P), Vc = ld_row_a<true, true>(v, j, i0, P),
              Vp = ld_row_a<false, false>(v, j + 1, i0, P);
  float4 dh4 = make_float4(0, 0, 0, 0), du4 = dh4, dv4 = dh4;
  if (!p.first_step) { dh4 = ld4(dh, off); du4 = ld4(du, off); dv4 = ld4(dv, off); }
  const float DHo[4] = {dh4.x, dh4.y, dh4.z, dh4.w}, DUo[4] = {du4.x, du4.y, du4.z, du4.w},
              DVo[4] = {dv4.x, dv4.y, dv4.z, dv4.w};
  const float cor_c = p.coriolis[j], cor_m = p.coriolis[j - 1];

  float FEc[5], FEp[5], FNc[5], FNm[5], Qc[5], Qm[4], KEc[5], KEp[4];
#pragma unroll
  for (int t = 0; t < 5; ++t) {
    // t <-> column i0 + t - 1  (i0-1 .. i0+3)
    FEc[t] = swe_fe(Hc.a[t], Hc.a[t + 1], Uc.a[t]);
    FEp[t] = swe_fe(Hp.a[t], Hp.a[t + 1], Up.a[t]);
    Qc[t] = swe_q(p, cor_c, Vc.a[t + 1], Vc.a[t], Up.a[t], Uc.a[t], Hc.a[t], Hc.a[t + 1], Hp.a[t], Hp.a[t + 1]);
  }
#pragma unroll
  for (int t = 0; t < 5; ++t) {
    const int x = t + 1;                       // columns i0 .. i0+4     (array index d + 1 = t + 1)
    FNc[t] = swe_fn(Hc.a[x], Hp.a[x], Vc.a[x]);
    FNm[t] = swe_fn(Hm.a[x], Hc.a[x], Vm.a[x]);
    KEc[t] = swe_ke(Uc.a[x], Uc.a[x - 1], Vc.a[x], Vm.a[x]);
  }
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const int x = t + 1;                       // columns i0 .. i0+3
    Qm[t] = swe_q(p, cor_m, Vm.a[x + 1], Vm.a[x], Uc.a[x], Um.a[x], Hm.a[x], Hm.a[x + 1], Hc.a[x], Hc.a[x + 1]);
    KEp[t] = swe_ke(Up.a[x], Up.a[x - 1], Vp.a[x], Vc.a[x]);
  }

  float Hn[4], Un[4], Vn[4], DH[4], DU[4], DV[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    SweK2In in;
    in.fe_c = FEc[k + 1]; in.fe_w = FEc[k]; in.fen_c = FEp[k + 1]; in.fen_w = FEp[k];
    in.fn_c = FNc[k]; in.fn_e = FNc[k + 1]; in.fns_c = FNm[k]; in.fns_e = FNm[k + 1];
    in.q_c = Qc[k + 1]; in.q_w = Qc[k]; in.qs_c = Qm[k];
    in.ke_c = KEc[k]; in.ke_e = KEc[k + 1]; in.ken_c = KEp[k];
    in.h_c = Hc.a[k + 1]; in.h_e = Hc.a[k + 2]; in.h_n = Hp.a[k + 1];
    in.u_o = Uc.a[k + 1]; in.v_o = Vc.a[k + 1];
    in.dh_o = DHo[k]; in.du_o = DUo[k]; in.dv_o = DVo[k];
    const SweK2Out o = swe_k2_cell(p, in);
    Hn[k] = o.h; Un[k] = o.u; Vn[k] = o.v;
    DH[k] = o.dh; DU[k] = o.du; DV[k] = o.dv;
  }
  st4(h_new, off, make_float4(Hn[0], Hn[1], Hn[2], Hn[3]));
  st4(u_new, off, make_float4(Un[0], Un[1], Un[2], Un[3]));
  st4(v_new, off, make_float4(Vn[0], Vn[1], Vn[2], Vn[3]));
  st4(dh_o, off, make_float4(DH[0], DH[1], DH[2], DH[3]));
  st4(du_o, off, make_float4(DU[0], DU[1], DU[2], DU[3]));
  st4(dv_o, off, make_float4(DV[0], DV[1], DV[2], DV[3]));
}

// friction phase for one aligned group of row j, all four lanes at least three cells from the edge
__device__ __forceinline__ void swe_k345_body(const B2SweParams& p, const float* __restrict__ u,
                                              float* __restrict__ u_new, const float* __restrict__ v,
                                              float* __restrict__ v_new, int j, int i0) {
  const int P = p.pitch;
  const size_t off = (size_t)j * P + i0;
  // U*[d + 2] <-> column i0 + d, d = -2 .. 4
  float Uc[7], Um[7], Up[7], Umm[7], Vc[7], Vp[7];
  {
    const Row6 r = ld_row<true, true>(u, j, i0, P);
    Uc[0] = i0 >= 2 ? u[off - 2] : 0.f;
    Uc[1] = r.w; Uc[2] = r.c0; Uc[3] = r.c1; Uc[4] = r.c2; Uc[5] = r.c3; Uc[6] = r.e;
    const Row6 m1 = ld_row<true, true>(u, j - 1, i0, P);
    Um[0] = 0.f; Um[1] = m1.w; Um[2] = m1.c0; Um[3] = m1.c1; Um[4] = m1.c2; Um[5] = m1.c3; Um[6] = m1.e;
    const Row6 p1 = ld_row<true, false>(u, j + 1, i0, P);
    Up[0] = 0.f; Up[1] = p1.w; Up[2] = p1.c0; Up[3] = p1.c1; Up[4] = p1.c2; Up[5] = p1.c3; Up[6] = 0.f;
    const float4 m2 = ld4(u, off - 2 * (size_t)P);
    Umm[0] = Umm[1] = 0.f; Umm[2] = m2.x; Umm[3] = m2.y; Umm[4] = m2.z; Umm[5] = m2.w; Umm[6] = 0.f;
    const Row6 vc = ld_row<false, true>(v, j, i0, P);
    Vc[0] = Vc[1] = 0.f; Vc[2] = vc.c0; Vc[3] = vc.c1; Vc[4] = vc.c2; Vc[5] = vc.c3; Vc[6] = vc.e;
    const float4 vp = ld4(v, off + P);
    Vp[0] = Vp[1] = 0.f; Vp[2] = vp.x; Vp[3] = vp.y; Vp[4] = vp.z; Vp[5] = vp.w; Vp[6] = 0.f;
  }
  float UNj[5], UNm[4], FE2[5];
#pragma unroll
  for (int t = 0; t < 5; ++t) {                 // u_new[j][i0 + t - 1]
    const int x = t + 1;
    UNj[t] = swe_friction_u(p, Uc[x], Uc[x + 1], Uc[x - 1], Up[x], Um[x], false, false);
    FE2[t] = swe_visc_flux(p.c_nux, Vc[x + 1], UNj[t]);  // fe2[j][i0 + t - 1]
  }
#pragma unroll
  for (int t = 0; t < 4; ++t) {                 // u_new[j-1][i0 + t]
    const int x = t + 2;
    UNm[t] = swe_friction_u(p, Um[x], Um[x + 1], Um[x - 1], Uc[x], Umm[x], false, false);
  }
  float Un[4], Vn[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float fn2_c = swe_visc_flux(p.c_nuy, Vp[k + 2], UNj[k + 1]);   // fn2[j][i]
    const float fn2_s = swe_visc_flux(p.c_nuy, Vc[k + 2], UNm[k]);       // fn2[j-1][i]
    const float vn = swe_apply_div(p, Vc[k + 2], FE2[k + 1], FE2[k], fn2_c, fn2_s);
    Un[k] = UNj[k + 1];
    Vn[k] = vn;
  }
  st4(u_new, off, make_float4(Un[0], Un[1], Un[2], Un[3]));
  st4(v_new, off, make_float4(Vn[0], Vn[1], Vn[2], Vn[3]));
}
