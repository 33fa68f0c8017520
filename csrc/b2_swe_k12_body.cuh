// mpi4jax_b200 -- shallow water, flux + tendency kernels fused ("K12"), EXPERIMENTAL / opt-in.
//
// The stand-alone step writes fe, fn, q, ke in K1 (4 array passes) only for K2 to read them back
// (4 more, next to h, u, v a second time): 23 passes for the two kernels, 32 per step.  K12
// recomputes the four quantities at the stencil neighbours of a cell from h, u, v (rows j-1..j+1)
// and applies the Adams-Bashforth update directly: 12 passes, 21 per step.
//
// The reference's discrete system evaluates the fluxes with u, v halos that are stale by the
// friction update and then EXCHANGES fe, fn, q, ke (examples/shallow_water.py:300-330); cells
// whose stencil touches an exchanged flux value therefore cannot be recomputed locally without
// changing the numbers.  So the domain is split:
//   * bulk  = interior cells at distance >= 2 from the block edge: K12, no exchange involved;
//   * frame = the one-cell ring next to the halo: K1 on a frame of width 2 -> the usual exchange of
//     (fe, fn, q, ke) -> K2 on the ring, a few hundred CTAs.
// K12 reads u, v at neighbouring cells, so u and v are ping-ponged like h (cf. the in-place race
// fixed in swe_k34_body).
//
// Bit-compatibility with the two-kernel path: the recomputed quantities use the explicit-rounding
// helpers swe_fe / swe_fn / swe_q / swe_ke of b2_swe_body.cuh and the consumer arithmetic is the
// shared swe_k2_cell below; swe_k1_body / swe_k2_body still use plain expressions whose FMA
// contraction is ptxas' choice (and differs between lanes), so results agree to rounding, not to
// the bit, until those two bodies are switched to the helpers as well (needs a GPU run to confirm).
#pragma once

#include "b2_swe_body.cuh"

__device__ __forceinline__ bool swe_is_bulk(const B2SweParams& p, int j, int i) {
  return j >= 2 && j <= p.ny - 3 && i >= 2 && i <= p.nx - 3;
}
__device__ __forceinline__ bool swe_is_ring(const B2SweParams& p, int j, int i) {
  const bool interior = j >= 1 && j <= p.ny - 2 && i >= 1 && i <= p.nx - 2;
  return interior && !swe_is_bulk(p, j, i);
}

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

// K12 for one aligned group of row j (2 <= j <= ny-3).  Writes ALL four lanes of the six outputs:
// bulk lanes get the update, every other lane a copy of its old value (tendencies of halo / pad
// lanes are zero, as in swe_k2_body) -- the ring lanes are overwritten by swe_k2_ring_body later.
__device__ __forceinline__ void swe_k12_body(const B2SweParams& p, const float* __restrict__ h,
                                             float* __restrict__ h_new, const float* __restrict__ u,
                                             float* __restrict__ u_new, const float* __restrict__ v,
                                             float* __restrict__ v_new, float* __restrict__ dh,
                                             float* __restrict__ du, float* __restrict__ dv, int j, int i0,
                                             const bool m[4]) {
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
    const bool bulk = swe_is_bulk(p, j, i0 + k);
    Hn[k] = bulk ? o.h : in.h_c;
    Un[k] = bulk ? o.u : in.u_o;
    Vn[k] = bulk ? o.v : in.v_o;
    // ring lanes keep their old tendencies (swe_k2_ring_body needs them), halo / pad lanes are zero
    DH[k] = bulk ? o.dh : (m[k] ? DHo[k] : 0.f);
    DU[k] = bulk ? o.du : (m[k] ? DUo[k] : 0.f);
    DV[k] = bulk ? o.dv : (m[k] ? DVo[k] : 0.f);
  }
  st4(h_new, off, make_float4(Hn[0], Hn[1], Hn[2], Hn[3]));
  st4(u_new, off, make_float4(Un[0], Un[1], Un[2], Un[3]));
  st4(v_new, off, make_float4(Vn[0], Vn[1], Vn[2], Vn[3]));
  st4(dh, off, make_float4(DH[0], DH[1], DH[2], DH[3]));
  st4(du, off, make_float4(DU[0], DU[1], DU[2], DU[3]));
  st4(dv, off, make_float4(DV[0], DV[1], DV[2], DV[3]));
}

// K2 on the ring cells of one group (reads the exchanged fe, fn, q, ke like swe_k2_body).  Lanes:
// ring -> update; halo / pad -> copies of the old h, u, v and zero tendencies (what swe_k2_body
// leaves there); bulk -> untouched (read back from the new buffers, written by swe_k12_body).
// Rows 1 / ny-2 also carry u's and v's halo rows over into the new buffers.
__device__ __forceinline__ void swe_k2_ring_body(const B2SweParams& p, const float* __restrict__ h,
                                                 float* __restrict__ h_new, const float* __restrict__ u,
                                                 float* __restrict__ u_new, const float* __restrict__ v,
                                                 float* __restrict__ v_new, float* __restrict__ dh,
                                                 float* __restrict__ du, float* __restrict__ dv,
                                                 const float* __restrict__ fe, const float* __restrict__ fn,
                                                 const float* __restrict__ q, const float* __restrict__ ke,
                                                 int j, int i0, const bool m[4]) {
  const int P = p.pitch;
  const size_t off = (size_t)j * P + i0;
  const Row6 fec = ld_row<true, false>(fe, j, i0, P), fen = ld_row<true, false>(fe, j + 1, i0, P);
  const Row6 fnc = ld_row<false, true>(fn, j, i0, P), fns = ld_row<false, true>(fn, j - 1, i0, P);
  const Row6 qc = ld_row<true, false>(q, j, i0, P), qs = ld_row<false, false>(q, j - 1, i0, P);
  const Row6 kec = ld_row<false, true>(ke, j, i0, P), ken = ld_row<false, false>(ke, j + 1, i0, P);
  const Row6 hc = ld_row<false, true>(h, j, i0, P), hn = ld_row<false, false>(h, j + 1, i0, P);
  const float4 u4 = ld4(u, off), v4 = ld4(v, off);
  const float4 dh4 = ld4(dh, off), du4 = ld4(du, off), dv4 = ld4(dv, off);
  const float4 hb4 = ld4(h_new, off), ub4 = ld4(u_new, off), vb4 = ld4(v_new, off);
  const float FE[5] = {fec.w, fec.c0, fec.c1, fec.c2, fec.c3}, FEN[5] = {fen.w, fen.c0, fen.c1, fen.c2, fen.c3};
  const float FN[5] = {fnc.c0, fnc.c1, fnc.c2, fnc.c3, fnc.e}, FNS[5] = {fns.c0, fns.c1, fns.c2, fns.c3, fns.e};
  const float Q[5] = {qc.w, qc.c0, qc.c1, qc.c2, qc.c3}, QS[4] = {qs.c0, qs.c1, qs.c2, qs.c3};
  const float KE[5] = {kec.c0, kec.c1, kec.c2, kec.c3, kec.e}, KEN[4] = {ken.c0, ken.c1, ken.c2, ken.c3};
  const float H[5] = {hc.c0, hc.c1, hc.c2, hc.c3, hc.e}, HN[4] = {hn.c0, hn.c1, hn.c2, hn.c3};
  const float Uo[4] = {u4.x, u4.y, u4.z, u4.w}, Vo[4] = {v4.x, v4.y, v4.z, v4.w};
  const float DHo[4] = {dh4.x, dh4.y, dh4.z, dh4.w}, DUo[4] = {du4.x, du4.y, du4.z, du4.w},
              DVo[4] = {dv4.x, dv4.y, dv4.z, dv4.w};
  const float Hb[4] = {hb4.x, hb4.y, hb4.z, hb4.w}, Ub[4] = {ub4.x, ub4.y, ub4.z, ub4.w},
              Vb[4] = {vb4.x, vb4.y, vb4.z, vb4.w};
  float Hn[4], Un[4], Vn[4], DH[4], DU[4], DV[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    SweK2In in;
    in.fe_c = FE[k + 1]; in.fe_w = FE[k]; in.fen_c = FEN[k + 1]; in.fen_w = FEN[k];
    in.fn_c = FN[k]; in.fn_e = FN[k + 1]; in.fns_c = FNS[k]; in.fns_e = FNS[k + 1];
    in.q_c = Q[k + 1]; in.q_w = Q[k]; in.qs_c = QS[k];
    in.ke_c = KE[k]; in.ke_e = KE[k + 1]; in.ken_c = KEN[k];
    in.h_c = H[k]; in.h_e = H[k + 1]; in.h_n = HN[k];
    in.u_o = Uo[k]; in.v_o = Vo[k];
    in.dh_o = p.first_step ? 0.f : DHo[k]; in.du_o = p.first_step ? 0.f : DUo[k];
    in.dv_o = p.first_step ? 0.f : DVo[k];
    const SweK2Out o = swe_k2_cell(p, in);
    const bool ring = swe_is_ring(p, j, i0 + k);
    if (ring) {
      Hn[k] = o.h; Un[k] = o.u; Vn[k] = o.v; DH[k] = o.dh; DU[k] = o.du; DV[k] = o.dv;
    } else if (!m[k]) {                                     // halo / pad lane
      Hn[k] = H[k]; Un[k] = Uo[k]; Vn[k] = Vo[k]; DH[k] = DU[k] = DV[k] = 0.f;
    } else {                                                // bulk lane: already final
      Hn[k] = Hb[k]; Un[k] = Ub[k]; Vn[k] = Vb[k]; DH[k] = DHo[k]; DU[k] = DUo[k]; DV[k] = DVo[k];
    }
    // "v" wall rule, applied after the update to every lane of the row (as swe_k2_body does)
    if (p.north_wall && j == p.ny - 2) Vn[k] = 0.f;
  }
  st4(h_new, off, make_float4(Hn[0], Hn[1], Hn[2], Hn[3]));
  st4(u_new, off, make_float4(Un[0], Un[1], Un[2], Un[3]));
  st4(v_new, off, make_float4(Vn[0], Vn[1], Vn[2], Vn[3]));
  st4(dh, off, make_float4(DH[0], DH[1], DH[2], DH[3]));
  st4(du, off, make_float4(DU[0], DU[1], DU[2], DU[3]));
  st4(dv, off, make_float4(DV[0], DV[1], DV[2], DV[3]));
  if (j == 1) {                                             // south halo rows of u, v
    st4(u_new, (size_t)i0, ld4(u, (size_t)i0));
    st4(v_new, (size_t)i0, ld4(v, (size_t)i0));
  }
  if (j == p.ny - 2) {                                      // north halo rows
    const size_t o2 = (size_t)(p.ny - 1) * P + i0;
    st4(u_new, o2, ld4(u, o2));
    st4(v_new, o2, ld4(v, o2));
  }
}

// friction-v update into the ping-pong partner of v (so that v returns to its home buffer after
// the step); arithmetic of swe_k5_body.
__device__ __forceinline__ void swe_k5_pp_body(const B2SweParams& p, const float* __restrict__ v,
                                               float* __restrict__ v_new, const float* __restrict__ fe2,
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
    const float vn = swe_apply_div(p, Vo[k], FE[k + 1], FE[k], FN[k], FNS[k]);
    Vn[k] = m[k] ? vn : Vo[k];
  }
  st4(v_new, off, make_float4(Vn[0], Vn[1], Vn[2], Vn[3]));
  if (j == 1) st4(v_new, (size_t)i0, ld4(v, (size_t)i0));
  if (j == p.ny - 2) {
    const size_t o2 = (size_t)(p.ny - 1) * P + i0;
    st4(v_new, o2, ld4(v, o2));
  }
}

// ---- friction phase fused (u' -> u and v' -> v in one kernel), bulk cells only ----------------
// The stand-alone friction phase is K34 (u' -> u, writes the friction-v fluxes fe2, fn2),
// exchange(fe2, fn2), K5 (v' -> v): 9 array passes.  A bulk cell's K5 needs fe2 / fn2 only at its
// own, its west and its south neighbour, all interior, so they can be recomputed from u_new at
// those three cells (three evaluations of u's 5-point friction stencil instead of one): 4 passes.
// The frame keeps K34 (on a frame of width 2, which also supplies fe2 / fn2 next to the ring)
// -> exchange -> K5 on the ring.

// one aligned group of row j (2 <= j <= ny-3); non-bulk lanes are written as copies of u', v'
__device__ __forceinline__ void swe_k345_body(const B2SweParams& p, const float* __restrict__ u,
                                              float* __restrict__ u_new, const float* __restrict__ v,
                                              float* __restrict__ v_new, int j, int i0, bool has_south) {
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
  const bool fnc0_j = p.north_wall && j == p.ny - 2, fnc0_m = p.north_wall && (j - 1) == p.ny - 2;
  const bool fns0_j = (j == 1) && !has_south, fns0_m = (j - 1 == 1) && !has_south;
  float UNj[5], UNm[4], FE2[5];
#pragma unroll
  for (int t = 0; t < 5; ++t) {                 // u_new[j][i0 + t - 1]
    const int x = t + 1;
    UNj[t] = swe_friction_u(p, Uc[x], Uc[x + 1], Uc[x - 1], Up[x], Um[x], fnc0_j, fns0_j);
    FE2[t] = swe_visc_flux(p.viscosity, Vc[x + 1], UNj[t], p.rdx);  // fe2[j][i0 + t - 1]
  }
#pragma unroll
  for (int t = 0; t < 4; ++t) {                 // u_new[j-1][i0 + t]
    const int x = t + 2;
    UNm[t] = swe_friction_u(p, Um[x], Um[x + 1], Um[x - 1], Uc[x], Umm[x], fnc0_m, fns0_m);
  }
  float Un[4], Vn[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float fn2_c = swe_visc_flux(p.viscosity, Vp[k + 2], UNj[k + 1], p.rdy);   // fn2[j][i]
    const float fn2_s = swe_visc_flux(p.viscosity, Vc[k + 2], UNm[k], p.rdy);       // fn2[j-1][i]
    const float vn = swe_apply_div(p, Vc[k + 2], FE2[k + 1], FE2[k], fn2_c, fn2_s);
    const bool bulk = swe_is_bulk(p, j, i0 + k);
    Un[k] = bulk ? UNj[k + 1] : Uc[k + 2];
    Vn[k] = bulk ? vn : Vc[k + 2];
  }
  st4(u_new, off, make_float4(Un[0], Un[1], Un[2], Un[3]));
  st4(v_new, off, make_float4(Vn[0], Vn[1], Vn[2], Vn[3]));
}

// K5 on the ring cells of one group: ring lanes get the friction-v update from the exchanged
// fe2 / fn2, halo / pad lanes copies of v', bulk lanes stay as swe_k345_body left them; rows 1 and
// ny-2 carry v's halo rows over.
__device__ __forceinline__ void swe_k5_ring_body(const B2SweParams& p, const float* __restrict__ v,
                                                 float* __restrict__ v_new, const float* __restrict__ fe2,
                                                 const float* __restrict__ fn2, int j, int i0,
                                                 const bool m[4]) {
  const int P = p.pitch;
  const size_t off = (size_t)j * P + i0;
  const Row6 fec = ld_row<true, false>(fe2, j, i0, P);
  const float4 fnc = ld4(fn2, off), fns = ld4(fn2, off - P), v4 = ld4(v, off), vb4 = ld4(v_new, off);
  const float FE[5] = {fec.w, fec.c0, fec.c1, fec.c2, fec.c3};
  const float FN[4] = {fnc.x, fnc.y, fnc.z, fnc.w}, FNS[4] = {fns.x, fns.y, fns.z, fns.w};
  const float Vo[4] = {v4.x, v4.y, v4.z, v4.w}, Vb[4] = {vb4.x, vb4.y, vb4.z, vb4.w};
  float Vn[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float vn = swe_apply_div(p, Vo[k], FE[k + 1], FE[k], FN[k], FNS[k]);
    Vn[k] = swe_is_ring(p, j, i0 + k) ? vn : (m[k] ? Vb[k] : Vo[k]);
  }
  st4(v_new, off, make_float4(Vn[0], Vn[1], Vn[2], Vn[3]));
  if (j == 1) st4(v_new, (size_t)i0, ld4(v, (size_t)i0));
  if (j == p.ny - 2) {
    const size_t o2 = (size_t)(p.ny - 1) * P + i0;
    st4(v_new, o2, ld4(v, o2));
  }
}

// ---- frame enumeration --------------------------------------------------------------------
// The frame of width w: rows [1, w] and [ny-1-w, ny-2] completely; of the rows in between, the
// groups that contain the columns [1, w] (group 0 for w <= 2) or [nx-1-w, nx-2].
struct SweFrame {
  int w, ngroups, g_lo, nside, nfull_rows;
  long long total;
};
__host__ __device__ inline SweFrame swe_frame(const B2SweParams& p, int w) {
  SweFrame f;
  f.w = w;
  f.ngroups = p.pitch >> 2;
  f.g_lo = (p.nx - 1 - w) >> 2;
  const int g_hi = (p.nx - 2) >> 2;
  f.nside = 1 + (g_hi - f.g_lo + 1);
  f.nfull_rows = 2 * w;
  f.total = (long long)f.nfull_rows * f.ngroups + (long long)(p.ny - 2 - 2 * w) * f.nside;
  return f;
}
__host__ __device__ inline bool swe_frame_task(const B2SweParams& p, const SweFrame& f, long long idx,
                                               int& j, int& i0, bool m[4]) {
  if (idx >= f.total) return false;
  const long long nfull = (long long)f.nfull_rows * f.ngroups;
  int g;
  if (idx < nfull) {
    const int r = (int)(idx / f.ngroups);
    g = (int)(idx % f.ngroups);
    j = r < f.w ? 1 + r : (p.ny - 1 - f.w) + (r - f.w);
  } else {
    const long long t = idx - nfull;
    const int s = (int)(t % f.nside);
    j = f.w + 1 + (int)(t / f.nside);
    g = s == 0 ? 0 : f.g_lo + (s - 1);
  }
  i0 = g << 2;
  for (int k = 0; k < 4; ++k) m[k] = (i0 + k >= 1) && (i0 + k <= p.nx - 2);
  return true;
}
// smallest block the split works for (distinct west / east frame groups, a non-empty bulk)
__host__ __device__ inline bool swe_k12_supported(const B2SweParams& p) {
  return p.ny >= 8 && p.nx >= 12 && ((p.nx - 3) >> 2) >= 1;
}
