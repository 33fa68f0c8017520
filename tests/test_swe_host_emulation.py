"""The shallow-water stencil bodies compiled for the HOST (tests/native/swe_host_emu.cpp: the
CUDA source with the qualifiers defined away) -- checks, on a machine without a GPU, the launch
sequence of the stand-alone kernels against the public-ops model and the communication-avoiding
step (csrc/b2_swe_ca_body.cuh: frame kernels with owner views, deep-halo exchange geometry,
bulk / frame partition, fused bulk bodies) against the stand-alone kernels, bit for bit."""

import ctypes
import os
import shutil
import subprocess

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CUDA_INC = next((d for d in (os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include"),
                             "/usr/local/cuda/include") if os.path.exists(os.path.join(d, "cuda_runtime.h"))), None)

pytestmark = pytest.mark.skipif(shutil.which("g++") is None or CUDA_INC is None,
                                reason="needs g++ and the CUDA headers")


class Params(ctypes.Structure):       # B2SweParams (csrc/b2_swe_body.cuh)
    _fields_ = [("ny", ctypes.c_int), ("nx", ctypes.c_int), ("pitch", ctypes.c_int),
                ("dx", ctypes.c_float), ("dy", ctypes.c_float), ("dt", ctypes.c_float),
                ("gravity", ctypes.c_float), ("viscosity", ctypes.c_float),
                ("rdx", ctypes.c_float), ("rdy", ctypes.c_float),
                ("ab_a", ctypes.c_float), ("ab_b", ctypes.c_float),
                ("first_step", ctypes.c_int), ("south_wall", ctypes.c_int), ("north_wall", ctypes.c_int),
                ("coriolis", ctypes.c_void_p)] + [(n, ctypes.c_float) for n in
                                                   ("c_gx", "c_gy", "c_nux", "c_nuy", "c_fx", "c_fy")]


def _folded(p):
    """The constant factors the model folds on the host (models/shallow_water.py)."""
    f32 = np.float32
    rdx, rdy, nu, dt, g = f32(p.rdx), f32(p.rdy), f32(p.viscosity), f32(p.dt), f32(p.gravity)
    p.c_gx, p.c_gy = float(-g * rdx), float(-g * rdy)
    p.c_nux, p.c_nuy = float(nu * rdx), float(nu * rdy)
    p.c_fx, p.c_fy = float(dt * nu * rdx * rdx), float(dt * nu * rdy * rdy)
    return p


def _build(tmp_path_factory, name, *defines):
    out = tmp_path_factory.mktemp(name) / f"lib{name}.so"
    cmd = ["g++", "-O1", "-ffp-contract=off", "-std=c++17", "-shared", "-fPIC", *defines, "-I",
           os.path.join(REPO, "csrc"), "-I", CUDA_INC, os.path.join(REPO, "tests", "native", "swe_host_emu.cpp"),
           "-o", str(out)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    return ctypes.CDLL(str(out))


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    lib = _build(tmp_path_factory, "swe_emu")
    from mpi4jax_b200._src import native

    assert ctypes.sizeof(Params) == ctypes.sizeof(native.B2SweParams)
    return lib


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _setup(ny, nx, first_step, south_wall, north_wall, seed=0):
    rng = np.random.default_rng(seed)
    pitch = (nx + 3) // 4 * 4 + 4             # some pad lanes beyond nx
    f = lambda lo, hi: rng.uniform(lo, hi, (ny, pitch)).astype(np.float32)  # noqa: E731
    fields = dict(h=f(90, 110), u=f(-2, 2), v=f(-2, 2), dh=f(-1, 1), du=f(-1, 1), dv=f(-1, 1))
    cor = rng.uniform(1e-4, 2e-4, ny).astype(np.float32)
    p = Params(ny=ny, nx=nx, pitch=pitch, dx=5e3, dy=5e3, dt=30.0, gravity=9.81, viscosity=1e3,
               rdx=np.float32(1) / np.float32(5e3), rdy=np.float32(1) / np.float32(5e3), ab_a=1.5 + 0.1,
               ab_b=-(0.5 + 0.1), first_step=int(first_step), south_wall=int(south_wall),
               north_wall=int(north_wall), coriolis=cor.ctypes.data)
    return _folded(p), fields, cor, rng



# ---- whole steps: the launch sequence of b2_swe_multistep, emulated on a process grid ------------
def _blocks(model, PY, PX):
    """Cut the model's global initial condition into (PY x PX) blocks with one halo cell, as
    ShallowWaterModel does per rank; returns per-rank dicts of pitch-padded float32 arrays."""
    h0, u0, v0 = model.initial_conditions_global()
    NY, NX = h0.shape
    nyi, nxi = (NY - 2) // PY, (NX - 2) // PX
    pitch = (nxi + 2 + 3) // 4 * 4
    ranks = []
    for py in range(PY):
        for px in range(PX):
            blk = {}
            for name, g in (("h", h0), ("u", u0), ("v", v0)):
                a = np.zeros((nyi + 2, pitch), np.float32)
                a[:, :nxi + 2] = g[py * nyi: py * nyi + nyi + 2, px * nxi: px * nxi + nxi + 2]
                blk[name] = a
            for name in ("h1", "u1", "v1", "dh", "du", "dv", "fe", "fn", "q", "ke", "fe2", "fn2"):
                blk[name] = np.zeros((nyi + 2, pitch), np.float32)
            y_global = np.arange(-1, NY - 1) * model.cfg.dy
            cor = (model.cfg.coriolis_f + y_global[py * nyi: py * nyi + nyi + 2] * model.cfg.coriolis_beta)
            blk["cor"] = cor.astype(np.float32)
            ranks.append(blk)
    return ranks, nyi + 2, nxi + 2, pitch


def _params(model, blk, ny, nx, pitch, py, PY, first):
    cfg = model.cfg
    return _folded(Params(ny=ny, nx=nx, pitch=pitch, dx=cfg.dx, dy=cfg.dy, dt=cfg.dt, gravity=cfg.gravity,
                          viscosity=cfg.lateral_viscosity, rdx=np.float32(1) / np.float32(cfg.dx),
                          rdy=np.float32(1) / np.float32(cfg.dy), ab_a=cfg.ab_a, ab_b=cfg.ab_b,
                          first_step=int(first), south_wall=int(py == 0), north_wall=int(py == PY - 1),
                          coriolis=blk["cor"].ctypes.data))


def _exchange(ranks, names, kinds, nx, PY, PX):
    from ._halo_sim import new_exchange

    for name, kind in zip(names, kinds):
        new_exchange([r[name][:, :nx] for r in ranks], PY, PX, kind)


def _emulate(emu, model, PY, PX, nsteps):
    """The launch sequence of b2_swe_multistep: K1 -> exchange -> K2 -> exchange -> K34 -> exchange -> K5."""
    from ._halo_sim import new_exchange

    ranks, ny, nx, pitch = _blocks(model, PY, PX)
    # reset(): halos of the initial state, ping-pong partners start as copies
    for name, kind in (("h", "h"), ("u", "u"), ("v", "v")):
        new_exchange([r[name][:, :nx] for r in ranks], PY, PX, kind)
    for r in ranks:
        r["h1"][:], r["u1"][:], r["v1"][:] = r["h"], r["u"], r["v"]
    hk, hnk = "h", "h1"
    for it in range(nsteps):
        ps = [_params(model, r, ny, nx, pitch, i // PX, PY, it == 0) for i, r in enumerate(ranks)]
        B = ctypes.byref
        for r, p in zip(ranks, ps):
            emu.emu_k1_all(B(p), _ptr(r[hk]), _ptr(r["u"]), _ptr(r["v"]), _ptr(r["fe"]), _ptr(r["fn"]),
                           _ptr(r["q"]), _ptr(r["ke"]))
        _exchange(ranks, ("fe", "fn", "q", "ke"), ("u", "v", "h", "h"), nx, PY, PX)
        for r, p in zip(ranks, ps):
            emu.emu_k2_all(B(p), _ptr(r[hk]), _ptr(r[hnk]), _ptr(r["u"]), _ptr(r["v"]), _ptr(r["dh"]),
                           _ptr(r["du"]), _ptr(r["dv"]), _ptr(r["fe"]), _ptr(r["fn"]), _ptr(r["q"]),
                           _ptr(r["ke"]))
        _exchange(ranks, (hnk, "u", "v"), ("h", "u", "v"), nx, PY, PX)
        for i, (r, p) in enumerate(zip(ranks, ps)):
            emu.emu_k34(B(p), _ptr(r["u"]), _ptr(r["u1"]), _ptr(r["v"]), _ptr(r["fe2"]), _ptr(r["fn2"]),
                        int(i // PX > 0))
            r["u"], r["u1"] = r["u1"], r["u"]
        _exchange(ranks, ("fe2", "fn2"), ("u", "v"), nx, PY, PX)
        for r, p in zip(ranks, ps):
            emu.emu_k5(B(p), _ptr(r["v"]), _ptr(r["fe2"]), _ptr(r["fn2"]))
        hk, hnk = hnk, hk
    return [dict(h=r[hk][:, :nx], u=r["u"][:, :nx], v=r["v"][:, :nx], dh=r["dh"][:, :nx], du=r["du"][:, :nx],
                 dv=r["dv"][:, :nx]) for r in ranks]


def _close(a, b, tol):
    return np.abs(a - b).max() <= tol * (np.abs(b).max() + 1e-30)


def test_emulated_native_step_matches_the_ops_backend(emu):
    """The kernel launch sequence of b2_swe_multistep (bodies + single-phase halo exchange),
    executed on the host, vs the model written with the public ops -- the GPU suite's
    native-vs-ops comparison, on a box without a GPU."""
    import torch

    from mpi4jax_b200.models import ShallowWaterConfig, ShallowWaterModel

    cfg = ShallowWaterConfig(nx=48, ny=24)
    model = ShallowWaterModel(cfg, device="cpu", backend="ops")
    nsteps = 6
    emu_state = _emulate(emu, model, 1, 1, nsteps)[0]
    model.multistep(nsteps)
    for name, t in model.state._asdict().items():
        got, want = emu_state[name], t.numpy()
        assert np.isfinite(got).all()
        tol = 2e-4 if name in ("h", "u", "v") else 2e-3
        assert _close(got[1:-1, 1:-1], want[1:-1, 1:-1], tol), name
    assert isinstance(model.h, torch.Tensor)



# ---- communication-avoiding step (csrc/b2_swe_ca_body.cuh, b2_swe_ca.cu) -----------------------------
class CAExt(ctypes.Structure):        # B2SweCA (csrc/b2_swe_body.cuh)
    _fields_ = [(n, ctypes.c_void_p) for n in ("hx", "upx", "vpx", "uppx", "vppx")] + [
        ("epitch", ctypes.c_int), ("cb1", ctypes.c_int)]


_SIDES = ("w", "e", "s", "n", "sw", "se", "nw", "ne")        # CA_W .. CA_NE: where a message lands


def _neighbours(r, PY, PX):
    py, px = divmod(r, PX)

    def at(dy_, dx_):
        iy = py + dy_
        return None if not 0 <= iy < PY else iy * PX + (px + dx_) % PX

    return dict(w=at(0, -1), e=at(0, 1), s=at(-1, 0), n=at(1, 0), sw=at(-1, -1), se=at(-1, 1), nw=at(1, -1),
                ne=at(1, 1))


def _ca_exchange(emu, ranks, names, ny, nx, pitch, epitch, PY, PX):
    """b2_k_halo_ca without the transport: every message read from the sender's arrays as the push loop
    does and scattered as the poll loop does."""
    F3 = ctypes.c_void_p * 3
    for r, blk in enumerate(ranks):
        nb = _neighbours(r, PY, PX)
        for side, key in enumerate(_SIDES):
            q = nb[key]
            if q is None:
                continue
            send = F3(*[ranks[q][n].ctypes.data for n in names])
            recv = F3(*[blk[n].ctypes.data for n in names])
            ext = F3(blk["hx"].ctypes.data, blk["upx"].ctypes.data, blk["vpx"].ctypes.data)
            emu.emu_ca_deliver(ny, nx, pitch, epitch, side, int(nb["s"] is not None), int(nb["n"] is not None),
                               send, recv, ext)


def _ca_ext(blk, epitch):
    return CAExt(hx=blk["hx"].ctypes.data, upx=blk["upx"].ctypes.data, vpx=blk["vpx"].ctypes.data,
                 uppx=blk["uppx"].ctypes.data, vppx=blk["vppx"].ctypes.data, epitch=epitch, cb1=0)


class EmuStep(ctypes.Structure):      # EmuStep (tests/native/swe_host_emu.cpp): one rank's arrays for one step
    _fields_ = [(n, ctypes.c_void_p) for n in ("h", "h_o", "u", "v", "dh", "du", "dv", "dub", "dvb", "upf", "vpf")]


def _emulate_ca(emu, model, PY, PX, nsteps, reverse=0):
    """The launch sequence of b2_swe_multistep_ca on a process grid: per step the bulk flux+tendency
    kernel, frame kernel A, the bulk friction kernel, the deep exchange, frame kernel D.  Only h is
    double-buffered; u, v and the tendencies are updated in place (kernel A keeps its own du, dv of the
    band-only cells).  ``reverse`` walks every kernel's tasks backwards AND swaps the kernels that run
    concurrently on the device (A before the bulk kernel, D before the bulk friction kernel): the
    result may not depend on either.  (D never runs before the bulk flux+tendency kernel: it overwrites
    u, v of the frame in place, which that kernel's stencil reads -- the B -> D edge of the schedule.)  (Messages are read before any rank scatters: all sends of a step
    come from h', u', v' frame cells, which no message writes.)"""
    from ._halo_sim import new_exchange

    ranks, ny, nx, pitch = _blocks(model, PY, PX)
    epitch = nx + 4
    for name, kind in (("h", "h"), ("u", "u"), ("v", "v")):
        new_exchange([r[name][:, :nx] for r in ranks], PY, PX, kind)
    for r in ranks:
        r["h1"][:] = r["h"]
        for n in ("hx", "upx", "vpx", "uppx", "vppx"):
            r[n] = np.full((ny + 4, epitch), np.nan, np.float32)     # NaN = never delivered / computed
        for n in ("upf", "vpf"):
            r[n] = np.full((ny, pitch), np.nan, np.float32)
        r["dub"], r["dvb"] = r["du"].copy(), r["dv"].copy()          # b2_swe_ca_init
    B = ctypes.byref
    _ca_exchange(emu, ranks, ("h", "u", "v"), ny, nx, pitch, epitch, PY, PX)       # b2_swe_ca_init
    for i, r in enumerate(ranks):
        p = _params(model, r, ny, nx, pitch, i // PX, PY, True)
        x = _ca_ext(r, epitch)
        emu.emu_ca_init_ext(B(p), B(x), _ptr(r["u"]), _ptr(r["v"]))
    hh = ("h", "h1")
    cur = 0
    for it in range(nsteps):
        nxt = cur ^ 1
        ps = [_params(model, r, ny, nx, pitch, i // PX, PY, it == 0) for i, r in enumerate(ranks)]
        xs = [_ca_ext(r, epitch) for r in ranks]
        es = [EmuStep(h=r[hh[cur]].ctypes.data, h_o=r[hh[nxt]].ctypes.data,
                      **{k: r[k].ctypes.data for k in ("u", "v", "dh", "du", "dv", "dub", "dvb", "upf", "vpf")})
              for r in ranks]
        for p, x, e in zip(ps, xs, es):
            if reverse:
                emu.emu_ca_tend_frame(B(p), B(x), B(e), reverse)
                emu.emu_ca_bulk_k12(B(p), B(e))
            else:
                emu.emu_ca_bulk_k12(B(p), B(e))
                emu.emu_ca_tend_frame(B(p), B(x), B(e), reverse)
            if not reverse:
                emu.emu_ca_bulk_fric(B(p), B(e))
        _ca_exchange(emu, ranks, (hh[nxt], "upf", "vpf"), ny, nx, pitch, epitch, PY, PX)
        for p, x, e in zip(ps, xs, es):
            emu.emu_ca_fric_frame(B(p), B(x), B(e), reverse)
            if reverse:
                emu.emu_ca_bulk_fric(B(p), B(e))
        cur = nxt
    return [dict(h=r[hh[cur]][:, :nx], **{k: r[k][:, :nx] for k in ("u", "v", "dh", "du", "dv")}) for r in ranks]


@pytest.mark.parametrize("shape", [(16, 24), (26, 50), (17, 29), (140, 33), (31, 300), (200, 530)])
def test_ca_bulk_and_frame_partition_the_interior(emu, shape):
    ny, nx = shape
    p, *_ = _setup(ny, nx, False, 0, 0)
    assert emu.emu_ca_supported(ctypes.byref(p)) == 1
    marks = np.zeros((ny, nx), np.int32)
    emu.emu_ca_marks(ctypes.byref(p), marks.ctypes.data_as(ctypes.POINTER(ctypes.c_int)))
    owner = marks & 0xff                     # 1 = frame kernels, 16 = bulk kernel (several CTAs)
    assert set(np.unique(owner[1:-1, 1:-1])) <= {1, 16}          # every interior cell written exactly once
    assert marks[0].sum() == marks[-1].sum() == marks[:, 0].sum() == marks[:, -1].sum() == 0
    cb1 = (nx - 4) // 4 * 4                  # east frame: 3 .. 6 columns (the bulk ends on a float4 boundary)
    jj, ii = np.nonzero(owner == 16)
    assert jj.min() == 4 and jj.max() == ny - 5 and ii.min() == 4 and ii.max() == cb1 - 1
    band = (marks & 256) != 0                # kernel A also computes u', v' two cells into the bulk
    want = np.zeros_like(band)
    want[1:-1, 1:-1] = True
    want[6:ny - 6, 6:cb1 - 2] = False
    assert np.array_equal(band, want)


@pytest.mark.parametrize("grid", [(1, 1), (2, 1), (1, 2), (2, 2), (3, 2), (2, 4)])
def test_emulated_ca_pipeline_is_bit_identical_to_the_standalone_pipeline(emu, grid):
    """One deep exchange per step + local recomputation (frame kernels with owner views) vs the three
    exchanges of b2_swe_multistep, same decomposition: the same bits, walls, periodic wrap and the
    reference's stale u / v halos included."""
    from mpi4jax_b200.models import ShallowWaterConfig, ShallowWaterModel

    PY, PX = grid
    model = ShallowWaterModel(ShallowWaterConfig(nx=48 * PX, ny=24 * PY), device="cpu", backend="ops")
    a = _emulate(emu, model, PY, PX, 6)
    b = _emulate_ca(emu, model, PY, PX, 6)
    c = _emulate_ca(emu, model, PY, PX, 6, reverse=1)
    for ra, rb, rc in zip(a, b, c):
        for name in ra:
            assert np.isfinite(rb[name][1:-1, 1:-1]).all(), (grid, name)
            assert np.array_equal(ra[name][1:-1, 1:-1], rb[name][1:-1, 1:-1]), (grid, name)
            assert np.array_equal(rb[name][1:-1, 1:-1], rc[name][1:-1, 1:-1]), (grid, name, "task order")
            # the main arrays' halos as well (h fresh; u, v stale by the friction step)
            if name in ("h", "u", "v"):
                assert np.array_equal(ra[name], rb[name]), (grid, name, "halo")
