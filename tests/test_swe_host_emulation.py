"""The shallow-water stencil bodies compiled for the HOST (tests/native/swe_host_emu.cpp: the
CUDA source with the qualifiers defined away) -- checks the indexing of the experimental fused
flux+tendency path (K12 on the bulk, K1 -> exchange -> K2 on the frame; csrc/b2_swe_k12_body.cuh)
against K1 -> K2 on every cell, on a machine without a GPU."""

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
                ("coriolis", ctypes.c_void_p)]


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    out = tmp_path_factory.mktemp("emu") / "libswe_emu.so"
    cmd = ["g++", "-O1", "-ffp-contract=off", "-std=c++17", "-shared", "-fPIC", "-I", os.path.join(REPO, "csrc"),
           "-I", CUDA_INC, os.path.join(REPO, "tests", "native", "swe_host_emu.cpp"), "-o", str(out)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    lib = ctypes.CDLL(str(out))
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
    return p, fields, cor, rng


def _fill_flux_halos(arrs, vals):
    """Stand-in for the halo exchange of (fe, fn, q, ke): the same 'received' numbers in both paths."""
    for a, r in zip(arrs, vals):
        a[0, :], a[-1, :] = r[0, :], r[-1, :]
        a[:, 0] = r[:, 0]
        a[:, arrs_nx(a) - 1] = r[:, arrs_nx(a) - 1]


_NX = {}


def arrs_nx(a):
    return _NX[id(a)]


@pytest.mark.parametrize("first_step", [False, True])
@pytest.mark.parametrize("walls", [(0, 0), (1, 0), (0, 1), (1, 1)])
@pytest.mark.parametrize("shape", [(19, 26), (8, 12), (33, 64), (12, 45)])
def test_fused_k12_path_equals_k1_k2(emu, shape, walls, first_step):
    ny, nx = shape
    p, fld, cor, rng = _setup(ny, nx, first_step, *walls, seed=ny * 1000 + nx)
    assert emu.emu_k12_supported(ctypes.byref(p)) == 1
    pitch = p.pitch
    recv = [rng.uniform(-5, 5, (ny, pitch)).astype(np.float32) for _ in range(4)]

    # ---- reference: K1 everywhere -> exchange -> K2 everywhere (u, v in place) -------------
    a = {k: x.copy() for k, x in fld.items()}
    flux = [np.zeros((ny, pitch), np.float32) for _ in range(4)]
    emu.emu_k1_all(ctypes.byref(p), _ptr(a["h"]), _ptr(a["u"]), _ptr(a["v"]), *[_ptr(x) for x in flux])
    for x in flux:
        _NX[id(x)] = nx
    _fill_flux_halos(flux, recv)
    h_ref = np.full((ny, pitch), np.nan, np.float32)
    emu.emu_k2_all(ctypes.byref(p), _ptr(a["h"]), _ptr(h_ref), _ptr(a["u"]), _ptr(a["v"]), _ptr(a["dh"]),
                   _ptr(a["du"]), _ptr(a["dv"]), *[_ptr(x) for x in flux])

    # ---- fused path: K12 on the bulk, K1 on the frame -> exchange -> K2 on the ring -----------
    b = {k: x.copy() for k, x in fld.items()}
    h_new = np.full((ny, pitch), np.nan, np.float32)
    u_new = np.full((ny, pitch), np.nan, np.float32)
    v_new = np.full((ny, pitch), np.nan, np.float32)
    emu.emu_k12_bulk(ctypes.byref(p), _ptr(b["h"]), _ptr(h_new), _ptr(b["u"]), _ptr(u_new), _ptr(b["v"]),
                     _ptr(v_new), _ptr(b["dh"]), _ptr(b["du"]), _ptr(b["dv"]))
    flux2 = [np.full((ny, pitch), np.nan, np.float32) for _ in range(4)]      # NaN = never computed
    emu.emu_k1_frame(ctypes.byref(p), _ptr(b["h"]), _ptr(b["u"]), _ptr(b["v"]), *[_ptr(x) for x in flux2])
    for x in flux2:
        _NX[id(x)] = nx
    _fill_flux_halos(flux2, recv)
    emu.emu_k2_ring(ctypes.byref(p), _ptr(b["h"]), _ptr(h_new), _ptr(b["u"]), _ptr(u_new), _ptr(b["v"]),
                    _ptr(v_new), _ptr(b["dh"]), _ptr(b["du"]), _ptr(b["dv"]), *[_ptr(x) for x in flux2])

    # inputs are untouched (ping-pong), outputs complete
    assert np.array_equal(b["h"], fld["h"]) and np.array_equal(b["u"], fld["u"]) and np.array_equal(b["v"], fld["v"])
    rows = slice(1, ny - 1)
    new = dict(h=h_new, u=u_new, v=v_new, dh=b["dh"], du=b["du"], dv=b["dv"])
    ref = dict(h=h_ref, u=a["u"], v=a["v"], dh=a["dh"], du=a["du"], dv=a["dv"])
    ring = np.zeros((ny, pitch), bool)
    ring[1:ny - 1, 1:nx - 1] = True
    ring[2:ny - 2, 2:nx - 2] = False
    bulk = np.zeros((ny, pitch), bool)
    bulk[2:ny - 2, 2:nx - 2] = True
    ncol = ((nx - 2) // 4 + 1) * 4           # groups without an interior lane are never touched
    for k in new:
        assert not np.isnan(new[k][rows, :ncol]).any(), k
        # frame cells run the very same arithmetic on the same numbers
        assert np.array_equal(new[k][ring], ref[k][ring]), k
        # bulk cells: recomputed fluxes (explicit fma placement) vs stored ones: rounding-level agreement
        scale = np.abs(ref[k][bulk]).max() + 1e-30
        assert np.abs(new[k][bulk] - ref[k][bulk]).max() <= 2e-6 * scale, k
        # halo / pad lanes of the processed rows: exactly what swe_k2_body leaves there
        outside = ~(ring | bulk)
        outside[0, :] = outside[-1, :] = False
        outside[:, ncol:] = False
        assert np.array_equal(new[k][outside], ref[k][outside]), k
    # u's and v's halo rows travel with the ping-pong
    for k, old in (("u", fld["u"]), ("v", fld["v"])):
        assert np.array_equal(new[k][0, :ncol], old[0, :ncol]) and np.array_equal(new[k][-1, :ncol], old[-1, :ncol])


@pytest.mark.parametrize("w", [1, 2])
@pytest.mark.parametrize("shape", [(19, 26), (8, 12), (33, 64), (12, 45), (9, 13)])
def test_frame_enumeration_is_exact(emu, shape, w):
    ny, nx = shape
    p, *_ = _setup(ny, nx, False, 0, 0)
    ngroups = p.pitch // 4
    marks = np.zeros((ny, ngroups), np.int32)
    total = emu.emu_frame_marks(ctypes.byref(p), w, marks.ctypes.data_as(ctypes.POINTER(ctypes.c_int)))
    assert marks.max() == 1 and marks.sum() == total          # no task twice
    want = np.zeros((ny, ngroups), bool)
    for j in range(1, ny - 1):
        for i in range(1, nx - 1):
            if j <= w or j >= ny - 1 - w or i <= w or i >= nx - 1 - w:
                want[j, i // 4] = True
    assert np.array_equal(marks.astype(bool) & want, want)     # every frame cell is covered
    extra = marks.astype(bool) & ~want
    assert not extra[w + 1:ny - 1 - w].any()                   # side rows: frame groups only


def test_friction_v_ping_pong_equals_in_place(emu):
    ny, nx = 17, 30
    p, fld, cor, rng = _setup(ny, nx, False, 0, 1)
    fe2 = rng.uniform(-1, 1, (ny, p.pitch)).astype(np.float32)
    fn2 = rng.uniform(-1, 1, (ny, p.pitch)).astype(np.float32)
    v_ref = fld["v"].copy()
    emu.emu_k5(ctypes.byref(p), _ptr(v_ref), _ptr(fe2), _ptr(fn2))
    v_new = np.full_like(v_ref, np.nan)
    emu.emu_k5_pp(ctypes.byref(p), _ptr(fld["v"]), _ptr(v_new), _ptr(fe2), _ptr(fn2))
    ncol = ((nx - 2) // 4 + 1) * 4           # the pure pad group is never touched
    assert np.array_equal(v_new[:, :ncol], v_ref[:, :ncol])
