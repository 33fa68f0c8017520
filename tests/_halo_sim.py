"""Pure-numpy simulation of the two halo-exchange schemes (used by test_halo_equivalence.py)."""
import numpy as np
rng = np.random.default_rng(0)

def neighbors(py, px, PY, PX):
    flat = lambda iy, ix: iy * PX + ix
    nb = dict(south=flat(py-1,px) if py>0 else None, north=flat(py+1,px) if py<PY-1 else None,
              west=flat(py,(px-1)%PX), east=flat(py,(px+1)%PX))
    def diag(dy,dx):
        iy=py+dy
        if not 0<=iy<PY: return None
        return flat(iy,(px+dx)%PX)
    dg = dict(sw=diag(-1,-1), se=diag(-1,1), nw=diag(1,-1), ne=diag(1,1))
    return nb, dg

def ref_exchange(arrs, PY, PX, kind):
    P = PY*PX
    send_order = ("west","north","east","south"); recv_order=("east","south","west","north")
    sidx = dict(south=(1,slice(None)), west=(slice(None),1), north=(-2,slice(None)), east=(slice(None),-2))
    ridx = dict(south=(0,slice(None)), west=(slice(None),0), north=(-1,slice(None)), east=(slice(None),-1))
    for sd, rd in zip(send_order, recv_order):
        msgs = {}
        for r in range(P):
            py,px = divmod(r,PX); nb,_ = neighbors(py,px,PY,PX)
            if nb[sd] is not None:
                msgs[(r, nb[sd])] = arrs[r][sidx[sd]].copy()
        for r in range(P):
            py,px = divmod(r,PX); nb,_ = neighbors(py,px,PY,PX)
            if nb[rd] is not None:
                arrs[r][ridx[rd]] = msgs[(nb[rd], r)]
    for r in range(P):
        py,px = divmod(r,PX)
        if kind=="v" and py==PY-1: arrs[r][-2,:]=0

def new_exchange(arrs, PY, PX, kind):
    P=PY*PX; ny,nx = arrs[0].shape
    snap = [a.copy() for a in arrs]   # all sends read pre-exchange values
    for r in range(P):
        py,px = divmod(r,PX); nb,dg = neighbors(py,px,PY,PX)
        a = arrs[r]
        jlo = 1 if nb["south"] is not None else 0
        jhi = ny-1 if nb["north"] is not None else ny
        north_wall = py==PY-1
        if nb["east"] is not None:
            col = snap[nb["east"]][:,1].copy()       # east neighbour's column 1
            for j in range(jlo,jhi):
                a[j,nx-1] = 0 if (kind=="v" and north_wall and j==ny-2) else col[j]
        if nb["west"] is not None:
            col = snap[nb["west"]][:,nx-2].copy()
            for j in range(jlo,jhi):
                a[j,0] = 0 if (kind=="v" and north_wall and j==ny-2) else col[j]
        if nb["south"] is not None:
            a[0,1:nx-1] = snap[nb["south"]][ny-2,1:nx-1]
        if nb["north"] is not None:
            a[ny-1,1:nx-1] = snap[nb["north"]][1,1:nx-1]
        if dg["sw"] is not None: a[0,0] = snap[dg["sw"]][ny-2,nx-2]
        if dg["se"] is not None: a[0,nx-1] = snap[dg["se"]][ny-2,1]
        if dg["nw"] is not None: a[ny-1,0] = snap[dg["nw"]][1,nx-2]
        if dg["ne"] is not None: a[ny-1,nx-1] = snap[dg["ne"]][1,1]
        if kind=="v" and north_wall:
            for i in range(nx):
                if (1<=i<nx-1) or (i==0 and nb["west"] is None) or (i==nx-1 and nb["east"] is None):
                    a[ny-2,i]=0

