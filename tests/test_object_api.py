"""mpi4py's lower-case object API on our communicators (scripts written for mpi4jax use mpi4py
for rank-0 gathers, parameter broadcasts, ...; cf. /root/reference/docs/sharp-bits.rst:74-135)."""

import pytest

from mpi4jax_b200 import MPI

world = MPI.COMM_WORLD
rank, size = world.Get_rank(), world.Get_size()


@pytest.fixture()
def comm():
    c = world.Clone()          # keep object traffic away from the tensor ops of other tests
    yield c
    c.Free()


def test_collectives_of_python_objects(comm):
    assert comm.bcast({"a": 1, "r": rank} if rank == 0 else None, root=0) == {"a": 1, "r": 0}
    assert comm.allgather(("x", rank)) == [("x", r) for r in range(size)]
    got = comm.gather(rank * 10, root=size - 1)
    assert got == ([r * 10 for r in range(size)] if rank == size - 1 else None)
    assert comm.scatter([f"item{r}" for r in range(size)] if rank == 0 else None, root=0) == f"item{rank}"
    assert comm.allreduce(rank + 1) == size * (size + 1) // 2
    assert comm.allreduce(rank + 1, op=MPI.MAX) == size
    assert comm.allreduce([rank], op=MPI.SUM) == list(range(size))          # list concatenation, like mpi4py
    red = comm.reduce(2, op=MPI.PROD, root=0)
    assert red == (2 ** size if rank == 0 else None)
    comm.barrier()
    with pytest.raises(ValueError):
        if rank == 0:
            comm.scatter([1] * (size + 1), root=0)
        else:
            raise ValueError("only the root validates")


def test_point_to_point_objects(comm):
    comm.send({"hello": rank}, dest=rank, tag=3)                 # self-send
    status = MPI.Status()
    assert comm.recv(source=rank, tag=3, status=status) == {"hello": rank}
    assert status.Get_source() == rank and status.Get_tag() == 3 and status.Get_count() > 0
    if size > 1:
        right, left = (rank + 1) % size, (rank - 1) % size
        assert comm.sendrecv(("from", rank), dest=right, source=left) == ("from", left)
        # tags are matched out of order, like MPI
        comm.send("first", dest=right, tag=1)
        comm.send("second", dest=right, tag=2)
        assert comm.recv(source=left, tag=2) == "second"
        assert comm.recv(source=left, tag=1) == "first"


def test_module_level_helpers():
    assert isinstance(MPI.Get_processor_name(), str) and MPI.Get_processor_name()
    t0 = MPI.Wtime()
    assert MPI.Wtime() >= t0
    assert MPI.Is_initialized() and not MPI.Is_finalized()
    assert MPI.SUM(2, 3) == 5 and MPI.MAX(2, 3) == 3 and MPI.BXOR(6, 3) == 5
