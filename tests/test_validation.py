"""Runtime argument checking of the public ops (``mpi4jax_b200._src.validation.enforce_types``).

Behaviour required by the reference's callers (/root/reference/mpi4jax/_src/validation.py:7-93):
named arguments are checked against a type or a tuple of types, numpy abstract types match Python
scalars by ``issubdtype`` (so ``np.integer`` takes ``3`` and ``np.int16(3)`` but neither ``True``
nor ``3.0``), the message names the argument, the accepted types and the offending type, a tensor
in a static slot earns a hint, and decorating a function with a name it does not have fails early.
"""

import types

import numpy as np
import pytest
import torch

import mpi4jax_b200 as m
from mpi4jax_b200._src.validation import enforce_types


def _noop(_):
    return None


def _sample(payload, count, callback=_noop):
    return payload, count, callback


_sample = enforce_types(count=(int, str), callback=types.FunctionType)(_sample)


@pytest.mark.parametrize("args, kwargs", [
    ((None, 3, _noop), {}),
    (("anything", "three", _noop), {}),
    ((object(),), {"count": 7, "callback": _noop}),
    ((b"bytes", 0), {"callback": lambda v: v}),
])
def test_accepted_calls_pass_through_unchanged(args, kwargs):
    out = _sample(*args, **kwargs)
    assert out[0] is args[0]


@pytest.mark.parametrize("args, kwargs, arg, expected, got", [
    ((0, _noop, _noop), {}, "count", "['int', 'str']", "<class 'function'>"),
    ((0, 1, 2), {}, "callback", "function", "<class 'int'>"),
    ((0,), {"count": 1.5, "callback": _noop}, "count", "['int', 'str']", "<class 'float'>"),
    ((0, "x"), {"callback": "not callable"}, "callback", "function", "<class 'str'>"),
])
def test_rejected_calls_name_argument_expectation_and_actual_type(args, kwargs, arg, expected, got):
    with pytest.raises(TypeError) as info:
        _sample(*args, **kwargs)
    text = str(info.value)
    assert f'unexpected type for argument "{arg}"' in text
    assert f"expected: {expected}, got: {got}" in text


def test_unchecked_arguments_are_left_alone_and_defaults_are_checked_too():
    # `payload` is not named in the decorator: anything goes; an omitted `callback` falls back to
    # its default, which is validated like a passed value
    assert _sample(torch.ones(1), 1)[2] is _noop

    @enforce_types(flag=bool)
    def bad_default(flag=0):
        return flag

    with pytest.raises(TypeError, match='argument "flag"'):
        bad_default()
    assert bad_default(True) is True


@pytest.mark.parametrize("value", [0, -4, np.int8(5), np.uint64(2 ** 40), np.int32(-1)])
def test_numpy_abstract_integer_accepts_python_and_numpy_ints(value):
    @enforce_types(rank=np.integer)
    def takes_rank(rank):
        return rank

    assert takes_rank(value) == value


@pytest.mark.parametrize("value, shown", [(True, "bool"), (2.0, "float"), ("1", "str"), (np.float32(1), "float32"),
                                          (None, "NoneType")])
def test_numpy_abstract_integer_rejects_everything_else(value, shown):
    @enforce_types(rank=np.integer)
    def takes_rank(rank):
        return rank

    with pytest.raises(TypeError) as info:
        takes_rank(value)
    assert "expected: integer" in str(info.value) and shown in str(info.value)


def test_optional_slots_are_written_as_tuples_with_nonetype():
    @enforce_types(comm=(type(None), str))
    def op(x, comm=None):
        return comm

    assert op(1) is None and op(1, comm=None) is None and op(1, "world") == "world"
    with pytest.raises(TypeError, match='argument "comm"'):
        op(1, comm=3)


def test_decorating_with_an_unknown_name_fails_at_definition_time():
    def two_args(a, b):
        return a, b

    with pytest.raises(ValueError) as info:
        enforce_types(c=int)(two_args)
    assert 'got unexpected argument "c"' in str(info.value)
    assert enforce_types(a=int, b=int)(two_args)(1, 2) == (1, 2)        # known names are fine


def test_wrapper_keeps_identity_of_the_function():
    assert _sample.__name__ == "_sample"
    assert _sample.__wrapped__.__code__.co_varnames[:3] == ("payload", "count", "callback")


def test_tensor_in_a_static_slot_gets_the_tracer_hint():
    """The reference tells users who pass a traced value where a static one is needed that an
    'abstract tracer was passed'; the torch counterpart of a tracer is a tensor."""
    @enforce_types(root=int)
    def op(x, root):
        return root

    assert op(None, 2) == 2
    for bad in (torch.tensor(2), torch.zeros(()), torch.ones(3, dtype=torch.int64)):
        with pytest.raises(TypeError) as info:
            op(None, bad)
        assert "abstract tracer was passed" in str(info.value)
        assert 'argument "root"' in str(info.value)


@pytest.mark.parametrize("call, arg", [
    (lambda: m.bcast(torch.ones(2), root=torch.tensor(0)), "root"),
    (lambda: m.send(torch.ones(2), dest=1.5), "dest"),
    (lambda: m.recv(torch.ones(2), source="0"), "source"),
    (lambda: m.send(torch.ones(2), 0, tag=None), "tag"),
    (lambda: m.allreduce(torch.ones(2), op="sum"), "op"),
    (lambda: m.allgather(torch.ones(2), comm="world"), "comm"),
    (lambda: m.reduce(torch.ones(2), m.MPI.SUM, root=True), "root"),
])
def test_public_ops_validate_their_static_arguments(call, arg):
    with pytest.raises(TypeError, match=f'unexpected type for argument "{arg}"'):
        call()
