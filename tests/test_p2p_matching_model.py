"""Property tests of the p2p matching protocol against an MPI-semantics oracle, on the executable
model in tests/_p2p_sim.py (the CUDA implementation it mirrors is exercised by
tests/collective_ops/test_send_and_recv.py on GPUs)."""

import random

import pytest

from ._p2p_sim import ANY_TAG, NSLOT, Fatal, Inbox, Pair


def _oracle(messages, recv_tags):
    """MPI matching: each receive takes the earliest not-yet-received message with its tag."""
    taken, order = set(), []
    for t in recv_tags:
        for k, (tag, _) in enumerate(messages):
            if k not in taken and (t == ANY_TAG or tag == t):
                taken.add(k)
                order.append(k)
                break
        else:
            raise AssertionError("test generator produced an unmatched receive")
    return order


def _run(messages, recv_order, slot_bytes=4):
    """messages: [(tag, nbytes)] sent in order; recv_order: indices in the order they are received
    (each receive uses the message's tag, or ANY_TAG when flagged)."""
    pair = Pair(slot_bytes)
    for tag, nb in messages:
        pair.post_send(tag, nb)
    starts, s = [], 0
    for tag, nb in messages:
        starts.append(s)
        s += pair.nfrag(nb)
    pending = list(recv_order)
    guard = 0
    while pending:
        pair.pump_sender()
        k, use_any = pending[0]
        tag, nb = messages[k]
        if pair.try_recv(ANY_TAG if use_any else tag, nb):
            pending.pop(0)
            guard = 0
        else:
            guard += 1
            assert guard < 4, "deadlock in the model: nothing can progress"
    pair.pump_sender()
    assert not pair.sendq, "sender still blocked after every message was received"
    assert pair.head == s and pair.ooo == 0, (pair.head, s, bin(pair.ooo))
    return [starts.index(seq0) for seq0, _, _ in pair.log]


def test_in_order_streaming_of_large_messages():
    msgs = [(1, 100), (2, 3), (3, 0), (4, 57)]              # 25 + 1 + 1 + 15 fragments of 4 bytes
    assert _run(msgs, [(0, False), (1, False), (2, True), (3, False)]) == [0, 1, 2, 3]


@pytest.mark.parametrize("order", [(2, 0, 1), (1, 2, 0), (2, 1, 0), (0, 2, 1)])
def test_three_small_messages_any_order(order):
    msgs = [(10, 4), (11, 2), (12, 0)]
    assert _run(msgs, [(k, False) for k in order]) == list(order)


def test_same_tag_never_overtaken():
    msgs = [(1, 4), (7, 4), (7, 4), (1, 4)]
    got = _run(msgs, [(1, False), (2, False), (0, False), (3, False)])
    assert got == [1, 2, 0, 3]


def test_small_message_behind_a_multi_fragment_one():
    msgs = [(3, 12), (5, 4)]                                 # 3 fragments, then 1
    assert _run(msgs, [(1, False), (0, False)]) == [1, 0]


def test_streamed_message_cannot_overtake():
    msgs = [(3, 4), (5, 12)]
    with pytest.raises(Fatal, match="cannot overtake"):
        _run(msgs, [(1, False), (0, False)])


def _reachable(msgs, nfr, received, k):
    """Can a tagged receive for message k see it?  Replays the scan of p2p_match on the send order:
    holes (already received, single-fragment) cost one slot, older unreceived messages are hopped
    over, a message whose fragments fill the rest of the window ends the scan."""
    first = min(j for j in range(len(msgs)) if j not in received)
    i = 0
    for j in range(first, len(msgs)):
        if i >= NSLOT:
            return False
        if j in received:
            i += 1
            continue
        if j == k:
            return i == 0 or nfr[k] == 1
        if msgs[j][0] == msgs[k][0]:
            return False                 # an older message with the same tag would be matched instead
        if nfr[j] >= NSLOT - i:
            return False
        i += nfr[j]
    return False


def test_random_schedules_match_the_mpi_oracle():
    rng = random.Random(1234)
    overtakes = 0
    for _ in range(3000):
        n = rng.randint(1, 12)
        msgs = [(rng.randint(0, 3), rng.choice([0, 1, 4, 4, 4, 9, 17, 40])) for _ in range(n)]
        nfr = [Pair(4).nfrag(nb) for _, nb in msgs]
        received, order = set(), []
        while len(received) < n:
            oldest = min(j for j in range(n) if j not in received)
            cands = [(oldest, rng.random() < 0.3)]
            cands += [(k, False) for k in range(oldest + 1, n)
                      if k not in received and _reachable(msgs, nfr, received, k)]
            k, use_any = rng.choice(cands)
            overtakes += k != oldest
            order.append((k, use_any))
            received.add(k)
        recv_tags = [ANY_TAG if a else msgs[k][0] for k, a in order]
        assert _run(msgs, order) == _oracle(msgs, recv_tags), (msgs, order)
    assert overtakes > 1000          # the generator really exercises out-of-order matching


def test_unreachable_message_is_a_deadlock_not_a_wrong_match():
    """More than NSLOT - 1 messages ahead of the head: the receive keeps spinning (the device
    watchdog turns that into a timeout diagnostic), it never matches something else."""
    msgs = [(0, 4)] * NSLOT + [(9, 4)]
    pair = Pair(4)
    for tag, nb in msgs:
        pair.post_send(tag, nb)
    pair.pump_sender()
    assert pair.match(9) is None
    assert pair.try_recv(0, 4)           # consuming the head opens the window
    pair.pump_sender()
    assert pair.match(9) == NSLOT


def test_any_source_receives_respect_per_source_order():
    """ANY_SOURCE with a tag (election scans every source's window): whatever source is picked, the
    message is that source's earliest pending one with the tag; every message is received once."""
    rng = random.Random(99)
    for _ in range(1500):
        nsrc = rng.randint(1, 5)
        inbox = Inbox(nsrc)
        pending = [[] for _ in range(nsrc)]           # per source: (tag, nbytes) still to be received
        for s in range(nsrc):
            for _ in range(rng.randint(0, 6)):        # <= 6 single-fragment messages: all inside the window
                tag, nb = rng.randint(0, 3), rng.choice([0, 1, 4])
                inbox.pairs[s].post_send(tag, nb)
                pending[s].append((tag, nb))
        total = sum(len(q) for q in pending)
        for _ in range(total):
            tags = sorted({t for q in pending for t, _ in q})
            want = ANY_TAG if rng.random() < 0.25 else rng.choice(tags)
            got = inbox.try_recv_any(want)
            assert got is not None, (pending, want)
            src, tag, nb = got
            assert want == ANY_TAG or tag == want
            if want == ANY_TAG:
                assert pending[src][0] == (tag, nb)            # the head of that source's queue
                pending[src].pop(0)
            else:
                k = next(i for i, (t, _) in enumerate(pending[src]) if t == want)
                assert pending[src][k] == (tag, nb)            # earliest with that tag from that source
                pending[src].pop(k)
        assert all(not q for q in pending)
        assert all(p.ooo == 0 and not p.sendq for p in inbox.pairs)


def test_any_source_skips_sources_without_a_match():
    inbox = Inbox(3)
    inbox.pairs[0].post_send(1, 4)
    inbox.pairs[1].post_send(2, 4)
    inbox.pairs[1].post_send(7, 4)
    inbox.pairs[2].post_send(3, 40)                    # streamed message at the head of source 2
    inbox.pairs[2].post_send(7, 40)                    # streamed, not at the head: not available yet
    assert inbox.try_recv_any(7) == (1, 7, 4)          # taken out of order from source 1
    assert inbox.try_recv_any(7) is None               # source 2's tag-7 message is behind a streamed one
    assert inbox.try_recv_any(3) == (2, 3, 40)
    assert inbox.try_recv_any(7) == (2, 7, 40)
