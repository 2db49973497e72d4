"""Host-side model of the TMEM plan of longconv_tc2_kernel (chimeralm_b200/csrc/longconv_tc2.cuh, header comment).

Two (channel, read-pair) items share the four 128-column TMEM units.  The kernel's safety rests on a static schedule:
who writes which unit in which slot, and what orders a write after the previous reader of the same unit (the in-order
tensor pipe, a barrier the data flow needs anyway, or one of the two early register drains).  This test replays that
schedule for many items and checks, event by event, that a unit is never written while another item's data in it is still
live, and that the orderings the kernel relies on really precede the writes.  It documents the plan; the GPU parity tests
(tests/test_gpu_kernels.py::test_longconv_tensor_core*) are what prove the kernel.  (Mutation check: dropping any one of the
three cross-item waits, or issuing step 7 of the previous item after step 3 instead of before it, makes this test fail.)"""

# slot order of one period (two items: A = item 2k, B = item 2k + 1, B- = item 2k - 1)
MMA_ORDER = [("M1", 0), ("M7", -1), ("M3", 0), ("M1", 1), ("M5", 0), ("M3", 1), ("M7", 0), ("M5", 1)]
EPI_ORDER = [("E3", -1), ("E1", 0), ("E4", -1), ("E2", 0), ("E1", 1), ("E3", 0), ("E2", 1), ("E4", 0)]


def units(step, typ):
    """(reads, writes) of a tensor step / epilogue phase, as unit indices; typ 0 = A-type item, 1 = B-type."""
    im = 2 if typ else 1          # where step 1 puts A_im
    z = 2 if typ else 0           # where step 7 puts z'
    return {
        "M1": ((), (0, im)),      # A_re -> U0, A_im -> U1 / U2
        "E1": ((0, im), (0,)),    # reads A, packs P1 over A_re (U0); A_im is pulled into registers first
        "M3": ((0,), (1, 3)),     # P1 (U0) -> S_re U1, S_im U3
        "E2": ((1, 3), (1,)),     # packs P2 over S_re (U1)
        "M5": ((1,), (2, 3)),     # P2 (U1) -> B_im U2, B_re U3
        "E3": ((2, 3), ()),       # B -> shared memory (BT); B_re (U3) is pulled into registers first
        "M7": ((), (z,)),         # BT -> z'
        "E4": ((z,), ()),         # z' -> registers -> output tile
    }[step]


def schedule(n):
    """Events in program order per role: (role, step, item)."""
    mma, epi = [], []
    for k in range(n // 2 + 1):
        for (st, d) in MMA_ORDER:
            it = 2 * k + d
            if 0 <= it < n:
                mma.append((st, it))
        for (st, d) in EPI_ORDER:
            it = 2 * k + d
            if 0 <= it < n:
                epi.append((st, it))
    return mma, epi


def test_every_item_runs_every_step_once_and_in_order():
    for n in (1, 2, 3, 4, 7, 28, 29):
        mma, epi = schedule(n)
        for it in range(n):
            assert [s for s, i in mma if i == it] == ["M1", "M3", "M5", "M7"], (n, it)
            assert [s for s, i in epi if i == it] == ["E1", "E2", "E3", "E4"], (n, it)


def test_no_unit_is_overwritten_while_another_items_data_is_live():
    """Replay with the kernel's dependencies: an epilogue phase runs after the tensor step that feeds it, a tensor step after
    the epilogue phase that produced its operand; tensor steps complete in issue order.  A write to a unit must come after
    the last read of the previous owner's data in that unit - by one of the mechanisms the kernel uses."""
    feeds = {"E1": "M1", "E2": "M3", "E3": "M5", "E4": "M7"}          # epilogue phase <- tensor step (x/y/x2/o_full)
    operand = {"M3": "E1", "M5": "E2", "M7": "E3"}                     # tensor step <- epilogue phase (p1/p2/bt_full)
    for n in (1, 2, 3, 4, 5, 8, 28):
        mma, epi = schedule(n)
        mma_pos = {ev: i for i, ev in enumerate(mma)}
        epi_pos = {ev: i for i, ev in enumerate(epi)}
        # explicit cross-item waits of the MMA issuer (longconv_tc2.cuh): step -> list of (epilogue event, "drain" | "end")
        def extra_waits(st, it):
            typ, k = it & 1, it >> 1
            w = []
            if st == "M1" and k >= 1:
                w.append((("E4", it - 2), "drain"))                    # e4_done of the previous same-type item
            if st == "M5" and typ == 0:
                if it + 1 < n:
                    w.append((("E1", it + 1), "drain"))                # imd: the B-type partner pulled A_im out of U2
                elif k >= 1:
                    w.append((("E4", it - 1), "drain"))                # no partner: U2 held the previous B-type z'
            if st == "M3" and typ == 1:
                w.append((("E3", it - 1), "drain"))                    # bdr: the A-type partner pulled B_re out of U3
            return w

        # ordered_before(epilogue event e, tensor event m): e is known to be over (or its drain done) when m executes
        def known_done(e, m, kind):
            st, it = m
            # direct: m waits on e (its operand barrier, which fires when the phase is complete, or an extra wait)
            if operand.get(st) and (operand[st], it) == e:
                return True
            if (e, kind) in extra_waits(st, it) or (kind == "drain" and (e, "drain") in extra_waits(st, it)):
                return True
            # transitive through the in-order tensor pipe: an EARLIER tensor step waited for e (complete), or for a
            # later phase of the same epilogue program order
            for m2 in mma[: mma_pos[m]]:
                st2, it2 = m2
                op = operand.get(st2)
                if op and epi_pos[(op, it2)] >= epi_pos[e]:
                    return True
                for (e2, k2) in extra_waits(st2, it2):
                    if epi_pos[e2] > epi_pos[e] or (e2 == e and (k2 == "end" or kind == "drain")):
                        return True
            return False

        readers = {u: [] for u in range(4)}   # events that read unit u since its last write: (role, event)
        # walk the tensor steps in issue order; before each write check all readers of the previous contents
        for m in mma:
            st, it = m
            rd, wr = units(st, it & 1)
            for u in wr:
                for (role, ev) in readers[u]:
                    if ev[1] == it:
                        continue                                   # the item's own earlier phases (data-flow ordered)
                    if role == "mma":
                        assert mma_pos[ev] < mma_pos[m], (n, m, u, ev)   # in-order tensor pipe
                    else:
                        # epilogue readers of another item's data: E1 (A_im), E3 (B_re) and E4 (z') pull the unit into
                        # registers first and signal; the others must be complete
                        drains = ((ev[0] == "E1" and u == (2 if ev[1] & 1 else 1)) or (ev[0] == "E3" and u == 3) or ev[0] == "E4")
                        assert known_done(ev, m, "drain" if drains else "end"), (n, m, u, ev)
                readers[u] = []
            for u in rd:
                readers[u].append(("mma", m))
            # the epilogue phase this step feeds reads (and possibly rewrites in place) the units it touches
            nxt = {v: k for k, v in feeds.items()}.get(st)
            if nxt:
                erd, _ = units(nxt, it & 1)
                for u in erd:
                    readers[u].append(("epi", (nxt, it)))


def test_at_most_four_units_live_at_any_slot_boundary():
    """Unit-count form of the plan: with the two early drains the two items never need a fifth unit."""
    use = [2, 1, 3, 1, 3, 0, 1, 0]      # units an item holds in slots M1 E1 M3 E2 M5 E3 M7 E4, epilogue slots AFTER their drain
    for d in (3, 5):                     # B is three slots behind A, the next A five behind B
        for s in range(8):
            assert use[s] + use[(s - d) % 8] <= 4, (d, s)
