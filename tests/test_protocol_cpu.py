"""The hand-over protocol of the persistent decoder-step kernels under a happens-before simulator (tools/protocol_sim.py):
random CTA schedules, vector clocks; no deadlock, every read sees the intended version, no unordered access."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))


def _clean(programs, G, seed, prep=None):
    import protocol_sim as ps

    sim = ps.Sim(G, seed)
    if prep:
        prep(sim)
    sim.run(programs)
    return sim.errors


def test_decode_mega3_protocol_is_race_free():
    import protocol_sim as ps

    G, D, H, ffn, ns, L = 20, 256, 4, 320, 4, 2

    def prep(sim):
        sim.write(0, "accA", range(D), "zeroA@start")
        sim.write(0, "accB", range(D), "zeroB@start")
        sim.vc = [list(sim.vc[0]) for _ in range(G)]

    for seed in range(2):
        assert _clean(ps.mega3_programs(G, D, H, ffn, ns, L), G, seed, prep) == []


def test_decode_mega_variants_protocol_is_race_free():
    import protocol_sim as ps

    G, D, H, ffn, ns, L = 20, 256, 4, 320, 4, 2
    for p2p, prod in ((False, False), (True, True)):
        assert _clean(ps.mega1_programs(G, D, H, ffn, ns, L, p2p, prod), G, 7) == []


def test_simulator_catches_a_missing_barrier_and_a_wrong_target():
    import protocol_sim as ps

    G, D, H, ffn, ns, L = 20, 256, 4, 320, 4, 2
    P = ps.mega1_programs(G, D, H, ffn, ns, L, True, True)
    for b in range(G):  # drop the barrier between out-proj (C) and LN2 + cross-q (D) of the first layer
        k, out = 0, []
        for op in P[b]:
            if op[0] == "barrier":
                k += 1
                if k == 2:
                    continue
            out.append(op)
        P[b] = out
    assert _clean(P, G, 3) != []
    P = ps.mega1_programs(G, D, H, ffn, ns, L, True, False)
    for b in range(G):
        P[b] = [(op[0], op[1], op[2] + 1) if op[0] == "wait" and op[1][0] == "xq" else op for op in P[b]]
    errs = _clean(P, G, 3)
    assert errs and "DEADLOCK" in errs[-1]
