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


def test_decode_mega_protocol_is_race_free():
    """The shipped protocol: 8 grid barriers per layer (tools/protocol_sim.py also models the round-1 experiments -- per-head
    counters, producer-only arrival, the fused third generation -- which were measured and removed in round 2)."""
    import protocol_sim as ps

    G, D, H, ffn, ns, L = 20, 256, 4, 320, 4, 2
    for seed in (7, 8):
        assert _clean(ps.mega1_programs(G, D, H, ffn, ns, L, False, False), G, seed) == []


def test_simulator_catches_a_missing_barrier():
    import protocol_sim as ps

    G, D, H, ffn, ns, L = 20, 256, 4, 320, 4, 2
    P = ps.mega1_programs(G, D, H, ffn, ns, L, False, False)
    for b in range(G):  # drop the barrier between out-proj (C) and LN2 + cross-q (D) of the first layer
        k, out = 0, []
        for op in P[b]:
            if op[0] == "barrier":
                k += 1
                if k == 4:
                    continue
            out.append(op)
        P[b] = out
    assert _clean(P, G, 3) != []
