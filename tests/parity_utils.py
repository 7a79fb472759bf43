"""Tie-aware replay of a pipeline run through the oracle (shared by the GPU pipeline tests and the CPU self-check)."""
import numpy as np


class DecodeRecorder:
    """Records every decode call of a pipeline run: the fp32 features the engine encoded, the prompts and the ids it
    generated -- so each call can be replayed through the oracle by teacher forcing."""

    def __init__(self, pipe):
        self.records = []
        self._mel = None
        gen, eng = pipe.generator, pipe.engine
        og, os_, od = gen.generate, eng.set_mel, gen._decode

        def generate(B, **kw):
            self._mel = kw["mel_f32"].float().cpu().numpy()[:B]
            return og(B, **kw)

        def set_mel(m):
            self._mel = m.float().cpu().numpy()
            return os_(m)

        def _decode(prompts, A, opts, max_new, num_beams):
            out = od(prompts, A, opts, max_new, num_beams)
            self.records.append({"mel": self._mel[:A].copy(), "prompts": np.array(prompts), "gen": [np.asarray(g) for g in out[0]],
                                 "eos_seen": list(out[2]), "opts": opts, "max_new": max_new})
            return out

        gen.generate, eng.set_mel, gen._decode = generate, set_mel, _decode


def assert_oracle_greedy(records, om, max_near_ties=3):
    """Every recorded token must be the oracle's processed arg-max given the same prefix (teacher forcing through the HF
    model + oracle/whisper_ref.process_logits), unless the oracle's own decision margin at that step -- top-1 vs top-2
    logit, or the timestamp-probability rule -- is below the bf16 logit tolerance.  Returns the number of such near ties."""
    from oracle import hf_ref, whisper_ref

    near = 0
    for rec in records:
        o = rec["opts"]
        for a in range(len(rec["gen"])):
            prompt = rec["prompts"][a].tolist()
            gen = rec["gen"][a].tolist()
            if rec["eos_seen"][a]:
                gen = gen + [o.eos_token]
            if not gen:
                continue
            plen = len(prompt)
            lg = hf_ref.teacher_forced_logits(om, rec["mel"][a], prompt + gen)
            for i, tok in enumerate(gen):
                row = lg[plen - 1 + i]
                tol = 0.16 * float(row.std()) + 2e-3
                s, pre, rule_margin = whisper_ref.process_logits(
                    row, prompt + gen[:i], plen, suppress=list(o.suppress_tokens), begin_suppress=list(o.begin_suppress_tokens),
                    ts_rules=bool(o.timestamp_rules), ts_begin=o.timestamp_begin, no_ts=o.no_timestamps_token, eos=o.eos_token,
                    max_initial_ts=(o.max_initial_timestamp_index if o.max_initial_timestamp_index >= 0 else None), details=True)
                order = np.argsort(-s)[:2]
                if tok == order[0]:
                    continue
                margin = float(s[order[0]] - s[order[1]])
                ok = tok == order[1] and margin < tol
                if not ok and o.timestamp_rules and abs(rule_margin) < tol:
                    alt = pre.copy()
                    if rule_margin <= 0:  # the rule did not fire in the oracle; it may fire under bf16 noise
                        alt[: o.timestamp_begin] = -np.inf
                    ok = tok == int(np.argmax(alt))
                assert ok, (a, i, tok, order.tolist(), margin, rule_margin, tol)
                near += 1
    assert near <= max_near_ties, near
    return near
