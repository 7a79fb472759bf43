"""A/B of the persistent decoder-step kernel's switches inside ONE process (large-v3 dims, random weights):
BW_MEGA_FLAGS (run-time bits) and BW_MEGA_VARIANT (compile-time variant, decode_mega.cu V_* bits) are re-read at every
decode_begin and each combination gets its own CUDA graph.  BW_AB="flags:variant,..." lists the combinations; the first
is the reference the others' logits and tokens are compared with."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from thewhisper_b200 import synthetic as S  # noqa: E402
from thewhisper_b200.engine import DecodeOptions, ModelDims, WhisperEngine, pack_weights  # noqa: E402
from tools.profile_decode import random_state_dict  # noqa: E402


def main():
    variants = [v for v in os.environ.get("BW_AB", "64:0,64:1,64:3,64:5,64:9,64:7,64:15,192:0").split(",") if v]
    A = int(os.environ.get("BW_A", "1"))
    dev = torch.device("cuda:0")
    dims = ModelDims.from_hf_config(S.make_hf_config(os.environ.get("BW_PRESET", "large-v3")))
    sd = random_state_dict(dims, dev)
    w = pack_weights(sd, dims, sd["model.encoder.embed_positions.weight"], dev)
    del sd
    eng = WhisperEngine({}, dims, chunk_length_s=30, device="cuda:0", max_audios=A, weights=w)
    g = S.make_generation_config("large-v3", eos_suppressed=True)
    opts = DecodeOptions(eos_token=S.EOS, pad_token=S.EOS, suppress_tokens=list(g.suppress_tokens), begin_suppress_tokens=list(g.begin_suppress_tokens))
    pcm = np.stack([S.synth_audio(30, seed=1000 + i) for i in range(A)])
    prompt = np.array([[S.SOT, S.LANG_EN, S.TRANSCRIBE, S.NOTIMESTAMPS]] * A, dtype=np.int32)
    eng.logmel(pcm)
    eng.encode(A)
    torch.cuda.synchronize()
    ref = None
    for v in variants:
        fl, nrep = v.split(":")
        os.environ["BW_MEGA_FLAGS"], os.environ["BW_MEGA_VARIANT"] = fl, nrep
        times = []
        for rep in range(4):
            eng.decode_begin(prompt, A, 1, opts)
            eng.decode_run(3)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            eng.decode_run(128)
            e1.record()
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1) / 128 * 1000)
        toks, _, pos = eng.decode_read()
        lg = eng.logits().float().cpu().numpy()
        if ref is None:
            ref = (toks.copy(), lg.copy())
            same = "reference"
        else:
            same = "tokens %s, logits max|d| %.3g" % ("equal" if np.array_equal(toks, ref[0]) else "DIFFER", float(np.abs(lg - ref[1]).max()))
        print("flags %3s variant %2s: %7.1f us/step (min), %7.1f (median)   pos %d   %s" % (fl, nrep, min(times[1:]), float(np.median(times[1:])), pos, same), flush=True)


if __name__ == "__main__":
    main()
