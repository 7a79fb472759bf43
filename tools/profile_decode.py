"""Small driver for ncu: large-v3-dims engine with random bf16 weights made directly on the GPU (no HF init), one
encode + a few decoder steps.  BW_NO_GRAPH=1 makes every kernel a separate launch for the per-launch duration list."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from thewhisper_b200 import synthetic as S  # noqa: E402
from thewhisper_b200.engine import DecodeOptions, ModelDims, WhisperEngine, pack_weights  # noqa: E402


def random_state_dict(dims: ModelDims, device, seed=0):
    """Shapes of a HF Whisper checkpoint, values N(0, 0.02) (LayerNorm weights 1), generated on the device."""
    g = torch.Generator(device=device).manual_seed(seed)
    D, F, V = dims.d_model, dims.ffn, dims.vocab

    def rn(*shape):
        return torch.randn(*shape, generator=g, device=device, dtype=torch.float32) * 0.02

    sd = {}
    e = "model.encoder."
    sd[e + "conv1.weight"], sd[e + "conv1.bias"] = rn(D, dims.n_mels, 3), rn(D)
    sd[e + "conv2.weight"], sd[e + "conv2.bias"] = rn(D, D, 3), rn(D)
    sd[e + "embed_positions.weight"] = rn(1500, D)
    sd[e + "layer_norm.weight"], sd[e + "layer_norm.bias"] = torch.ones(D, device=device), torch.zeros(D, device=device)

    def attn(p, kbias=False):
        for n in ("q_proj", "k_proj", "v_proj", "out_proj"):
            sd[p + n + ".weight"] = rn(D, D)
            if n != "k_proj":
                sd[p + n + ".bias"] = rn(D)

    def ln(p):
        sd[p + ".weight"], sd[p + ".bias"] = torch.ones(D, device=device), torch.zeros(D, device=device)

    for i in range(dims.enc_layers):
        p = f"{e}layers.{i}."
        attn(p + "self_attn.")
        ln(p + "self_attn_layer_norm")
        ln(p + "final_layer_norm")
        sd[p + "fc1.weight"], sd[p + "fc1.bias"] = rn(F, D), rn(F)
        sd[p + "fc2.weight"], sd[p + "fc2.bias"] = rn(D, F), rn(D)
    d = "model.decoder."
    sd[d + "embed_tokens.weight"] = rn(V, D)
    sd[d + "embed_positions.weight"] = rn(dims.max_target_positions, D)
    ln(d + "layer_norm")
    for i in range(dims.dec_layers):
        p = f"{d}layers.{i}."
        attn(p + "self_attn.")
        attn(p + "encoder_attn.")
        ln(p + "self_attn_layer_norm")
        ln(p + "encoder_attn_layer_norm")
        ln(p + "final_layer_norm")
        sd[p + "fc1.weight"], sd[p + "fc1.bias"] = rn(F, D), rn(F)
        sd[p + "fc2.weight"], sd[p + "fc2.bias"] = rn(D, F), rn(D)
    return sd


def main():
    preset = os.environ.get("BW_PRESET", "large-v3")
    steps = int(os.environ.get("BW_STEPS", "8"))
    A = int(os.environ.get("BW_A", "1"))
    G = int(os.environ.get("BW_G", "1"))  # sequences per audio (beams): the step runs A * G rows, cross K/V shared per audio
    chunk_s = int(os.environ.get("BW_CHUNK_S", "30"))
    dev = torch.device("cuda:0")
    dims = ModelDims.from_hf_config(S.make_hf_config(preset))
    sd = random_state_dict(dims, dev)
    w = pack_weights(sd, dims, sd["model.encoder.embed_positions.weight"], dev)
    del sd
    if chunk_s != 30:
        from thewhisper_b200.engine import interpolate_positions

        w["enc.pos"] = interpolate_positions(w["enc.pos"], chunk_s).to(dev)
    eng = WhisperEngine({}, dims, chunk_length_s=chunk_s, device="cuda:0", max_audios=A, max_beams=G, weights=w)
    g = S.make_generation_config(preset, eos_suppressed=True)
    opts = DecodeOptions(eos_token=S.EOS, pad_token=S.EOS, suppress_tokens=list(g.suppress_tokens), begin_suppress_tokens=list(g.begin_suppress_tokens))
    pcm = np.stack([S.synth_audio(chunk_s, seed=1000 + i) for i in range(A)])
    prompt = np.array([[S.SOT, S.LANG_EN, S.TRANSCRIBE, S.NOTIMESTAMPS]] * (A * G), dtype=np.int32)
    for it in range(2):
        eng.logmel(pcm)
        eng.encode(A)
        eng.decode_begin(prompt, A, G, opts)
        eng.decode_run(3 + steps)
        torch.cuda.synchronize()
    if os.environ.get("BW_TIME"):
        for name, fn in (("logmel", lambda: eng.logmel(pcm)), ("encode", lambda: eng.encode(A))):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                fn()
            e1.record()
            torch.cuda.synchronize()
            print(f"{name}: {e0.elapsed_time(e1) / 5:.3f} ms")
        eng.decode_begin(prompt, A, G, opts)
        eng.decode_run(3)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        eng.decode_run(128)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 128 * 1000
        d, L, V, ffn = dims.d_model, dims.dec_layers, dims.vocab, dims.ffn
        gb = (2.0 * (L * (6 * d * d + 2 * d * ffn) + V * d) + A * 2.0 * L * 2 * eng.S * d + A * G * 2.0 * L * 2 * 68 * d) / 1e9
        print(f"decode step (A={A} G={G} S={eng.S}): {us:.1f} us, {A * 1e6 / us:.0f} tok/s, algorithmic {gb:.2f} GB/step -> {gb / us * 1e6:.0f} GB/s")
    toks, fin, pos = eng.decode_read()
    fl = eng.lib.bw_runtime_flags()
    print("pos", pos, "tokens", toks[0, :12].tolist(), "| cooperative launch", bool(fl & 1), "| programmatic dependent launch", bool(fl & 2))


if __name__ == "__main__":
    main()
