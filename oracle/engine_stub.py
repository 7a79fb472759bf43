"""ORACLE (test infrastructure): a CPU stand-in with the WhisperEngine interface, built from the installed
transformers modules (encoder / decoder forward) plus oracle/whisper_ref.py (rules, DTW).

Purpose: run the product's HOST logic (thewhisper_b200.generation / nvidia.ASRPipeline / streaming) on the CPU box and
compare it end to end with the golden outputs of the real reference, without a GPU.  It is never imported by the
product; `tests/` monkey-patches it in place of the CUDA engine.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np
import torch

from oracle import hf_ref, whisper_ref
from thewhisper_b200.engine import DecodeOptions, ModelDims
from thewhisper_b200.features import HOP, mel_filter_bank


class StubEngine:
    def __init__(self, model, chunk_length_s=30, max_audios=1, max_beams=1, alignment_heads=None, **_):
        self.model = model
        if chunk_length_s < 30 and model.config.max_source_positions == 1500:
            hf_ref.interpolate_positions(model, chunk_length_s)
        self.dims = ModelDims.from_hf_config(model.config)
        self.S = model.config.max_source_positions
        self.frames = 2 * self.S
        self.n_samples = self.frames * HOP
        self.max_audios, self.max_beams = max_audios, max_beams
        self.alignment_heads = [list(p) for p in (alignment_heads or [])]
        self.max_align_steps = 448
        self.bank = mel_filter_bank(self.dims.n_mels)
        self.weights = None
        self.mel = None
        self.enc = None
        self._align = None

    def close(self):
        pass

    # ---- stages -------------------------------------------------------------------------------------------------
    def logmel(self, pcm: np.ndarray, return_f32: bool = False):
        from transformers import WhisperFeatureExtractor

        fe = WhisperFeatureExtractor(feature_size=self.dims.n_mels, chunk_length=self.n_samples // 16000)
        mel = np.stack([hf_ref.logmel(fe, p) for p in pcm])
        self.mel = torch.from_numpy(mel)
        return self.mel.clone() if return_f32 else None

    def set_mel(self, mel: torch.Tensor):
        self.mel = mel.detach().float().cpu()

    @torch.no_grad()
    def encode(self, B: int):
        self.enc = self.model.model.encoder(self.mel[:B]).last_hidden_state

    # ---- decoding -------------------------------------------------------------------------------------------------
    def decode_begin(self, prompts, A, G, opts: DecodeOptions, begin_index=None):
        self._prompts = np.asarray(prompts)
        self._opts = opts
        self._A = A
        self._G = G
        self._plen = self._prompts.shape[1]
        self._seqs = [list(map(int, r)) for r in self._prompts]
        self._slot_rows = []

    @torch.no_grad()
    def decode_beam_step(self, run_scores):
        """full re-forward of every sequence (no cache: test sizes only); log-softmax, rules, per-sequence top 2G"""
        A, G, o = self._A, self._G, self._opts
        ids = torch.tensor(self._seqs, dtype=torch.long)
        enc = self.enc[:A].repeat_interleave(G, dim=0)
        want_align = bool(self.alignment_heads) and getattr(o, "record_alignment", False)
        if want_align:
            self.model.config._attn_implementation = "eager"
        out = self.model.model.decoder(input_ids=ids, encoder_hidden_states=enc, output_attentions=want_align)
        if want_align:  # cross-attention of the newest position of every SLOT at this step (the engine keeps one block per slot)
            row = np.stack([np.stack([out.cross_attentions[l][q, h, -1].float().numpy() for l, h in self.alignment_heads]) for q in range(A * G)])
            if len(self._seqs[0]) > self._plen:  # row 0 = the forward pass that consumed the first generated token
                self._slot_rows.append(row)  # [Q, Ha, S]
        lg = self.model.proj_out(out.last_hidden_state[:, -1]).float()
        lp = torch.log_softmax(lg, dim=-1).numpy()
        K = 2 * G
        cs = np.full((A * G, K), -np.inf, dtype=np.float32)
        ct = np.full((A * G, K), -1, dtype=np.int32)
        for q in range(A * G):
            s = whisper_ref.process_logits(
                lp[q], self._seqs[q], self._plen, suppress=o.suppress_tokens, begin_suppress=o.begin_suppress_tokens,
                ts_rules=o.timestamp_rules, ts_begin=o.timestamp_begin, no_ts=o.no_timestamps_token, eos=o.eos_token,
                max_initial_ts=o.max_initial_timestamp_index if o.max_initial_timestamp_index >= 0 else None)
            order = np.argsort(-s, kind="stable")[:K]
            cs[q] = s[order] + np.float32(run_scores[q])
            ct[q] = np.where(np.isfinite(s[order]), order, -1)
        return cs, ct

    def decode_reorder(self, parent, next_token):
        old = self._seqs
        self._seqs = [list(old[int(p)]) + [int(t)] for p, t in zip(parent, next_token)]

    @torch.no_grad()
    def decode_run(self, n):
        if getattr(self, "_G", 1) > 1:
            return  # beam mode: decode_beam_step re-forwards the whole sequences
        ids = torch.from_numpy(self._prompts).long()
        out = self.model.model.decoder(input_ids=ids, encoder_hidden_states=self.enc[: self._A])
        self._logits = self.model.proj_out(out.last_hidden_state[:, -1]).float()

    def logits(self):
        return self._logits

    @torch.no_grad()
    def greedy(self, prompts: np.ndarray, A: int, opts: DecodeOptions, max_new_tokens: int, poll_every: int = 32):
        from transformers.cache_utils import DynamicCache, EncoderDecoderCache

        plen = prompts.shape[1]
        max_new = max(0, min(max_new_tokens, self.dims.max_target_positions - plen))
        seqs = [list(map(int, r)) for r in prompts]
        finished = [False] * A
        enc = self.enc[:A]
        cache = EncoderDecoderCache(DynamicCache(), DynamicCache())
        self.model.config._attn_implementation = "eager"
        rows: List[List[np.ndarray]] = [[] for _ in range(A)]
        ids = torch.tensor(seqs, dtype=torch.long)
        done = 0
        for step in range(max_new):
            out = self.model.model.decoder(input_ids=ids, encoder_hidden_states=enc, past_key_values=cache, use_cache=True,
                                           output_attentions=True)
            lg = self.model.proj_out(out.last_hidden_state[:, -1]).float().numpy()
            if step > 0 and self.alignment_heads:
                for a in range(A):
                    rows[a].append(np.stack([out.cross_attentions[l][a, h, -1].float().numpy() for l, h in self.alignment_heads]))
            nxt = []
            for a in range(A):
                s = whisper_ref.process_logits(
                    lg[a], seqs[a], plen, suppress=opts.suppress_tokens, begin_suppress=opts.begin_suppress_tokens,
                    ts_rules=opts.timestamp_rules, ts_begin=opts.timestamp_begin, no_ts=opts.no_timestamps_token, eos=opts.eos_token,
                    max_initial_ts=opts.max_initial_timestamp_index if opts.max_initial_timestamp_index >= 0 else None)
                t = int(np.argmax(s))
                if finished[a]:
                    t = opts.pad_token
                elif t == opts.eos_token:
                    finished[a] = True
                seqs[a].append(t)
                nxt.append(t)
            done += 1
            ids = torch.tensor(nxt, dtype=torch.long)[:, None]
            if all(finished):
                break
        self._align = [np.stack(r, axis=1) if r else None for r in rows]  # [Ha, T, S]
        toks = np.full((A, self.dims.max_target_positions), opts.pad_token, dtype=np.int32)
        gen = []
        for a in range(A):
            toks[a, : len(seqs[a])] = seqs[a]
            row = np.asarray(seqs[a][plen:], dtype=np.int32)
            cut = np.where(row == opts.eos_token)[0]
            gen.append(row[: cut[0]] if len(cut) else row)
        return gen, toks, done

    def word_timestamps_batch(self, audios, n_tokens, num_frames, time_precision: float = 0.02) -> np.ndarray:
        out = np.zeros((len(audios), int(max(n_tokens)) + 1), dtype=np.float32)
        for i, (a, t, f) in enumerate(zip(audios, n_tokens, num_frames)):
            out[i, : t + 1] = self.word_timestamps(a, t, f, time_precision)
        return out

    def word_timestamps_gather(self, slot_map, n_tokens, num_frames, time_precision: float = 0.02) -> np.ndarray:
        out = np.zeros((len(n_tokens), int(max(n_tokens)) + 1), dtype=np.float32)
        for i, (t, f) in enumerate(zip(n_tokens, num_frames)):
            w = np.stack([self._slot_rows[r][int(slot_map[i][r])] for r in range(t)], axis=1)  # [Ha, T, S]
            out[i, : t + 1] = whisper_ref.token_timestamps(w, f, time_precision)
        return out

    def word_timestamps(self, audio: int, n_tokens: int, num_frames: int, time_precision: float = 0.02) -> np.ndarray:
        w = self._align[audio][:, :n_tokens]
        return whisper_ref.token_timestamps(w, num_frames, time_precision)
