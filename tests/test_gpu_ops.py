"""Operator-level GPU tests: each CUDA kernel (through the C ABI) against a plain PyTorch fp32 reference of the
same op.  Tolerances: fp32 outputs see only the 16-bit rounding of the operands they were given (inputs are
pre-rounded, so the fp32 result is exact up to accumulation order); 16-bit outputs add one rounding."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _need_experimental(what):
    """The kernel variants that lost their A/B measurements are compiled only into MMR_EXPERIMENTAL builds
    (csrc/build.py); the default build and test matrix cover what ships."""
    from kddcup_2020_multimodalitiesrecall_2nd_place_b200 import _lib
    if not _lib.load().mmr_experimental_build():
        pytest.skip(f"{what}: only in MMR_EXPERIMENTAL builds")


def _rel(got, ref):
    return ((got.float() - ref.float()).abs().max() / (ref.float().abs().max() + 1e-30)).item()


def _gemm_case(M, N, K, dtype, act=0, residual=False, bias=True, lda_pad=0):
    from kddcup_2020_multimodalitiesrecall_2nd_place_b200 import ops
    torch.manual_seed(0)
    a = torch.randn(M, K + lda_pad, device="cuda").to(dtype)[:, :K]
    w = (torch.randn(N, K, device="cuda") * 0.05).to(dtype)
    b = torch.randn(N, device="cuda") * 0.1 if bias else None
    r = torch.randn(M, N, device="cuda") if residual else None
    o16, o32 = ops.gemm(a, w, b, r, act=act, want16=True, want32=True)
    torch.cuda.synchronize()
    ref = a.float() @ w.float().t()
    if bias:
        ref = ref + b
    ref = {0: lambda x: x, 1: F.relu, 2: lambda x: F.gelu(x, approximate="tanh"), 3: F.gelu, 4: torch.tanh}[act](ref)
    if residual:
        ref = ref + r
    assert _rel(o32, ref) < 2e-5, (M, N, K, act)
    assert _rel(o16, ref) < (4e-3 if dtype == torch.bfloat16 else 6e-4), (M, N, K, act)


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_gemm_shapes_and_tails(dtype):
    _gemm_case(128, 256, 64, dtype, bias=False)
    _gemm_case(256, 768, 768, dtype)
    _gemm_case(300, 272, 128, dtype, lda_pad=8)      # M tail, N tail (272 = 256 + 16), strided A
    _gemm_case(100, 16, 64, dtype)                   # single narrow tile
    _gemm_case(4, 768, 768, dtype, act=4)            # pooler at cfg1 (B=4)


@pytest.mark.parametrize("act", [1, 2, 3, 4])
def test_gemm_fused_activations(act):
    _gemm_case(384, 512, 256, torch.float16, act=act)


def test_gemm_residual_in_place_and_baseline_shapes():
    from kddcup_2020_multimodalitiesrecall_2nd_place_b200 import ops, _lib
    _gemm_case(384, 768, 3072, torch.float16, residual=True)
    # in-place residual update, as the model driver uses it (out32 aliases residual)
    M, N, K = 17408, 768, 3072
    torch.manual_seed(1)
    a = torch.randn(M, K, device="cuda").half()
    w = (torch.randn(N, K, device="cuda") * 0.02).half()
    b = torch.randn(N, device="cuda") * 0.1
    x = torch.randn(M, N, device="cuda")
    ref = a.float() @ w.float().t() + b + x
    lib = _lib.load()
    _lib.check(lib.mmr_gemm(a.data_ptr(), K, w.data_ptr(), K, M, N, K, b.data_ptr(), x.data_ptr(), N, 0, 0,
                            x.data_ptr(), N, 0, _lib.DT_FP16, torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    assert _rel(x, ref) < 2e-5


@pytest.mark.parametrize("M,N,K,act", [(300, 256, 128, 0), (1000, 768, 768, 2), (17408, 2304, 768, 0),
                                       (17408, 3072, 768, 2), (9216, 768, 2048, 1), (8192, 3072, 768, 3),
                                       (19200, 512, 64, 0), (129, 1024, 192, 4)])
@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_gemm_16bit_output_path(M, N, K, act, dtype):
    """out16-only GEMM (QKV / FFN-in class): CTA-pair kernel with the TMA-store epilogue, including the split tail
    wave (N = 2304 -> 2 x 128-wide sub-tiles, N = 3072 -> 4 x 64-wide), ragged M and a strided output view."""
    from kddcup_2020_multimodalitiesrecall_2nd_place_b200 import _lib
    lib = _lib.load()
    torch.manual_seed(M + N + K)
    a = torch.randn(M, K, device="cuda").to(dtype)
    w = (torch.randn(N, K, device="cuda") * 0.05).to(dtype)
    b = torch.randn(N, device="cuda") * 0.1
    ldo = N + 64                                        # output is a column slice of a wider buffer
    out = torch.full((M + 3, ldo), 7.0, device="cuda", dtype=dtype)
    _lib.check(lib.mmr_gemm(a.data_ptr(), K, w.data_ptr(), K, M, N, K, b.data_ptr(), 0, 0, out.data_ptr(), ldo, 0, 0,
                            act, _lib.DT_FP16 if dtype == torch.float16 else _lib.DT_BF16,
                            torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    ref = a.float() @ w.float().t() + b
    ref = {0: lambda x: x, 1: F.relu, 2: lambda x: F.gelu(x, approximate="tanh"), 3: F.gelu, 4: torch.tanh}[act](ref)
    assert _rel(out[:M, :N], ref) < (4e-3 if dtype == torch.bfloat16 else 6e-4)
    # nothing outside [M, N] was touched (TMA stores clip at the tensor-map bounds)
    assert (out[M:] == 7.0).all() and (out[:, N:] == 7.0).all()


def test_gemm_rejects_bad_arguments():
    from kddcup_2020_multimodalitiesrecall_2nd_place_b200 import ops
    from kddcup_2020_multimodalitiesrecall_2nd_place_b200._lib import MmrError
    a = torch.randn(64, 100, device="cuda").half()      # K not a multiple of 64
    w = torch.randn(32, 100, device="cuda").half()
    with pytest.raises(MmrError, match="multiple of 64"):
        ops.gemm(a, w)
    with pytest.raises(MmrError, match="multiple of 16"):
        ops.gemm(torch.randn(64, 64, device="cuda").half(), torch.randn(24, 64, device="cuda").half())


def test_layernorm():
    from kddcup_2020_multimodalitiesrecall_2nd_place_b200 import ops
    torch.manual_seed(1)
    for (M, H) in [(1000, 768), (33, 1536), (1, 768)]:
        x = torch.randn(M, H, device="cuda") * 3 + 0.5
        g = torch.rand(H, device="cuda") + 0.5
        b = torch.randn(H, device="cuda") * 0.1
        o16, o32 = ops.layernorm(x, g, b, dtype=torch.float16)
        torch.cuda.synchronize()
        ref = F.layer_norm(x, (H,), g, b, 1e-12)
        assert _rel(o32, ref) < 1e-6
        assert _rel(o16, ref) < 6e-4
        acc = torch.ones(M, H, device="cuda")
        ops.layernorm(x, g, b, dtype=torch.float16, scale=1.0 / 3.0, accumulate_into=acc)
        torch.cuda.synchronize()
        assert _rel(acc, 1.0 + ref / 3.0) < 1e-6


@pytest.mark.parametrize("kernel", ["tcgen05_pipelined", "tcgen05", "mma_sync", "mma_sync_tma"])
@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_attention(dtype, kernel):
    """All four attention kernels: tcgen05 / TMEM pipelined four items deep with P kept in TMEM, the first tcgen05
    version (one item per CTA, P through shared memory), one CTA per (pair, head) with mma.sync, and the persistent
    TMA-pipelined mma.sync variant."""
    from kddcup_2020_multimodalitiesrecall_2nd_place_b200 import ops, _lib
    if kernel != "tcgen05_pipelined":
        _need_experimental(f"attention kernel '{kernel}'")
    lib = _lib.load()
    default_tc = lib.mmr_get_tuning(_lib.TUNE_ATTN_TC)
    _lib.check(lib.mmr_set_tuning(_lib.TUNE_ATTN_TC, {"tcgen05_pipelined": 2, "tcgen05": 1}.get(kernel, 0)))
    _lib.check(lib.mmr_set_tuning(_lib.TUNE_ATTN_TMA, 1 if kernel == "mma_sync_tma" else 0))
    try:
        _attention_cases(dtype, ops)
    finally:
        _lib.check(lib.mmr_set_tuning(_lib.TUNE_ATTN_TC, default_tc))
        _lib.check(lib.mmr_set_tuning(_lib.TUNE_ATTN_TMA, 0))


def _attention_cases(dtype, ops):
    torch.manual_seed(2)
    H = 12
    for (B, Sq, Sk, masked) in [(3, 68, 68, True), (2, 32, 36, True), (2, 36, 32, False), (2, 104, 104, False),
                                (1, 128, 128, True), (2, 10, 23, True), (2, 1, 1, False), (2, 28, 28, True),
                                (256, 68, 68, True), (300, 36, 32, True)]:
        qkv_q = torch.randn(B * Sq, 3 * 768, device="cuda").to(dtype)
        qkv_k = qkv_q if Sq == Sk else torch.randn(B * Sk, 3 * 768, device="cuda").to(dtype)
        mask = None
        if masked:
            lens = torch.randint(1, Sk + 1, (B,), device="cuda")
            mask = (torch.arange(Sk, device="cuda")[None, :] < lens[:, None]).int().contiguous()
        out = ops.attention(qkv_q[:, :768], qkv_k[:, 768:1536], qkv_k[:, 1536:], mask, B, Sq, Sk, H)
        torch.cuda.synchronize()
        q = qkv_q[:, :768].float().view(B, Sq, H, 64).transpose(1, 2)
        k = qkv_k[:, 768:1536].float().view(B, Sk, H, 64).transpose(1, 2)
        v = qkv_k[:, 1536:].float().view(B, Sk, H, 64).transpose(1, 2)
        s = q @ k.transpose(-1, -2) / 8.0
        if mask is not None:
            s = s + (1.0 - mask.float())[:, None, None, :] * -10000.0      # additive -10000, not -inf
        ref = (torch.softmax(s, -1) @ v).transpose(1, 2).reshape(B * Sq, 768)
        assert _rel(out, ref) < (5e-3 if dtype == torch.bfloat16 else 1.5e-3), (B, Sq, Sk, masked)


def test_attention_fully_masked_row_matches_additive_mask_semantics():
    """Quirk 8: with every key masked the reference's additive -10000 gives a uniform-ish softmax over
    the raw scores, not NaN."""
    from kddcup_2020_multimodalitiesrecall_2nd_place_b200 import ops
    torch.manual_seed(3)
    B, S, H = 1, 16, 12
    qkv = torch.randn(B * S, 3 * 768, device="cuda").half()
    mask = torch.zeros(B, S, dtype=torch.int32, device="cuda")
    out = ops.attention(qkv[:, :768], qkv[:, 768:1536], qkv[:, 1536:], mask, B, S, S, H)
    torch.cuda.synchronize()
    q = qkv[:, :768].float().view(B, S, H, 64).transpose(1, 2)
    k = qkv[:, 768:1536].float().view(B, S, H, 64).transpose(1, 2)
    v = qkv[:, 1536:].float().view(B, S, H, 64).transpose(1, 2)
    ref = (torch.softmax(q @ k.transpose(-1, -2) / 8.0 - 10000.0, -1) @ v).transpose(1, 2).reshape(B * S, 768)
    assert torch.isfinite(out).all() and _rel(out, ref) < 2e-3


def test_cast16_is_round_to_nearest_even():
    from kddcup_2020_multimodalitiesrecall_2nd_place_b200 import ops
    x = torch.randn(36 * 7, 2048, device="cuda")
    for dt in (torch.float16, torch.bfloat16):
        assert torch.equal(ops.cast16(x, dt), x.to(dt))


@pytest.mark.parametrize("variant", [(1, 0), (2, 0), (2, 331), (2, 511), (2, 412), (2, 322)])
@pytest.mark.parametrize("M,K", [(17408, 768), (17408, 3072), (8192, 768), (300, 768), (1000, 3072), (129, 768)])
def test_fused_gemm_layernorm(M, K, variant):
    """One kernel = dense + bias + residual + LayerNorm against the unfused fp32 reference; in place on the residual
    stream, ragged M included.  Variant (1, 0): three CTA pairs per 256-row block meeting through a global table
    (default); (2, cfg): one pair owns the block and all 768 columns, for every shared-memory split it is built in."""
    from kddcup_2020_multimodalitiesrecall_2nd_place_b200 import ops, _lib
    if variant[0] != 1:
        _need_experimental("row-owner GEMM+LayerNorm")
    lib = _lib.load()
    if not lib.mmr_gemm_layernorm_supported(M, K, _lib.DT_FP16):
        pytest.skip("device cannot co-schedule a 6-CTA cluster with this kernel's shared memory")
    _lib.check(lib.mmr_set_tuning(_lib.TUNE_GEMM_LN, variant[0]))
    _lib.check(lib.mmr_set_tuning(_lib.TUNE_LN_ROW_CFG, variant[1]))
    try:
        _fused_gemm_layernorm_case(M, K, ops)
    finally:
        _lib.check(lib.mmr_set_tuning(_lib.TUNE_GEMM_LN, 1))
        _lib.check(lib.mmr_set_tuning(_lib.TUNE_LN_ROW_CFG, 0))


def _fused_gemm_layernorm_case(M, K, ops):
    torch.manual_seed(M + K)
    a = torch.randn(M, K, device="cuda").half()
    w = (torch.randn(768, K, device="cuda") * 0.03).half()
    b = torch.randn(768, device="cuda") * 0.1
    x = torch.randn(M, 768, device="cuda") * 2 + 0.3
    g = torch.rand(768, device="cuda") + 0.5
    be = torch.randn(768, device="cuda") * 0.1
    ref = F.layer_norm(a.float() @ w.float().t() + b + x, (768,), g, be, 1e-12)
    x16, x32 = ops.gemm_layernorm(a, w, b, x, g, be)
    torch.cuda.synchronize()
    assert x32.data_ptr() == x.data_ptr()
    assert _rel(x32, ref) < 1e-5
    assert _rel(x16, ref) < 6e-4
    # every row is normalised: mean ~ beta-weighted, but the pre-affine statistics must be exact
    z = (x32 - be) / g
    assert z.mean(1).abs().max().item() < 1e-4 and (z.var(1, unbiased=False) - 1).abs().max().item() < 1e-3


@pytest.mark.parametrize("knob,value", [("TUNE_GEMM_CLUSTER", 2), ("TUNE_GEMM_TAIL", 0), ("TUNE_GEMM_P16", 0),
                                        ("TUNE_GEMM_PAIR", 0), ("TUNE_GEMM_LN", 0), ("TUNE_PDL", 0)])
def test_alternate_kernel_paths_agree(knob, value):
    """Every kernel-selection knob (mmr_set_tuning) selects a path that computes the same thing: 4-CTA clusters with
    TMA-multicast W (odd and even row-block counts), unsplit tail, the general pair kernel, the single-CTA kernel,
    and GEMM + separate LayerNorm."""
    from kddcup_2020_multimodalitiesrecall_2nd_place_b200 import ops, _lib
    if knob == "TUNE_GEMM_CLUSTER":
        _need_experimental("4-CTA multicast GEMM")
    lib = _lib.load()
    k = getattr(_lib, knob)
    default = {"TUNE_GEMM_CLUSTER": 1}.get(knob, 1)
    torch.manual_seed(5)
    try:
        _lib.check(lib.mmr_set_tuning(k, value))
        for (M, N, K, act) in [(17408, 2304, 768, 0), (19200, 3072, 768, 2), (700, 512, 128, 1)]:
            a = torch.randn(M, K, device="cuda").half()
            w = (torch.randn(N, K, device="cuda") * 0.05).half()
            b = torch.randn(N, device="cuda") * 0.1
            o16, _ = ops.gemm(a, w, b, None, act=act, want16=True, want32=False)
            torch.cuda.synchronize()
            ref = a.float() @ w.float().t() + b
            ref = {0: lambda x: x, 1: F.relu, 2: lambda x: F.gelu(x, approximate="tanh")}[act](ref)
            assert _rel(o16, ref) < 6e-4, (knob, M, N, K)
        if knob == "TUNE_GEMM_LN":
            assert not lib.mmr_gemm_layernorm_supported(17408, 768, _lib.DT_FP16)
    finally:
        _lib.check(lib.mmr_set_tuning(k, default))
