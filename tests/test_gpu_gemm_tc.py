"""GPU: the tcgen05/TMA dense transform (FITGNN_GEMM_BF16X3) against an fp64 product of the ORIGINAL fp32 operands.
Tolerance: the path's 1e-3 relative bound with a 10x margin (bf16 hi/lo split keeps ~2^-17 per operand)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


@pytest.fixture(scope="module")
def fg():
    import fitgnn_b200
    return fitgnn_b200


def run_tc(fg, A, W, b, act=0, head=0):
    K = A.shape[1]
    kp = (K + 7) // 8 * 8
    a = fg.ops.split_bf16(A.to(DEV), ldo=kp)
    w = fg.ops.split_bf16(W.to(DEV), ldo=kp)
    out = fg.ops.gemm_bias_act(a, w, None if b is None else b.to(DEV), act, head, precision=fg.ops.GEMM_BF16X3, K=kp)
    torch.cuda.synchronize()
    return out.detach().cpu().numpy().astype(np.float64)


@pytest.mark.parametrize("M,K,N", [(128, 64, 16), (1, 8, 16), (5, 3, 7), (129, 72, 130), (300, 100, 512), (1000, 512, 512),
                                   (4097, 1433, 512), (20000, 512, 512), (333, 512, 256), (640, 200, 48), (2000, 64, 96)])
def test_bf16x3_gemm_matches_fp64(fg, M, K, N):
    g = torch.Generator().manual_seed(M * 7 + K * 3 + N)
    A, W, b = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g) / K ** 0.5, torch.randn(N, generator=g)
    want = (A.double() @ W.double().T + b.double()).numpy()
    got = run_tc(fg, A, W, b)
    scale = np.abs(want).max()
    assert np.abs(got - want).max() <= 1e-4 * scale, np.abs(got - want).max() / scale
    got = run_tc(fg, A, W, b, act=fg.ops.ACT_ELU)
    want_elu = np.where(want > 0, want, np.expm1(np.minimum(want, 0)))
    assert np.abs(got - want_elu).max() <= 1e-4 * scale


@pytest.mark.parametrize("M,K,N", [(40000, 100, 512), (80000, 64, 256), (50001, 128, 512), (160000, 40, 48)])
def test_bf16x3_w_stationary_plan(fg, M, K, N):
    """Small K and many row blocks: the kernel keeps its weight block resident in shared memory (W-stationary)."""
    g = torch.Generator().manual_seed(M + K + N)
    A, W, b = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g) / K ** 0.5, torch.randn(N, generator=g)
    want = (A.double() @ W.double().T + b.double()).numpy()
    got = run_tc(fg, A, W, b)
    scale = np.abs(want).max()
    assert np.abs(got - want).max() <= 1e-4 * scale, np.abs(got - want).max() / scale


@pytest.mark.parametrize("M,K,N,head", [(777, 512, 47, 1), (1000, 512, 7, 1), (300, 64, 3, 2), (2500, 512, 256, 1), (64, 32, 1, 0)])
def test_bf16x3_head(fg, M, K, N, head):
    g = torch.Generator().manual_seed(M + N)
    A, W, b = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g) / K ** 0.5, torch.randn(N, generator=g)
    logits = A.double() @ W.double().T + b.double()
    want = {0: logits, 1: torch.log_softmax(logits, 1), 2: torch.softmax(logits, 1)}[head].numpy()
    got = run_tc(fg, A, W, b, head=head)
    assert np.abs(got - want).max() <= 1e-4 * max(1.0, np.abs(want).max())


def test_bf16x3_engine_matches_fp32_engine(fg):
    """Whole forward with tensor-core GEMMs vs exact-fp32 GEMMs on the same pack (both orders of layer 0)."""
    from oracle import fitgnn_oracle as fo
    for F in (100, 600):
        n = 6000
        ei = fg.synth.powerlaw_graph(n, 15000, seed=1)
        partition, comps, C_list = fg.synth.neighborhood_partition(ei, n, 0.3, seed=1)
        X = fg.synth.features(n, F, seed=1).to(DEV)
        sd = fo.init_state_dict(F, 512, 47, seed=1)
        pack = fg.build_pack(torch.tensor(ei, device=DEV), torch.tensor(partition.part), partition.k, "extra")
        a = fg.PackedForward(pack, sd, precision="fp32")(X).detach().cpu().numpy()
        b = fg.PackedForward(pack, sd, precision="bf16x3")(X).detach().cpu().numpy()
        assert np.abs(a - b).max() <= 1e-4 * np.abs(a).max(), np.abs(a - b).max()


@pytest.mark.parametrize("M,K,N", [(128, 64, 64), (1000, 512, 512), (4097, 104, 512), (333, 256, 40), (50, 64, 256)])
def test_bf16x3_split_output(fg, M, K, N):
    """Epilogue emitting bf16 hi/lo planes (the A operand of the next tensor-core GEMM): hi + lo == the fp32 result."""
    g = torch.Generator().manual_seed(M + K + N)
    A, W, b = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g) / K ** 0.5, torch.randn(N, generator=g)
    kp = (K + 7) // 8 * 8
    a = fg.ops.split_bf16(A.to(DEV), ldo=kp)
    w = fg.ops.split_bf16(W.to(DEV), ldo=kp)
    ref = fg.ops.gemm_bias_act(a, w, b.to(DEV), fg.ops.ACT_ELU, precision=fg.ops.GEMM_BF16X3, K=kp)
    hi, lo = fg.ops.gemm_bias_act(a, w, b.to(DEV), fg.ops.ACT_ELU, precision=fg.ops.GEMM_BF16X3, K=kp, split_out=True)
    torch.cuda.synchronize()
    want_hi = ref.to(torch.bfloat16)
    assert torch.equal(hi, want_hi)
    assert torch.equal(lo, (ref - want_hi.float()).to(torch.bfloat16))


@pytest.mark.parametrize("n,F,H,mode", [(60000, 100, 512, "none"), (40000, 64, 512, "extra"), (45000, 128, 512, "extra")])
def test_fused_gcn_layer_matches_spmm_plus_gemm(fg, n, F, H, mode):
    """fitgnn_gcn_layer_fused (gather warps feeding the tensor-core GEMM) vs the two-kernel path and vs fp64."""
    from oracle import fitgnn_oracle as fo
    ei, part, cw, k = fg.synth.planted_partition(n, 4 * n, 0.4, seed=n, device=DEV, locality=16)
    part = fg.synth.relabel_partition_reference_order(part)
    pack = fg.build_pack(ei, part, k, mode)
    g = torch.Generator().manual_seed(F)
    X = torch.rand(n, F, generator=g).to(DEV)
    W = (torch.randn(H, F, generator=g) / F ** 0.5).to(DEV)
    b = torch.randn(H, generator=g).to(DEV)
    kp = (F + 7) // 8 * 8
    Xp = torch.zeros(n, kp, device=DEV); Xp[:, :F] = X
    Wp = fg.ops.split_bf16(W, ldo=kp)
    rows = pack.core_rows
    fused = fg.ops.gcn_layer_fused(pack.rowptr, pack.col, pack.dinv, Xp, kp, Wp, b, fg.ops.ACT_ELU, src_index=pack.gid,
                                   out_rows=rows)
    A = fg.ops.spmm_symnorm(pack.rowptr, pack.col, pack.dinv, Xp, kp, pack.gid, out_rows=rows, split=True)
    two = fg.ops.gemm_bias_act(A, Wp, b, fg.ops.ACT_ELU, precision=fg.ops.GEMM_BF16X3, K=kp)
    torch.cuda.synchronize()
    assert torch.equal(fused, two)  # same arithmetic, same order -> bit-identical
    # fp64 check on a sample of rows
    sel = torch.randperm(rows.numel(), generator=torch.Generator().manual_seed(1))[:2000]
    Ad = fg.ops.spmm_symnorm(pack.rowptr, pack.col, pack.dinv, Xp, kp, pack.gid, out_rows=rows[sel.to(DEV)].contiguous())
    want = Ad.double().cpu() @ torch.nn.functional.pad(W, (0, kp - F)).double().cpu().T + b.double().cpu()
    want = torch.where(want > 0, want, torch.expm1(want))
    got = fused[sel.to(DEV)].double().cpu()
    assert (got - want).abs().max() <= 1e-4 * want.abs().max()
    hi, lo = fg.ops.gcn_layer_fused(pack.rowptr, pack.col, pack.dinv, Xp, kp, Wp, b, fg.ops.ACT_ELU, src_index=pack.gid,
                                    out_rows=rows, split_out=True)
    assert torch.equal(hi, fused.to(torch.bfloat16))
    # the engine's opt-in use of the fused layer gives the same logits as the default schedule
    sd = fo.init_state_dict(F, H, 7, seed=3)
    a = fg.PackedForward(pack, sd, precision="bf16x3", fuse_aggregate=False)(X)
    f = fg.PackedForward(pack, sd, precision="bf16x3", fuse_layer0=True, fuse_aggregate=False)
    bq = f(X)
    assert f.fused_layer0 is True and torch.equal(a, bq)
    # ineligible shapes are refused loudly at the ABI (callers fall back)
    with pytest.raises(fg._lib.FitgnnError):
        fg.ops.gcn_layer_fused(pack.rowptr, pack.col, pack.dinv, Xp[:1000], kp, Wp, b, out_rows=rows[:1000].contiguous())


@pytest.mark.parametrize("M,K,N", [(4096, 512, 512), (5000, 512, 512), (4224, 192, 384), (100003, 512, 512), (8192, 1024, 256)])
def test_cta_pair_gemm_bit_identical_to_single_cta(fg, M, K, N):
    """cta_group::2 kernel (two SMs share a 256-row MMA, each staging half of the B tile) against the single-CTA kernel
    (same MMA order per output element -> bit-identical) and fp64."""
    g = torch.Generator().manual_seed(M)
    A = torch.randn(M, K, generator=g)
    W = torch.randn(N, K, generator=g) / K ** 0.5
    b = torch.randn(N, generator=g) * 0.1
    A_pl, W_pl, bd = fg.ops.split_bf16(A.to(DEV)), fg.ops.split_bf16(W.to(DEV)), b.to(DEV)
    for split in (False, True):
        old = fg._lib.set_tuning("gemm_pair", 0)
        try:
            y1 = fg.ops.gemm_bias_act(A_pl, W_pl, bd, fg.ops.ACT_ELU, precision=fg.ops.GEMM_BF16X3, split_out=split)
            fg._lib.set_tuning("gemm_pair", 1)
            y2 = fg.ops.gemm_bias_act(A_pl, W_pl, bd, fg.ops.ACT_ELU, precision=fg.ops.GEMM_BF16X3, split_out=split)
        finally:
            fg._lib.set_tuning("gemm_pair", old)
        if split:
            assert torch.equal(y1[0], y2[0]) and torch.equal(y1[1], y2[1])
            got = y2[0].float() + y2[1].float()
        else:
            assert torch.equal(y1, y2)
            got = y2
        want = torch.nn.functional.elu(A[:1500].double() @ W.double().T + b.double())
        assert (got[:1500].cpu().double() - want).abs().max() <= 1e-4 * want.abs().max()
