"""Stage-2 training step on the GPU: every backward kernel against torch autograd of the same op in fp32,
the whole denoiser block (forward + backward) and a short AdamW run against the CPU oracle (oracle/train.py), and the
drop-in command line (main_denoiser.py) incl. checkpoint layout and resume.
Tolerances: bf16 tensor-core GEMMs / attention with fp32 accumulation against fp32 references -> cosine >= 0.99 on
gradients (>= 0.999 on forward values), fp32 elementwise kernels to ~1e-5 relative."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _cos(a, b):
    return F.cosine_similarity(a.flatten().double(), b.flatten().double(), dim=0).item()


def _dev_ok():
    from dvt import _lib
    torch.cuda.synchronize()
    assert _lib.device_error() == 0


@pytest.mark.parametrize("B,N,heads", [(2, 257, 2), (1, 1369, 3), (2, 128, 1), (1, 300, 2), (3, 100, 1)])
def test_attention_backward_matches_autograd(B, N, heads):
    from dvt import train_ops
    g = torch.Generator(device="cuda").manual_seed(N + heads)
    C = heads * 64
    qkv = (torch.randn(B, N, 3 * C, device="cuda", generator=g) * 1.2).bfloat16()
    dout = (torch.randn(B, N, C, device="cuda", generator=g)).bfloat16()
    out, lse = train_ops.attention_fwd_lse(qkv, heads)
    dqkv = train_ops.attention_bwd(qkv, out, dout, lse, heads)
    _dev_ok()
    x = qkv.float().requires_grad_(True)
    q, k, v = x.reshape(B, N, 3, heads, 64).permute(2, 0, 3, 1, 4).unbind(0)
    s = (q @ k.transpose(-1, -2)) * 0.125
    ref_lse = torch.logsumexp(s, dim=-1) * 1.4426950408889634           # log2 domain
    ref = (s.softmax(-1) @ v).transpose(1, 2).reshape(B, N, C)
    ref.backward(dout.float())
    assert (out.float() - ref).abs().max().item() < 3e-2
    assert (lse - ref_lse).abs().max().item() < 2e-2
    gr = x.grad
    for name, sl in (("dq", slice(0, C)), ("dk", slice(C, 2 * C)), ("dv", slice(2 * C, 3 * C))):
        a, b = dqkv[..., sl].float(), gr[..., sl]
        assert _cos(a, b) > 0.998, f"{name}: cosine {_cos(a, b)}"
        assert (a - b).abs().max().item() < 0.03 * b.abs().max().item() + 1e-3, name


@pytest.mark.parametrize("rows,C", [(1370, 768), (333, 128), (2000, 384)])
def test_layernorm_backward(rows, C):
    from dvt import train_ops
    g = torch.Generator(device="cuda").manual_seed(rows)
    x = (torch.randn(rows, C, device="cuda", generator=g) * 2 + 0.5).requires_grad_(True)
    w = (1 + 0.2 * torch.randn(C, device="cuda", generator=g)).requires_grad_(True)
    b = torch.zeros(C, device="cuda", requires_grad=True)
    dy = torch.randn(rows, C, device="cuda", generator=g)
    F.layer_norm(x, (C,), w, b, 1e-6).backward(dy)
    acc0 = torch.randn(rows, C, device="cuda", generator=g)
    acc = acc0.clone()
    dg, db = train_ops.layernorm_bwd_(acc, x.detach(), w.detach(), dy)
    _dev_ok()
    assert (acc - acc0 - x.grad).abs().max().item() < 2e-4 * max(1.0, x.grad.abs().max().item())
    assert (dg - w.grad).abs().max().item() < 2e-4 * w.grad.abs().max().item() + 1e-4
    assert (db - b.grad).abs().max().item() < 2e-4 * b.grad.abs().max().item() + 1e-4


def test_bias_gelu_and_gemm_gradients():
    from dvt import train_ops
    g = torch.Generator(device="cuda").manual_seed(0)
    rows, n_out, n_in = 1000, 384, 256
    dy = torch.randn(rows, n_out, device="cuda", generator=g).bfloat16()
    x = torch.randn(rows, n_in, device="cuda", generator=g).bfloat16()
    w = (torch.randn(n_out, n_in, device="cuda", generator=g) / 16).bfloat16()
    pre = torch.randn(rows, n_in, device="cuda", generator=g).bfloat16()
    # column sums (bias gradients), bf16 and f32, odd row counts
    assert (train_ops.colsum(dy) - dy.float().sum(0)).abs().max().item() < 1e-2
    f = dy.float()[:777, :130].contiguous()
    assert (train_ops.colsum(f) - f.sum(0)).abs().max().item() < 1e-3
    # GELU forward
    assert (train_ops.gelu(pre).float() - F.gelu(pre.float())).abs().max().item() < 2e-2
    # data gradient (+ fused GELU derivative) and weight gradient
    ref_dx = dy.float() @ w.float()
    got = train_ops.dgrad(dy, w, torch.float32)
    assert (got - ref_dx).abs().max().item() < 2e-2 * ref_dx.abs().max().item()
    pf = pre.float().requires_grad_(True)
    F.gelu(pf).backward(ref_dx)
    got_g = train_ops.dgrad(dy, w, torch.bfloat16, gelu_preact=pre)
    assert _cos(got_g.float(), pf.grad) > 0.9995
    ref_dw = dy.float().t() @ x.float()
    got_w = train_ops.wgrad(dy, x)
    assert got_w.shape == (n_out, n_in) and (got_w - ref_dw).abs().max().item() < 2e-3 * ref_dw.abs().max().item() + 1e-2
    _dev_ok()


def test_loss_and_adamw_match_torch():
    from dvt import train_ops
    g = torch.Generator(device="cuda").manual_seed(1)
    pred = torch.randn(4, 5, 6, 768, device="cuda", generator=g).requires_grad_(True)
    tgt = torch.randn(4, 5, 6, 768, device="cuda", generator=g)
    loss, l2, cs = train_ops.denoise_loss(pred, tgt)
    (loss * 3.0).backward()
    got_grad = pred.grad.clone()
    pred.grad = None
    r_l2 = F.mse_loss(pred, tgt)
    r_cs = 1 - F.cosine_similarity(pred, tgt, dim=-1).mean()
    ((r_l2 + r_cs) * 3.0).backward()
    assert abs(l2.item() - r_l2.item()) < 1e-5 and abs(cs.item() - r_cs.item()) < 1e-5 and abs(loss.item() - (r_l2 + r_cs).item()) < 2e-5
    assert (got_grad - pred.grad).abs().max().item() < 1e-6 + 1e-4 * pred.grad.abs().max().item()
    # AdamW: five steps against torch.optim.AdamW with a changing learning rate
    p_ref = torch.randn(1000, device="cuda", generator=g).requires_grad_(True)
    opt = torch.optim.AdamW([p_ref], betas=(0.9, 0.999), weight_decay=1e-2)
    p = p_ref.detach().clone()
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    for step in range(1, 6):
        gr = torch.randn(1000, device="cuda", generator=g)
        lr = 1e-3 * step
        opt.param_groups[0]["lr"] = lr
        p_ref.grad = gr.clone()
        opt.step()
        train_ops.adamw_(p, gr, m, v, lr=lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, step=step)
    assert (p - p_ref.detach()).abs().max().item() < 2e-6
    _dev_ok()


def _denoiser(C, hw, nb, sd):
    import dvt.models as DVT
    m = DVT.Denoiser(hw[0], hw[1], C, vit=None, enable_pe=True, num_blocks=nb)
    m.load_state_dict(sd, strict=True)
    return m.cuda().train()


@pytest.mark.parametrize("nb,hw_in", [(1, (5, 6)), (2, (5, 6)), (1, (7, 9))], ids=["one-block", "two-blocks", "resampled-pe"])
def test_block_gradients_match_oracle(nb, hw_in):
    """Forward value, loss terms and the gradient of EVERY parameter (and of the input) of `Denoiser` against autograd
    through the CPU oracle."""
    from dvt import train_ops
    from oracle import denoiser as OD
    from oracle import train as OT
    C, hw, B = 128, (5, 6), 3
    sd = OD.random_state_dict(C, hw, nb, seed=10 + nb)
    g = torch.Generator().manual_seed(3)
    x = torch.randn(B, hw_in[0], hw_in[1], C, generator=g)
    tgt = torch.randn(B, hw_in[0], hw_in[1], C, generator=g)
    (r_loss, r_l2, r_cos), r_grads, r_pred = OT.gradients(sd, x, tgt, hw, nb)
    m = _denoiser(C, hw, nb, sd)
    xin = x.cuda().requires_grad_(True)
    pred = m(xin)
    loss, l2, cs = train_ops.denoise_loss(pred, tgt.cuda())
    loss.backward()
    _dev_ok()
    assert F.cosine_similarity(pred.detach().cpu().reshape(-1, C), r_pred.reshape(-1, C), dim=-1).min().item() > 0.999
    assert abs(loss.item() - r_loss) < 2e-2 * abs(r_loss) and abs(cs.item() - r_cos) < 2e-2 * abs(r_cos) + 1e-3
    for name, p in m.named_parameters():
        c = _cos(p.grad.cpu(), r_grads[name])
        assert c > 0.99, f"{name}: gradient cosine {c}"
        ratio = p.grad.norm().item() / (r_grads[name].norm().item() + 1e-12)
        assert 0.95 < ratio < 1.05, f"{name}: gradient norm ratio {ratio}"
    assert _cos(xin.grad.cpu(), r_grads["__input__"]) > 0.99


def test_training_run_matches_oracle():
    """30 AdamW steps (warm-up + cosine schedule) on synthetic pairs: loss trajectory and final parameters against the
    oracle loop (torch autograd + torch.optim.AdamW in fp32 on the CPU)."""
    from dvt import train_ops
    from dvt.optim import FusedAdamW
    from dvt.utils import misc
    from oracle import denoiser as OD
    from oracle import train as OT
    C, hw, B, T = 128, (5, 6), 4, 30
    sd = OD.random_state_dict(C, hw, 1, seed=21)
    g = torch.Generator().manual_seed(5)
    clean = torch.randn(8, hw[0], hw[1], C, generator=g)
    noise = 0.5 * torch.randn(1, hw[0], hw[1], C, generator=g)          # a position-dependent artifact, shared by all images
    batches = []
    for s in range(T):
        ids = torch.randint(0, 8, (B,), generator=g)
        batches.append((clean[ids] + noise, clean[ids]))
    sched = dict(base_value=2e-3, final_value=1e-5, total_iters=T, warmup_iters=int(T * 0.15), start_warmup_value=0)
    lrs = OT.cosine_schedule(**sched)
    ref_sd, ref_logs = OT.train(sd, batches, hw, lr_values=lrs, weight_decay=1e-5)
    m = _denoiser(C, hw, 1, sd)
    opt = FusedAdamW(m.parameters(), betas=(0.9, 0.999), weight_decay=1e-5)
    logs = []
    for s, (x, t) in enumerate(batches):
        lr = misc.cosine_schedule(s, **sched)
        assert abs(lr - lrs[s]) < 1e-12
        misc.apply_optim_scheduler(opt, lr)
        loss, l2, cs = train_ops.denoise_loss(m(x.cuda()), t.cuda())
        opt.zero_grad()
        loss.backward()
        opt.step()
        logs.append((loss.item(), l2.item(), cs.item()))
    _dev_ok()
    logs = np.array(logs)
    assert np.allclose(logs[:, 0], ref_logs[:, 0], rtol=0.03, atol=2e-3), np.abs(logs[:, 0] - ref_logs[:, 0]).max()
    assert logs[-1, 0] < 0.8 * logs[0, 0], "the loss must fall"
    got_sd = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    for k, v in ref_sd.items():
        upd_ref, upd_got = v - sd[k], got_sd[k] - sd[k]
        if k.endswith("attn.qkv.bias"):
            # softmax is invariant to a shift of all keys: the exact gradient of the k-bias is zero, what Adam normalises
            # there is rounding noise -- compare the q and v thirds only
            keep = torch.cat([torch.arange(0, C), torch.arange(2 * C, 3 * C)])
            upd_ref, upd_got = upd_ref[keep], upd_got[keep]
        assert _cos(upd_got, upd_ref) > 0.97, f"{k}: update cosine {_cos(upd_got, upd_ref)}"
    # optimiser state in torch.optim layout
    st = opt.state_dict()
    assert set(st) == {"state", "param_groups"} and len(st["state"]) == len(list(m.parameters()))
    assert float(st["state"][0]["step"]) == T and st["state"][0]["exp_avg"].shape == next(m.parameters()).shape


def test_stage2_cli_trains_and_checkpoints(tmp_path, capsys):
    """main_denoiser.py on a small feature store written by the stage-1 writer: loss falls, checkpoint has the reference's
    layout ({"denoiser", "optimizer", "step"}, latest.pth), `--resume` continues from it."""
    sys.path.insert(0, ROOT)
    import main_denoiser as M
    from argparse import Namespace
    from dvt.store import FeatureStoreWriter
    from dvt.utils import misc
    model = "vit_small_patch14_dinov2.lvd142m"
    h = w = 5                                                     # input 70 x 70, patch 14
    data_root = str(tmp_path / "data") + "/"
    sargs = Namespace(data_root=data_root, save_root=str(tmp_path / "feats"), model=model)
    g = torch.Generator().manual_seed(0)
    noise = 0.5 * torch.randn(h, w, 384, generator=g)
    wr = FeatureStoreWriter()
    rels = [f"img/{i}.jpg" for i in range(6)]
    for rel in rels:
        clean = torch.randn(h, w, 384, generator=g)
        wr.submit(*misc.feature_paths(sargs, os.path.join(data_root, rel)), clean + noise, clean[None])
    wr.close()
    lst = tmp_path / "list.txt"
    lst.write_text("".join(f"{r} 0\n" for r in rels))
    argv = ["--model", model, "--input_size", "70", "--stride_size", "14", "--data_root", data_root, "--feat_root",
            f"{sargs.save_root}/denoised_features/{model}/", "--data_list_path", str(lst), "--batch_size", "4",
            "--num_iterations", "40", "--blr", "0.02", "--output_root", str(tmp_path / "work"), "--run_name", "t",
            "--save_freq", "20", "--num_workers", "0", "--log_freq", "10"]
    M.main(M.get_args(argv))
    out = capsys.readouterr().out
    vals = [float(ln.split("loss: ")[1].split()[0]) for ln in out.splitlines() if ln.startswith("Train [")]
    assert len(vals) >= 4 and vals[-1] < vals[0]
    ck_dir = tmp_path / "work" / "denosing-vit" / "t" / "checkpoints"
    assert sorted(os.listdir(ck_dir)) == ["ckpt_000000.pth", "ckpt_000020.pth", "ckpt_000039.pth", "latest.pth"]
    assert os.path.islink(ck_dir / "latest.pth")
    ck = torch.load(str(ck_dir / "latest.pth"), map_location="cpu")
    assert set(ck) == {"denoiser", "optimizer", "step"} and ck["step"] == 39
    assert "pos_embed" in ck["denoiser"] and "denoiser.attn.qkv.weight" in ck["denoiser"] and not any("vit." in k for k in ck["denoiser"])
    assert set(ck["optimizer"]) == {"state", "param_groups"} and ck["optimizer"]["param_groups"][0]["betas"] == (0.9, 0.999)
    # resume: continues at step 40 of a longer schedule
    M.main(M.get_args(argv[:argv.index("--num_iterations") + 1] + ["44"] + argv[argv.index("--num_iterations") + 2:]
                      + ["--resume", str(ck_dir / "latest.pth")]))
    out2 = capsys.readouterr().out
    assert "Resumed from" in out2 and "at step 40" in out2 and os.path.isfile(ck_dir / "ckpt_000043.pth")
