"""torch.library operators (safe_dreamer_b200/ops.py): registered with schemas, fake implementations that propagate
shapes without touching CUDA (what torch.compile traces), and an autograd formula on the differentiable ones."""
from types import SimpleNamespace as NS

import torch
from torch._subclasses.fake_tensor import FakeTensorMode

from oracle import rssm_oracle as O


def _module():
    from safe_dreamer_b200.rssm import RSSM
    c = O.Cfg(D=64, U=16, S=4, K=8, G=2, E=24, A=3)
    cfg = NS(stoch=c.S, deter=c.D, hidden=c.U, discrete=c.K, act="SiLU", unimix_ratio=c.unimix, initial="learned",
             device="cpu", obs_layers=1, img_layers=2, dyn_layers=1, blocks=c.G)
    return c, RSSM(cfg, c.E, c.A)


def test_operators_are_registered_with_fake_and_autograd():
    from safe_dreamer_b200 import ops
    names = ["observe", "observe_bwd", "prior", "prior_bwd", "kl_loss", "kl_loss_bwd", "imagine", "heads_lambda", "lambda_return"]
    for n in names:
        op = getattr(torch.ops.safedreamer, n)
        assert op.default._schema.name == f"safedreamer::{n}"
    c, rssm = _module()
    key = ops.module_key(rssm)
    B, T, H, N = 3, 5, 4, 7
    with FakeTensorMode():
        params = [torch.empty(p.shape) for p in rssm.parameters()]
        st, dt, lg = torch.ops.safedreamer.observe(torch.empty(B, T, c.E), torch.empty(B, T, c.A), torch.empty(B, c.S, c.K),
                                                   torch.empty(B, c.D), torch.empty(B, T, dtype=torch.uint8),
                                                   torch.empty(B, T, c.S, c.K), params, key, True)
        assert st.shape == (B, T, c.S, c.K) and dt.shape == (B, T, c.D) and lg.shape == (B, T, c.S, c.K)
        de, dis, did, wg = torch.ops.safedreamer.observe_bwd(st, dt, lg, params, key, True, True, True)
        assert de.shape == (B, T, c.E) and dis.shape == (B, c.S, c.K) and did.shape == (B, c.D)
        assert [tuple(g.shape) for g in wg] == [tuple(p.shape) for p in params]
        ps, pl = torch.ops.safedreamer.prior(dt, torch.empty(B, T, c.S, c.K), params, key, True)
        assert ps.shape == pl.shape == (B, T, c.S, c.K)
        dyn, rep = torch.ops.safedreamer.kl_loss(lg, pl, 1.0, key)
        assert dyn.shape == rep.shape == (B, T)
        feats, acts = torch.ops.safedreamer.imagine(torch.empty(N, c.S, c.K), torch.empty(N, c.D), torch.empty(N, H, c.S, c.K),
                                                    torch.empty(N, H, c.A), H, key)
        assert feats.shape == (N, H, c.S * c.K + c.D) and acts.shape == (N, H, c.A)
        outs = torch.ops.safedreamer.heads_lambda(feats, 0.99, 0.95, key)
        assert [tuple(o.shape) for o in outs] == [(N, H, 1)] * 5 + [(N, H - 1, 1)]
        ret = torch.ops.safedreamer.lambda_return(*[torch.empty(N, H, 1)] * 5, 0.99, 0.95, key)
        assert ret.shape == (N, H - 1, 1)


def test_forward_only_entries_refuse_gradients():
    c, rssm = _module()
    stoch = torch.zeros(2, c.S, c.K, requires_grad=True)
    try:
        rssm.img_step(stoch, torch.zeros(2, c.D), torch.zeros(2, c.A))
    except RuntimeError as e:
        assert "forward-only" in str(e)
    else:
        raise AssertionError("img_step accepted an input that requires grad")


def test_cnn_encoder_operator_fake_shapes():
    from safe_dreamer_b200 import ops
    from safe_dreamer_b200.encoder import ConvEncoder
    cfg = NS(act="SiLU", norm=True, kernel_size=5, minres=4, depth=16, mults=[2, 3, 4, 4])
    enc = ConvEncoder(cfg, (64, 64, 3))
    assert enc.out_dim == 1024
    assert sorted(enc.state_dict()) == sorted(f"layers.{4 * i + j}.{n}" for i in range(4) for j, n in ((0, "weight"), (0, "bias"), (2, "weight")))
    key = ops.module_key(enc)
    assert torch.ops.safedreamer.cnn_encoder.default._schema.name == "safedreamer::cnn_encoder"
    with FakeTensorMode():
        params = [torch.empty(p.shape) for p in enc._tensors()]
        emb = torch.ops.safedreamer.cnn_encoder(torch.empty(2, 5, 64, 64, 3), params, key, True)
        assert emb.shape == (2, 5, 1024)
        d_obs, wg = torch.ops.safedreamer.cnn_encoder_bwd(emb, torch.empty(2, 5, 64, 64, 3), params, key, True, True)
        assert d_obs.shape == (10, 64, 64, 3)
        assert [tuple(g.shape) for g in wg] == [tuple(p.shape) for p in params]
    import copy
    twin = copy.deepcopy(enc)          # dreamer.py:263 clone_and_freeze deep-copies the encoder
    assert twin._eng is None and twin._ops_key is None and sorted(twin.state_dict()) == sorted(enc.state_dict())
