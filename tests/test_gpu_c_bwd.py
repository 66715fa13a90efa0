"""GPU parity of the reverse-time backward scans through the C ABI (fp32 path) against the oracle's manual
backward AND the reference's torch.autograd gradients stored in the goldens.
Tolerance: rtol 3e-3 / atol 3e-5 x max(1, |grad|_2) (fp32 accumulation order differs)."""
import numpy as np
import pytest
import torch

from oracle import rssm_oracle as O
from tests.helpers import cu, golden_initial, golden_params, load_golden, make_engine

pytestmark = pytest.mark.gpu
CASES = ["tiny_cont", "tiny_onehot", "base_cont"]


def _np(t):
    return t.detach().cpu().numpy()


@pytest.mark.parametrize("tag", CASES)
def test_observe_bwd(tag):
    c, z = load_golden(tag)
    P = golden_params(c, z)
    B, T = int(z["B"]), int(z["T"])
    eng = make_engine(c, P, max_rows=16, max_steps=8, max_tape_rows=8)
    embed, action, reset, u = O.synth_observe_inputs(c, B, T, seed=2)
    reset = reset.copy(); reset[0, 0] = False
    s0, d0 = golden_initial(c, B)
    g = np.random.Generator(np.random.Philox(13))
    c_st = g.standard_normal((B, T, c.S, c.K), dtype=np.float32)
    c_dt = g.standard_normal((B, T, c.D), dtype=np.float32) * np.float32(0.1)
    c_lg = g.standard_normal((B, T, c.S, c.K), dtype=np.float32) * np.float32(0.1)
    for flags in (0, 4, 4):
        st, dt, lg = eng.observe(cu(embed), cu(action), cu(s0), cu(d0), cu(reset), cu(u), flags=2 | flags)
        np.testing.assert_array_equal(_np(st).argmax(-1).astype(np.int8), z["obs_stoch_idx"])
        wg = {n: torch.zeros(P["rssm"][n].shape, device="cuda") for n in eng.weight_names(0)}
        d_embed, d_is, d_id = eng.observe_bwd(B, T, cu(c_st), cu(c_dt), cu(c_lg), True, True, wg, flags=flags)
        torch.cuda.synchronize()
        np.testing.assert_allclose(_np(d_embed), z["bwd_d_embed"], rtol=3e-3, atol=3e-5)
        np.testing.assert_allclose(_np(d_is), z["bwd_d_init_stoch"], rtol=3e-3, atol=3e-5)
        np.testing.assert_allclose(_np(d_id), z["bwd_d_init_deter"], rtol=3e-3, atol=3e-5)
        for name in O.rssm_param_shapes(c):
            got = _np(wg[name])
            gn = float(z["bwd_gn/" + name])
            if "bwd_g/" + name in z.files:
                np.testing.assert_allclose(got, z["bwd_g/" + name], rtol=3e-3, atol=3e-5 * max(1.0, gn), err_msg=name)
            else:
                sl = got.reshape(-1)[:: max(1, got.size // 2048)][:2048]
                np.testing.assert_allclose(sl, z["bwd_gs/" + name], rtol=3e-3, atol=3e-5 * max(1.0, gn), err_msg=name)
            assert abs(np.sqrt((got.astype(np.float64) ** 2).sum()) - gn) <= 3e-3 * max(gn, 1e-3), name
    # dgrad-only call (frozen weights) gives the same input grads
    d_embed2, d_is2, d_id2 = eng.observe_bwd(B, T, cu(c_st), cu(c_dt), cu(c_lg), True, True, None)
    np.testing.assert_array_equal(_np(d_embed2), _np(d_embed))
    # no tape -> loud error
    eng.observe(cu(embed), cu(action), cu(s0), cu(d0), cu(reset), cu(u), flags=0)
    with pytest.raises(RuntimeError, match="SAVE_TAPE"):
        eng.observe_bwd(B, T + 1, cu(c_st), cu(c_dt), cu(c_lg))


def test_module_autograd_matches_reference_grads():
    """The drop-in RSSM module: loss.backward() through observe reproduces the reference's autograd."""
    from types import SimpleNamespace as NS
    from safe_dreamer_b200.rssm import RSSM
    c, z = load_golden("tiny_cont")
    P = golden_params(c, z)
    B, T = int(z["B"]), int(z["T"])
    cfg = NS(stoch=c.S, deter=c.D, hidden=c.U, discrete=c.K, act="SiLU", unimix_ratio=c.unimix, initial="learned",
             device="cuda", obs_layers=c.obs_layers, img_layers=c.img_layers, dyn_layers=1, blocks=c.G)
    rssm = RSSM(cfg, c.E, c.A).cuda()
    rssm.load_state_dict({k: cu(v) for k, v in P["rssm"].items()})
    import copy
    frozen = copy.deepcopy(rssm)  # dreamer.py:276 clone_and_freeze must keep working
    assert frozen._rt.engine is None
    embed, action, reset, u = O.synth_observe_inputs(c, B, T, seed=2)
    reset = reset.copy(); reset[0, 0] = False
    s0, d0 = golden_initial(c, B)
    rssm.noise_source = lambda shape, dev: cu(u).reshape(shape)
    e = cu(embed).requires_grad_(True)
    st, dt, lg = rssm.observe(e, cu(action), (cu(s0), cu(d0)), cu(reset)[..., None])
    g = np.random.Generator(np.random.Philox(13))
    c_st = g.standard_normal(st.shape, dtype=np.float32)
    c_dt = g.standard_normal(dt.shape, dtype=np.float32) * np.float32(0.1)
    c_lg = g.standard_normal(lg.shape, dtype=np.float32) * np.float32(0.1)
    ((st * cu(c_st)).sum() + (dt * cu(c_dt)).sum() + (lg * cu(c_lg)).sum()).backward()
    np.testing.assert_allclose(_np(e.grad), z["bwd_d_embed"], rtol=3e-3, atol=3e-5)
    for name, p in rssm.named_parameters():
        if name.startswith("_img_net"):
            continue
        np.testing.assert_allclose(_np(p.grad), z["bwd_g/" + name], rtol=3e-3,
                                   atol=3e-5 * max(1.0, float(z["bwd_gn/" + name])), err_msg=name)


def test_module_static_fast_path_same_grads():
    """The low-overhead module settings bench.py's end-to-end loop uses (CUDA graphs, engine-owned outputs, caller-owned
    pointer-stable inputs, cached parameter dicts, p.grad aliasing the gradient bucket) give bit-identical outputs and
    gradients to the default settings, step after step (second step with different inputs in the same buffers)."""
    from types import SimpleNamespace as NS
    from safe_dreamer_b200.rssm import RSSM
    c, z = load_golden("tiny_cont")
    P = golden_params(c, z)
    B, T = int(z["B"]), int(z["T"])
    cfg = NS(stoch=c.S, deter=c.D, hidden=c.U, discrete=c.K, act="SiLU", unimix_ratio=c.unimix, initial="learned",
             device="cuda", obs_layers=c.obs_layers, img_layers=c.img_layers, dyn_layers=1, blocks=c.G)
    mods = []
    for fast in (False, True):
        m = RSSM(cfg, c.E, c.A).cuda()
        m.load_state_dict({k: cu(v) for k, v in P["rssm"].items()})
        if fast:
            m.use_graph, m.auto_refresh, m.static_outputs = True, False, True
            m.stage_inputs, m.cache_params, m.static_grads = False, True, True
        mods.append(m)
    s0, d0 = golden_initial(c, B)
    bufs = None
    for step in range(3):
        embed, action, reset, u = O.synth_observe_inputs(c, B, T, seed=20 + step)
        g = np.random.Generator(np.random.Philox(90 + step))
        cot = [g.standard_normal(sh, dtype=np.float32) * np.float32(0.1) for sh in ((B, T, c.S, c.K), (B, T, c.D), (B, T, c.S, c.K))]
        outs = []
        for fast, m in zip((False, True), mods):
            m.noise_source = lambda shape, dev: cu(u).reshape(shape)
            for p_ in m.parameters():
                p_.grad = None
            if fast:   # persistent device inputs, refilled in place
                if bufs is None:
                    bufs = [cu(embed), cu(action), cu(reset)[..., None].contiguous(), cu(s0), cu(d0)]
                else:
                    with torch.no_grad():
                        bufs[0].copy_(cu(embed)); bufs[1].copy_(cu(action)); bufs[2].copy_(cu(reset)[..., None])
                bufs[0].grad = None
                e, a, r, si, di = bufs
                m.refresh_weights(force=True)
            else:
                e, a, r, si, di = cu(embed), cu(action), cu(reset)[..., None], cu(s0), cu(d0)
            e.requires_grad_(True)
            st, dt, lg = m.observe(e, a, (si, di), r)
            torch.autograd.backward((st, dt, lg), tuple(cu(x) for x in cot))
            outs.append(([_np(st).copy(), _np(dt).copy(), _np(lg).copy(), _np(e.grad).copy()],
                         {n: _np(p_.grad).copy() for n, p_ in m.named_parameters() if p_.grad is not None}))
        for x, y in zip(outs[0][0], outs[1][0]):
            np.testing.assert_array_equal(x, y)
        assert outs[0][1].keys() == outs[1][1].keys()
        for n in outs[0][1]:
            np.testing.assert_array_equal(outs[0][1][n], outs[1][1][n], err_msg=n)


@pytest.mark.parametrize("tag", CASES)
def test_imagine_bwd(tag):
    """dgrad-only backward through the imagination rollout (frozen weights; the attack shape)."""
    c, z = load_golden(tag)
    P = golden_params(c, z)
    N, H = int(z["N"]), int(z["H"])
    eng = make_engine(c, P, max_rows=16, max_steps=8, max_tape_rows=8)
    st0, dt0, u, noise = O.synth_imagine_inputs(c, N, H, seed=3)
    g2 = np.random.Generator(np.random.Philox(17))
    c_f = g2.standard_normal((N, H, c.F), dtype=np.float32) * np.float32(0.1)
    c_a = g2.standard_normal((N, H, c.A), dtype=np.float32)
    for flags in (0, 4, 4):
        feats, acts = eng.imagine(cu(st0), cu(dt0), cu(u), cu(noise), H, flags=2 | flags)
        ds, dd = eng.imagine_bwd(N, H, cu(c_f), cu(c_a), flags=flags)
        torch.cuda.synchronize()
        np.testing.assert_allclose(_np(feats)[..., c.SK:], z["imag_deter"], atol=5e-5, rtol=0)
        np.testing.assert_allclose(_np(ds), z["imag_bwd_d_stoch"], rtol=3e-3, atol=3e-5)
        np.testing.assert_allclose(_np(dd), z["imag_bwd_d_deter"], rtol=3e-3, atol=3e-5)
    with pytest.raises(RuntimeError, match="SAVE_TAPE"):
        eng.imagine_bwd(N, H + 1, cu(c_f), cu(c_a))


@pytest.mark.parametrize("tag", ["tiny_cont", "base_cont"])
def test_prior_bwd(tag):
    """Batched prior (dreamer.py:485) forward + backward: d_deter and the _img_net weight grads vs the oracle."""
    c, z = load_golden(tag)
    P = golden_params(c, z)
    R = 24
    eng = make_engine(c, P, max_rows=8, max_steps=4, max_tape_rows=1)
    rng = np.random.Generator(np.random.Philox(31))
    deter = np.tanh(rng.standard_normal((R, c.D), dtype=np.float32)).astype(np.float32)
    u = O.clamp_u(rng.random((R, c.S, c.K), dtype=np.float32))
    d_lg = rng.standard_normal((R, c.S, c.K), dtype=np.float32) * np.float32(0.1)
    d_st = rng.standard_normal((R, c.S, c.K), dtype=np.float32)
    tp = {}
    lg_o = O.img_logit(c, P["rssm"], deter, tp)
    st_o, idx_o, y_o, _ = O.sample_onehot(lg_o, u, c.unimix)
    for with_stoch in (False, True):
        G = {}
        g_l = d_lg + (O.sample_bwd(d_st, lg_o, y_o, c.unimix) if with_stoch else 0)
        dd_o = O.mlp_logits_bwd(g_l.reshape(R, -1), P["rssm"], G, tp["img_acts"], tp["img_last_in"], "_img_net.img_net_",
                                "_img_net.img_net_n_", "_img_net.img_net_logit")
        st, lg = eng.prior(cu(deter), cu(u), flags=2)
        np.testing.assert_allclose(_np(lg), lg_o, atol=2e-4, rtol=0)
        np.testing.assert_array_equal(_np(st).argmax(-1), idx_o)
        wg = {n: torch.zeros(P["rssm"][n].shape, device="cuda") for n in eng.weight_names(0)}
        dd = eng.prior_bwd(R, cu(d_st) if with_stoch else None, cu(d_lg), True, wg)
        torch.cuda.synchronize()
        np.testing.assert_allclose(_np(dd), dd_o, rtol=3e-3, atol=3e-5)
        for name, g in G.items():
            gn = float(np.sqrt((g.astype(np.float64) ** 2).sum()))
            np.testing.assert_allclose(_np(wg[name]), g, rtol=3e-3, atol=3e-5 * max(1.0, gn), err_msg=name)
        for name in wg:
            if not name.startswith("_img_net"):
                assert float(wg[name].abs().sum()) == 0.0, name
    with pytest.raises(RuntimeError, match="SAVE_TAPE"):
        eng.prior_bwd(R + 1, None, cu(d_lg))


def test_world_model_update_autograd():
    """observe -> batched prior -> kl_loss -> backward through the drop-in module (dreamer.py:483-486,667):
    gradients reach embed, the initial state and every RSSM parameter (incl. _img_net through the prior)."""
    from types import SimpleNamespace as NS
    from safe_dreamer_b200.rssm import RSSM
    c, z = load_golden("tiny_cont")
    P = golden_params(c, z)
    B, T = 4, 5
    cfg = NS(stoch=c.S, deter=c.D, hidden=c.U, discrete=c.K, act="SiLU", unimix_ratio=c.unimix, initial="learned",
             device="cuda", obs_layers=c.obs_layers, img_layers=c.img_layers, dyn_layers=1, blocks=c.G)
    rssm = RSSM(cfg, c.E, c.A).cuda()
    rssm.load_state_dict({k: cu(v) for k, v in P["rssm"].items()})
    rssm.max_rows, rssm.max_steps = 32, 8
    embed, action, reset, u = O.synth_observe_inputs(c, B, T, seed=2)
    up = O.clamp_u(np.random.Generator(np.random.Philox(41)).random((B, T, c.S, c.K), dtype=np.float32))
    queue = [u, up]
    rssm.noise_source = lambda shape, dev: cu(queue.pop(0)).reshape(shape)
    e = cu(embed).requires_grad_(True)
    st, dt, lg = rssm.observe(e, cu(action), rssm.initial(B), cu(reset)[..., None])
    _, plog = rssm.prior(dt)
    dyn, rep = rssm.kl_loss(lg, plog, 1.0)
    (dyn.mean() + 0.1 * rep.mean() + (st * 0.01).sum()).backward()
    assert e.grad is not None and torch.isfinite(e.grad).all() and float(e.grad.abs().sum()) > 0
    for name, p in rssm.named_parameters():
        assert p.grad is not None and torch.isfinite(p.grad).all(), name
        assert float(p.grad.abs().sum()) > 0, name
    # the same gradient from the oracle: d(loss)/d(post_logit), d/d(prior_logit) by torch on the tiny KL term,
    # then the oracle's manual backward through prior and observe
    tapes = []
    zero = (np.zeros((B, c.S, c.K), np.float32), np.zeros((B, c.D), np.float32))
    st_o, dt_o, lg_o, idx_o = O.observe(c, P["rssm"], embed, action, zero, reset, u, tapes)
    np.testing.assert_array_equal(_np(st).argmax(-1), idx_o)
    tp = {}
    plog_o = O.img_logit(c, P["rssm"], dt_o.reshape(B * T, -1), tp).reshape(B, T, c.S, c.K)
    a = torch.from_numpy(lg_o).requires_grad_(True)
    b = torch.from_numpy(plog_o).requires_grad_(True)
    from safe_dreamer_b200.distributions import kl
    dyn_o = torch.clip(kl(a.detach(), b).sum(-1), min=1.0).mean()
    rep_o = torch.clip(kl(a, b.detach()).sum(-1), min=1.0).mean()
    (dyn_o + 0.1 * rep_o).backward()
    G = {}
    dd_prior = O.mlp_logits_bwd(b.grad.numpy().reshape(B * T, -1), P["rssm"], G, tp["img_acts"], tp["img_last_in"],
                                "_img_net.img_net_", "_img_net.img_net_n_", "_img_net.img_net_logit")
    G2, d_embed, d_is, d_id = O.observe_bwd(c, P["rssm"], tapes, np.full(st_o.shape, 0.01, np.float32),
                                            dd_prior.reshape(B, T, -1), a.grad.numpy())
    np.testing.assert_allclose(_np(e.grad), d_embed, rtol=5e-3, atol=1e-6)
    for name, p in rssm.named_parameters():
        ref = G.get(name, 0) + G2[name]
        gn = float(np.sqrt((ref.astype(np.float64) ** 2).sum()))
        np.testing.assert_allclose(_np(p.grad), ref, rtol=5e-3, atol=5e-5 * max(gn, 1e-3), err_msg=name)


def test_imagine_bwd_tcgen05_dgrad():
    """Large-row imagination backward with the dgrads on tcgen05 (bf16 copies of the gradients, fp32
    accumulate): fp32 forward (so the sampled path equals the oracle's), bf16 backward.
    Tolerance: relative L2 error <= 3 %, cosine >= 0.999 against the fp32 oracle gradients."""
    c, z = load_golden("base_cont")
    P = golden_params(c, z)
    N, H = 128, 3
    eng = make_engine(c, P, max_rows=N, max_steps=4, max_tape_rows=N)
    st0, dt0, u, noise = O.synth_imagine_inputs(c, N, H, seed=33)
    g2 = np.random.Generator(np.random.Philox(34))
    c_f = g2.standard_normal((N, H, c.F), dtype=np.float32) * np.float32(0.1)
    c_a = g2.standard_normal((N, H, c.A), dtype=np.float32)
    tapes = []
    feats_o, acts_o = O.imagine(c, P["rssm"], P["actor"], (st0, dt0), H, u, noise, tapes)
    ds_o, dd_o = O.imagine_bwd(c, P["rssm"], P["actor"], tapes, c_f, c_a)
    feats, acts = eng.imagine(cu(st0), cu(dt0), cu(u), cu(noise), H, flags=2)
    np.testing.assert_array_equal(_np(feats)[..., :c.SK], feats_o[..., :c.SK])
    for flags, tol in ((0, 3e-3), (1, 3e-2)):
        ds, dd = eng.imagine_bwd(N, H, cu(c_f), cu(c_a), flags=flags)
        torch.cuda.synchronize()
        for got, ref, name in ((_np(ds), ds_o, "d_stoch0"), (_np(dd), dd_o, "d_deter0")):
            rel = np.linalg.norm(got - ref) / np.linalg.norm(ref)
            cos = float((got * ref).sum() / (np.linalg.norm(got) * np.linalg.norm(ref)))
            print(f"imagine_bwd flags={flags} {name}: rel L2 err {rel:.2e}, cosine {cos:.6f}")
            assert rel <= tol and cos >= 0.999, (flags, name, rel, cos)


def _tiny_module(max_rows, max_steps):
    from types import SimpleNamespace as NS
    from safe_dreamer_b200.rssm import RSSM
    c, z = load_golden("tiny_cont")
    P = golden_params(c, z)
    cfg = NS(stoch=c.S, deter=c.D, hidden=c.U, discrete=c.K, act="SiLU", unimix_ratio=c.unimix, initial="learned",
             device="cuda", obs_layers=c.obs_layers, img_layers=c.img_layers, dyn_layers=1, blocks=c.G)
    rssm = RSSM(cfg, c.E, c.A).cuda()
    rssm.load_state_dict({k: cu(v) for k, v in P["rssm"].items()})
    rssm.max_rows, rssm.max_steps = max_rows, max_steps
    return c, P, rssm


def test_module_update_with_more_rows_than_max_rows():
    """B*T > rssm.max_rows: the batched prior / kl_loss must not rebuild the engine between observe's forward and its
    backward (the tape lives in the engine).  Gradients equal those of a module whose max_rows covers B*T."""
    B, T = 5, 8
    grads = []
    for max_rows in (16, 64):       # 40 rows: above / below the limit
        c, P, rssm = _tiny_module(max_rows, 8)
        embed, action, reset, u = O.synth_observe_inputs(c, B, T, seed=2)
        up = O.clamp_u(np.random.Generator(np.random.Philox(41)).random((B, T, c.S, c.K), dtype=np.float32))
        queue = [u, up]
        rssm.noise_source = lambda shape, dev: cu(queue.pop(0)).reshape(shape)
        e = cu(embed).requires_grad_(True)
        st, dt, lg = rssm.observe(e, cu(action), rssm.initial(B), cu(reset)[..., None])
        eng0 = rssm._rt.engine
        _, plog = rssm.prior(dt)
        dyn, rep = rssm.kl_loss(lg, plog, 1.0)
        assert rssm._rt.engine is eng0
        (dyn.mean() + 0.1 * rep.mean() + (st * 0.01).sum()).backward()
        grads.append([_np(e.grad)] + [_np(p.grad) for p in rssm.parameters()])
    for a, b in zip(*grads):
        np.testing.assert_allclose(a, b, rtol=1e-5, atol=1e-7)


def test_stale_tape_is_detected():
    """The library keeps one tape per handle: a second grad-enabled observe() before the first backward overwrites it, and the
    first backward must fail loudly instead of returning gradients of the wrong call."""
    B, T = 3, 4
    c, P, rssm = _tiny_module(16, 8)
    embed, action, reset, u = O.synth_observe_inputs(c, B, T, seed=2)
    e1 = cu(embed).requires_grad_(True)
    e2 = cu(embed * 0.5).requires_grad_(True)
    out1 = rssm.observe(e1, cu(action), rssm.initial(B), cu(reset)[..., None])
    out2 = rssm.observe(e2, cu(action), rssm.initial(B), cu(reset)[..., None])
    with pytest.raises(RuntimeError, match="tape"):
        out1[1].sum().backward()
    out2[1].sum().backward()          # the latest forward still owns the tape
    assert e2.grad is not None and torch.isfinite(e2.grad).all()
