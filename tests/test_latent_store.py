"""Replay latent write-back: Buffer.update / `initial` of utils/buffer.py:40,44-53.
CPU: the numpy oracle against the golden written by the reference's own Buffer.update (tests/golden/make_golden.py:
run_latent_store).  GPU: sd_latent_writeback / sd_latent_gather (safe_dreamer_b200.replay.LatentStore) against the oracle,
incl. repeated slots (last row wins), out-of-range indices (IndexError) and the full B*T = 1024 x (32x16, 2048) size."""
import os

import numpy as np
import pytest

from oracle import rssm_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))


def _inputs():
    # restated from tests/golden/make_golden.py:latent_store_inputs
    rng = np.random.Generator(np.random.Philox(4242))
    B, T, S, K, D, n_len, n_env = 3, 5, 4, 8, 12, 20, 4
    env = np.array([0, 2, 3], np.int64)[:, None].repeat(T, 1)
    t0 = np.array([1, 7, 13], np.int64)[:, None]
    time = t0 + np.arange(T, dtype=np.int64)[None]
    cls = rng.integers(0, K, size=(B, T, S))
    stoch = np.eye(K, dtype=np.float32)[cls]
    deter = rng.standard_normal((B, T, D), dtype=np.float32)
    store_stoch = rng.standard_normal((n_len, n_env, S, K), dtype=np.float32)
    store_deter = rng.standard_normal((n_len, n_env, D), dtype=np.float32)
    return env, time, stoch, deter, store_stoch, store_deter


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(HERE, "golden", "latent_store.npz"))


def test_oracle_writeback_matches_reference_update(golden):
    env, time, stoch, deter, store_stoch, store_deter = _inputs()
    idx = np.zeros(store_stoch.shape[:3], np.uint8)
    O.latent_writeback([env, time], stoch, deter, store_deter, idx, store_stoch)
    np.testing.assert_array_equal(store_stoch, golden["store_stoch"])
    np.testing.assert_array_equal(store_deter, golden["store_deter"])
    # the class-index form decodes to the same rows the reference stored
    st, dt = O.latent_initial([env, time], idx, store_deter, stoch.shape[-1])
    np.testing.assert_array_equal(st, golden["store_stoch"][time.reshape(-1), env.reshape(-1)])
    np.testing.assert_array_equal(dt, golden["store_deter"][time.reshape(-1), env.reshape(-1)])


def test_oracle_repeated_slot_last_row_wins():
    S, K, D = 2, 4, 3
    env = np.array([[1, 1, 0]], np.int64); time = np.array([[2, 2, 2]], np.int64)
    stoch = np.eye(K, dtype=np.float32)[np.array([[[0, 1], [2, 3], [1, 1]]])]
    deter = np.arange(9, dtype=np.float32).reshape(1, 3, D)
    sd = np.zeros((4, 2, D), np.float32); si = np.zeros((4, 2, S), np.uint8)
    O.latent_writeback([env, time], stoch, deter, sd, si)
    np.testing.assert_array_equal(sd[2, 1], deter[0, 1]); np.testing.assert_array_equal(si[2, 1], [2, 3])
    np.testing.assert_array_equal(sd[2, 0], deter[0, 2]); np.testing.assert_array_equal(si[2, 0], [1, 1])


def test_latent_store_has_no_cpu_fallback():
    """The product path fails loudly off the GPU instead of falling back to torch indexing."""
    from safe_dreamer_b200.replay import LatentStore
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        LatentStore(8, 2, 4, 8, 12, device="cpu")
    with pytest.raises(ValueError, match="uint8"):
        LatentStore(8, 2, 4, 300, 12, device="cpu")


@pytest.mark.gpu
def test_cuda_latent_store_golden_and_duplicates(golden):
    import torch
    from safe_dreamer_b200.replay import LatentStore
    env, time, stoch, deter, store_stoch, store_deter = _inputs()
    n_len, n_env, S, K = store_stoch.shape
    ls = LatentStore(n_len, n_env, S, K, deter.shape[-1], keep_onehot=True)
    ls.onehot.copy_(torch.from_numpy(store_stoch)); ls.deter.copy_(torch.from_numpy(store_deter))
    index = [torch.from_numpy(env).cuda(), torch.from_numpy(time).cuda()]
    ls.update(index, torch.from_numpy(stoch).cuda(), torch.from_numpy(deter).cuda())
    np.testing.assert_array_equal(ls.onehot.cpu().numpy(), golden["store_stoch"])
    np.testing.assert_array_equal(ls.deter.cpu().numpy(), golden["store_deter"])
    st, dt = ls.initial(index)
    np.testing.assert_array_equal(st.cpu().numpy(), stoch.reshape(-1, S, K))
    np.testing.assert_array_equal(dt.cpu().numpy(), deter.reshape(-1, deter.shape[-1]))
    # repeated slots: the later row wins (oracle = sequential assignment)
    rng = np.random.Generator(np.random.Philox(7))
    R = 300
    e2 = rng.integers(0, n_env, size=(1, R)); t2 = rng.integers(0, 6, size=(1, R))     # 24 slots, many repeats
    s2 = np.eye(K, dtype=np.float32)[rng.integers(0, K, size=(1, R, S))]
    d2 = rng.standard_normal((1, R, deter.shape[-1]), dtype=np.float32)
    od = ls.deter.cpu().numpy().copy(); oi = ls.idx.cpu().numpy().copy(); oo = ls.onehot.cpu().numpy().copy()
    O.latent_writeback([e2, t2], s2, d2, od, oi, oo)
    ls.update([torch.from_numpy(e2), torch.from_numpy(t2)], torch.from_numpy(s2).cuda(), torch.from_numpy(d2).cuda())
    np.testing.assert_array_equal(ls.deter.cpu().numpy(), od)
    np.testing.assert_array_equal(ls.idx.cpu().numpy(), oi)
    np.testing.assert_array_equal(ls.onehot.cpu().numpy(), oo)
    # out-of-range index: skipped and reported like the reference's IndexError
    before = ls.deter.clone()
    with pytest.raises(IndexError):
        ls.update([torch.tensor([[0, n_env]]), torch.tensor([[1, 1]])], torch.from_numpy(s2[:, :2]).cuda(),
                  torch.from_numpy(d2[:, :2] + 5).cuda())
    bad_free = before.clone(); bad_free[1, 0] = torch.from_numpy(d2[0, 0] + 5).cuda()
    assert torch.equal(ls.deter, bad_free)


@pytest.mark.gpu
def test_cuda_latent_store_full_size_roundtrip():
    """BASELINE sizes: B*T = 1024 rows of (32 x 16, 2048) into a (1000, 16) storage; write -> gather is the identity on
    one-hot rows, untouched slots keep their contents, the index storage holds the arg-max classes."""
    import torch
    from safe_dreamer_b200.replay import LatentStore
    B, T, S, K, D, n_len, n_env = 16, 64, 32, 16, 2048, 1000, 16
    g = torch.Generator().manual_seed(5)
    ls = LatentStore(n_len, n_env, S, K, D)
    ls.deter.fill_(-3.0)
    t0 = torch.randint(0, n_len - T, (B,), generator=g)
    time = t0[:, None] + torch.arange(T)[None]
    env = torch.arange(B)[:, None].expand(B, T)            # one slice per environment: all slots distinct
    cls = torch.randint(0, K, (B, T, S), generator=g)
    stoch = torch.nn.functional.one_hot(cls, K).float().cuda()
    deter = torch.randn(B, T, D, generator=g).cuda()
    ls.update([env, time], stoch, deter)
    st, dt = ls.initial([env, time])
    assert torch.equal(st, stoch.reshape(-1, S, K)) and torch.equal(dt, deter.reshape(-1, D))
    assert torch.equal(ls.idx[time.reshape(-1), env.reshape(-1)].cpu().long(), cls.reshape(-1, S))
    touched = torch.zeros(n_len, n_env, dtype=torch.bool); touched[time.reshape(-1), env.reshape(-1)] = True
    assert bool((ls.deter[~touched.cuda()] == -3.0).all()) and int(touched.sum()) == B * T
