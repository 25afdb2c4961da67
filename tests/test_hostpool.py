"""Host scoring pool (CPU): parallel, ordered, same values as the serial loop."""
import numpy as np
import torch

from diffndm_b200.hostpool import PooledReward, radius_of_gyration_score


def _batch(seed=0):
    rng = np.random.default_rng(seed)
    sizes = rng.integers(1, 30, size=37)
    x = torch.from_numpy(rng.normal(size=(int(sizes.sum()), 3)).astype(np.float32))
    types = torch.from_numpy(rng.integers(0, 10, size=int(sizes.sum())))
    mask = torch.from_numpy(np.repeat(np.arange(len(sizes)), sizes))
    return x, types, mask, sizes


def test_pool_matches_serial_and_keeps_order():
    x, types, mask, sizes = _batch()
    serial = PooledReward(radius_of_gyration_score, workers=0)(x, types, mask)
    assert len(serial) == len(sizes)
    ref = [radius_of_gyration_score(x[mask == i].numpy(), types[mask == i].numpy()) for i in range(len(sizes))]
    assert np.allclose(serial, ref, atol=1e-6)
    with PooledReward(radius_of_gyration_score, workers=2, chunk=5) as pool:
        par = pool(x, types, mask)
        assert par == serial
        assert pool(x, types, mask) == serial            # reusable


def test_split_handles_single_molecule():
    x = np.zeros((4, 3), np.float32)
    mols = PooledReward.split(x, np.arange(4), np.zeros(4, np.int64))
    assert len(mols) == 1 and mols[0][0].shape == (4, 3)


def test_submit_is_non_blocking_and_ordered():
    x, types, mask, sizes = _batch(1)
    with PooledReward(radius_of_gyration_score, workers=2, chunk=4) as pool:
        h1 = pool.submit(x, types, mask)
        h2 = pool.submit(x, types, mask)
        assert h1.result() == h2.result() == pool(x, types, mask)
        assert h1.result() is h1.result()                                   # cached
    serial = PooledReward(radius_of_gyration_score, workers=0)
    assert serial.submit(x, types, mask).result() == serial(x, types, mask)
