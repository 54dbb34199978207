"""Counter-based draws of the vector environments (no GPU): the tensor expression of `envs/vector.py` against the library's
`rbc_checkpoint_draw` (the function the step kernel evaluates), and the noise generator's statistics."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")


def test_python_draw_equals_library_draw():
    from rbc_gym_b200 import backend
    from rbc_gym_b200.envs.vector import _VectorBase
    L = backend.load_library()

    class Stub(_VectorBase):
        pass
    s = Stub()
    s.torch, s.seed = torch, 12345
    s.env_ids = torch.arange(5000, 5064, dtype=torch.int64)
    s._episode = torch.arange(64, dtype=torch.int64) * 7
    s.sim = type("S", (), {"n_episodes": 20})()
    got = s._draw_checkpoints(torch.arange(64)).tolist()
    want = [L.rbc_checkpoint_draw(12345, 5000 + i, 7 * i, 20) for i in range(64)]
    assert got == want and len(set(got)) > 10
    assert L.rbc_checkpoint_draw(1, 2, 3, 0) == -1


def test_counter_normal_is_standard_normal_and_stateless():
    from rbc_gym_b200.envs.vector import counter_normal
    keys = torch.tensor([11, 12, 2 ** 40 + 5], dtype=torch.int64)
    x = counter_normal(torch, keys, 200_000, 0)
    assert x.shape == (3, 200_000) and torch.isfinite(x).all()
    assert abs(float(x.mean())) < 5e-3 and abs(float(x.std()) - 1) < 5e-3
    assert abs(float((x ** 4).mean()) - 3) < 0.1                                   # kurtosis of a Gaussian
    assert abs(float(np.corrcoef(x[0, :-1].numpy(), x[0, 1:].numpy())[0, 1])) < 0.01
    assert abs(float(np.corrcoef(x[0].numpy(), x[1].numpy())[0, 1])) < 0.01       # neighbouring keys are independent
    y = counter_normal(torch, keys[1:2], 1000, 0)
    assert torch.equal(y[0], x[1, :1000])                                          # same key, same numbers, any subset
    assert not torch.equal(counter_normal(torch, keys[:1], 1000, 1)[0], x[0, :1000])
